#!/bin/bash
# A/B the path kernel across variant libraries (tools/build_variant.sh) on the GPU box:  tools/ab.sh OUT name[:ENV=1] ...
out=$1; shift
: > "$out"
for spec in "$@"; do
  name=${spec%%:*}; envs=""
  [[ "$spec" == *:* ]] && envs=${spec#*:}
  lib=build/ab/librt_$name.so
  [[ "$name" == "cur" ]] && lib=ray-tracer-v1_b200/csrc/librt_b200.so
  for cfg in "complex 16" "complex 64" "chandelier 16"; do
    set -- $cfg
    echo "== $spec $cfg" >> "$out"
    env $envs RT_B200_LIB=$PWD/$lib python tools/time_path.py --scene $1 --spp $2 --reps 4 --schedules 0 $AB_ARGS 2>&1 | grep -v "^$" >> "$out"
  done
done
