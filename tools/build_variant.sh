#!/bin/bash
# Build an A/B variant of librt_b200.so: only rt_f32.cu is recompiled with the extra flags, the other objects are reused.
#   tools/build_variant.sh NAME [extra nvcc flags...]   ->  build/ab/librt_NAME.so   (select with RT_B200_LIB=...)
set -e
name=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
src=$here/ray-tracer-v1_b200/csrc
out=$here/build/ab
mkdir -p "$out"
ARCH="-gencode arch=compute_100a,code=sm_100a"
(cd "$src" && make -s librt_b200.so >/dev/null)
nvcc $ARCH -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v "$@" -c "$src/rt_f32.cu" -o "$out/rt_f32_$name.o" 2> "$out/$name.ptxas.log"
nvcc $ARCH -shared -o "$out/librt_$name.so" "$src/rt_api.o" "$out/rt_f32_$name.o" "$src/rt_f64.o" "$src/rt_lbvh.o" -Xcompiler -fPIC -cudart static
grep -A3 "path_kernelIfLi[03]ELb1ELb0" "$out/$name.ptxas.log" | grep Used | sed "s/^/$name: /"
