#!/bin/bash
# compute-sanitizer over the GPU test suite ("race detection"):  tools/sanitize.sh OUT
#   memcheck   out-of-bounds / misaligned accesses and leaks of device allocations (every -m gpu test but the full-size frames)
#   racecheck  shared-memory hazards (scene staging, static shared arrays of the path kernel)
#   synccheck  invalid __syncthreads / __syncwarp / vote usage (the warp-converged pair votes of brute_select_pkc)
#   initcheck  reads of uninitialised device memory
# The full-size BASELINE frames are left out (the tools slow kernels down 10-100x); every kernel and code path they
# use is also exercised by the small frames.
out=${1:-gpurun_out/sanitizer.txt}
: > "$out"
SEL='not full_size and not c5_size and not reference_rng'
for tool in memcheck racecheck synccheck initcheck; do
  echo "== compute-sanitizer --tool $tool: pytest tests -m gpu -k \"$SEL\"" >> "$out"
  extra=""
  [[ $tool == memcheck ]] && extra="--leak-check full"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 20 --error-exitcode 7 \
      python -m pytest tests -m gpu -x -q -k "$SEL" -p no:cacheprovider > gpurun_out/san_$tool.log 2>&1
  echo "exit code $?" >> "$out"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY|passed|failed|=========     [A-Z]" gpurun_out/san_$tool.log | sort | uniq -c | sort -rn | head -25 >> "$out"
done
