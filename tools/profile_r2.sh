#!/bin/bash
# Round-2 evidence run on ONE B200 (every ncu capture comes AFTER its command has run clean without ncu).
#   tools/profile_r2.sh TAG   -> gpurun_out/*_TAG.*
tag=${1:-r2}
o=gpurun_out
set -x
python bench.py --steps 5 --warmup 3 > $o/bench_$tag.json 2> $o/bench_$tag.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $o/bench_ref_$tag.json 2> $o/bench_ref_$tag.err
python tools/bench_configs.py > $o/configs_$tag.jsonl 2> $o/configs_$tag.err
python tools/time_env.py > $o/env_$tag.txt 2>&1
# launch list of the bench command (per-launch times are cold-cache and serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $o/ncu_launches_$tag.log 2>&1
# the dominant kernel, one 64-spp launch of the bench workload
ncu --set full --clock-control none --import-source on -k regex:path_kernel --launch-skip 1 --launch-count 1 -f \
    -o $o/prof_path_$tag python tools/time_path.py --scene complex --spp 64 --reps 1 --schedules 0 > $o/ncu_path_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:path_kernel --launch-skip 1 --launch-count 1 -f \
    -o $o/prof_chandelier_$tag python tools/time_path.py --scene chandelier --spp 64 --reps 1 --schedules 0 > $o/ncu_chand_$tag.log 2>&1
# secondary kernels: fused env step (FB + RL flavours), Algorithm-A frame passes
ncu --set full --clock-control none -k regex:env_step_kernel --launch-skip 30 --launch-count 2 -f \
    -o $o/prof_env_$tag python tools/time_env.py > $o/ncu_env_$tag.log 2>&1
ncu --set full --clock-control none -k regex:whitted --launch-skip 6 --launch-count 4 -f \
    -o $o/prof_whitted_$tag python tools/bench_configs.py --only C2 --no-cpu > $o/ncu_whitted_$tag.log 2>&1
# summaries are made on the box; only the dominant kernel's report travels back (gpurun_out is capped at 64 MiB)
for k in path chandelier env whitted; do
  python tools/ncu_summary.py $o/prof_${k}_$tag.ncu-rep > $o/ncu_${k}_summary_$tag.txt 2>&1
done
python tools/ncu_source.py $o/prof_path_$tag.ncu-rep lines 60 > $o/ncu_path_source_lines_$tag.txt 2>&1
python tools/ncu_stalls.py $o/prof_path_$tag.ncu-rep stall_math 12 > $o/ncu_path_stalls_$tag.txt 2>&1
rm -f $o/prof_chandelier_$tag.ncu-rep $o/prof_env_$tag.ncu-rep $o/prof_whitted_$tag.ncu-rep
ls -la $o/*_$tag*
