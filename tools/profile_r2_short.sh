#!/bin/bash
# The 1-GPU core of tools/profile_r2.sh (bench, reference arm, launch list, one ncu --set full capture of the path kernel
# with its summaries), each ncu pass after its command has run clean without ncu:  tools/profile_r2_short.sh TAG
tag=${1:-r2s}
o=gpurun_out
set -x
python bench.py --steps 5 --warmup 3 > $o/bench_$tag.json 2> $o/bench_$tag.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $o/bench_ref_$tag.json 2> $o/bench_ref_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $o/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:path_kernel --launch-skip 1 --launch-count 1 -f \
    -o $o/prof_path_$tag python tools/time_path.py --scene complex --spp 64 --reps 1 --schedules 0 > $o/ncu_path_$tag.log 2>&1
python tools/ncu_summary.py $o/prof_path_$tag.ncu-rep > $o/ncu_path_summary_$tag.txt 2>&1
python tools/ncu_source.py $o/prof_path_$tag.ncu-rep lines 60 > $o/ncu_path_source_lines_$tag.txt 2>&1
python tools/ncu_stalls.py $o/prof_path_$tag.ncu-rep stall_math 12 > $o/ncu_path_stalls_$tag.txt 2>&1
ls -la $o/*_$tag*
