#!/bin/bash
# ncu evidence for the bench command: launch list (durations) + one full capture of each top kernel
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/bench_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gdn_tc_kernel|mwa_tc_kernel" -s 12 -c 4 -o gpurun_out/prof_bench -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/ncu_launch.log gpurun_out/ncu_full.log; wc -l gpurun_out/launches.csv
