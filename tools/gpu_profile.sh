#!/bin/bash
# ncu evidence for the bench command: launch list (durations) + full captures of the two top kernels, then the
# micro-benchmarks and the bench line itself (all numbers outside ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-clocks --no-graph"
$CMD > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/bench_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"mwa_ws_kernel" -s 8 -c 2 -o gpurun_out/prof_attn -f $CMD > gpurun_out/ncu_full_attn.log 2>&1
$CMD > gpurun_out/bench_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gdn_tc_kernel" -s 12 -c 3 -o gpurun_out/prof_gdn -f $CMD > gpurun_out/ncu_full_gdn.log 2>&1
tail -n 2 gpurun_out/ncu_full_attn.log gpurun_out/ncu_full_gdn.log | cut -c1-200; wc -l gpurun_out/launches.csv
timeout -s KILL 400 python tools/kbench.py attn gdn round --iters 10 --no-simt > gpurun_out/kbench_all.log 2>&1; tail -n 5 gpurun_out/kbench_all.log
timeout -s KILL 100 python tools/phase_times.py attn8 1.0 > gpurun_out/phase8_dense.log 2>&1
timeout -s KILL 100 python tools/phase_times.py attn8 0.5 > gpurun_out/phase8_sparse.log 2>&1
timeout -s KILL 300 python tools/train_bench.py > gpurun_out/train_bench.log 2>&1; tail -c 600 gpurun_out/train_bench.log
timeout -s KILL 400 python bench.py > gpurun_out/bench_full.log 2>&1; tail -c 1200 gpurun_out/bench_full.log
