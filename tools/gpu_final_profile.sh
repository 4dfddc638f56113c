#!/bin/bash
# end-of-round evidence: full ncu captures of the two top kernels of the bench command + the kernel micro-benchmarks
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-clocks --no-graph"
$CMD > gpurun_out/bench_plain2.log 2>&1 && \
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:"mwa_ws_kernel" -s 8 -c 2 -o gpurun_out/prof_attn -f $CMD > gpurun_out/ncu_full_attn.log 2>&1
$CMD > gpurun_out/bench_plain3.log 2>&1 && \
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:"gdn_tc_kernel" -s 12 -c 3 -o gpurun_out/prof_gdn -f $CMD > gpurun_out/ncu_full_gdn.log 2>&1
tail -n 2 gpurun_out/ncu_full_attn.log gpurun_out/ncu_full_gdn.log | cut -c1-200
timeout -s KILL 400 python tools/kbench.py attn gdn round gate pyramid --iters 10 --no-simt > gpurun_out/kbench_all.log 2>&1; tail -n 3 gpurun_out/kbench_all.log | cut -c1-200
