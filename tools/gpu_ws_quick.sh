#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 150 python tools/dbg_attn.py > gpurun_out/dbg_attn.log 2>&1; echo "dbg exit $?" >> gpurun_out/dbg_attn.log
grep -c "bad windows 0/" gpurun_out/dbg_attn.log; tail -n 1 gpurun_out/dbg_attn.log
timeout -s KILL 200 python tools/kbench.py attn --iters 10 --no-simt > gpurun_out/kbench_attn.log 2>&1; echo "kbench exit $?" >> gpurun_out/kbench_attn.log
grep "tcgen05 " gpurun_out/kbench_attn.log
