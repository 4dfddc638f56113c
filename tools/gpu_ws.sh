#!/bin/bash
# development: correctness + timing of the warp-specialised attention kernel (each step under its own watchdog)
mkdir -p gpurun_out
timeout -s KILL 150 python tools/dbg_attn.py > gpurun_out/dbg_attn.log 2>&1; echo "dbg exit $?" >> gpurun_out/dbg_attn.log
tail -n 20 gpurun_out/dbg_attn.log
if grep -q "dbg exit 0" gpurun_out/dbg_attn.log; then
  timeout -s KILL 200 python tools/kbench.py attn --iters 10 --no-simt > gpurun_out/kbench_attn.log 2>&1; echo "kbench exit $?" >> gpurun_out/kbench_attn.log
  grep -v simt gpurun_out/kbench_attn.log | tail -n 16
  timeout -s KILL 100 python tools/phase_times.py attn8 > gpurun_out/phase8.log 2>&1; cat gpurun_out/phase8.log
  timeout -s KILL 100 python tools/phase_times.py attn4 > gpurun_out/phase4.log 2>&1; cat gpurun_out/phase4.log
fi
