#!/bin/bash
# development: correctness + timing + one full ncu capture (source-level stall sampling) of the ws attention kernel
mkdir -p gpurun_out
bash tools/gpu_ws.sh
if grep -q "dbg exit 0" gpurun_out/dbg_attn.log; then
  timeout -s KILL 120 python tools/prof_one.py attn8 tc > gpurun_out/prof_plain.log 2>&1 && \
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:mwa_ws_kernel -s 2 -c 1 -f -o gpurun_out/prof_ws python tools/prof_one.py attn8 tc > gpurun_out/ncu_ws.log 2>&1
  tail -n 3 gpurun_out/ncu_ws.log
fi
