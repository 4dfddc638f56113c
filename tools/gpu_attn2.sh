#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python tools/dbg_attn.py > gpurun_out/dbg_attn.log 2>&1
timeout -s KILL 120 python tools/prof_one.py attn8 tc > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mwa_tc_kernel -s 2 -c 1 -o gpurun_out/prof_attn_v1 -f python tools/prof_one.py attn8 tc > gpurun_out/ncu_attn.log 2>&1
tail -n 12 gpurun_out/dbg_attn.log; tail -n 4 gpurun_out/ncu_attn.log
