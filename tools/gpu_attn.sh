#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 ./build/umma_probe > gpurun_out/probe.log 2>&1; echo "probe exit $?" >> gpurun_out/probe.log
timeout -s KILL 200 python tools/dbg_attn.py > gpurun_out/dbg_attn.log 2>&1; echo "dbg exit $?" >> gpurun_out/dbg_attn.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "attention or alpha_one or window_attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_attn.log
timeout -s KILL 300 python tools/kbench.py attn --iters 10 > gpurun_out/kbench_attn.log 2>&1; echo "kbench exit $?" >> gpurun_out/kbench_attn.log
timeout -s KILL 120 python tools/phase_times.py attn8 > gpurun_out/phase8.log 2>&1
timeout -s KILL 120 python tools/phase_times.py attn4 > gpurun_out/phase4.log 2>&1
tail -n 8 gpurun_out/probe.log; tail -n 9 gpurun_out/dbg_attn.log; tail -n 4 gpurun_out/pytest_attn.log; grep -v simt gpurun_out/kbench_attn.log | tail -n 13; cat gpurun_out/phase8.log; head -4 gpurun_out/phase4.log
