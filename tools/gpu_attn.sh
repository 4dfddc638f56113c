#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "attention or alpha_one or window_attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_attn.log
timeout -s KILL 300 python tools/kbench.py attn --iters 10 > gpurun_out/kbench_attn.log 2>&1; echo "kbench exit $?" >> gpurun_out/kbench_attn.log
tail -n 40 gpurun_out/pytest_attn.log; tail -n 30 gpurun_out/kbench_attn.log
