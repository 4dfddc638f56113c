#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gdn or rounding" > gpurun_out/pytest_gdn.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gdn.log
timeout 600 python tools/kbench.py gdn round --iters 10 > gpurun_out/kbench_gdn.log 2>&1; echo "kbench exit $?" >> gpurun_out/kbench_gdn.log
tail -n 12 gpurun_out/pytest_gdn.log; tail -n 30 gpurun_out/kbench_gdn.log
