#!/bin/bash
# full GPU validation: smoke, all gpu tests, bench (logs under gpurun_out/)
mkdir -p gpurun_out
timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -n 3 gpurun_out/smoke.log; tail -n 25 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "e2e", "cpu_baseline", "clocks")})
    for r in d["rooflines"]: print(r)
    print(d["per_op_ms"])
except Exception as e:
    print("bench parse failed", e)
PY
