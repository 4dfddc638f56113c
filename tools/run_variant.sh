#!/bin/bash
# development: run tools/prof_one.py against a variant library:  tools/run_variant.sh <variant> <prof_one args...>
v=$1; shift
CUDA_LAUNCH_BLOCKING=1 timeout -s KILL 100 python - "$@" <<PY 2>&1 | tail -3
import sys, os
sys.path.insert(0, ".")
import mwa_b200 as pkg
pkg._abi.LIB_PATH = os.path.abspath("build/variants/$v.so")
sys.argv = ["prof_one.py"] + sys.argv[1:]
exec(open("tools/prof_one.py").read())
PY
