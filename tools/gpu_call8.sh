#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts tests; timeout 400 python -m pytest tests/test_components.py tests/test_parity_gpu.py tests/test_solve.py -m gpu -x -q -k "extended or components or gram or identification" > gpurun_out/r02_pytest_ext.log 2>&1; tail -4 gpurun_out/r02_pytest_ext.log
ts bench; for c in c6 c7; do timeout 120 python tools/bench_ext.py $c; done 2>&1 | tee gpurun_out/r02_ext_bench.log
ts done
