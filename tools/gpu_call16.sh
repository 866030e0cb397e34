#!/bin/bash
cd "$(dirname "$0")/.."
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts tests; timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py tests/test_solve.py -m gpu -x -q -k "(gram and not 1e6) or reproducible or identification or sharded or group_single" > gpurun_out/r02_pytest_rolled.log 2>&1; tail -4 gpurun_out/r02_pytest_rolled.log
ts timing
for v in cur rolled; do
  for dbg in 0 2; do RDB_GRAM_DEBUG=$dbg timeout 120 python tools/bench_gram.py 16000000 5 --lib build/var_$v/librosdyn_b200.so 2>&1 | sed "s/^/$v /"; done
done | tee gpurun_out/r02_rolled.log
ts done
