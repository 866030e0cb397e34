#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts tests; RDB_GRAM_IMPL=slots timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py tests/test_components.py -m gpu -x -q -k "(gram and not 1e6) or extended" > gpurun_out/r02_pytest_z.log 2>&1; tail -3 gpurun_out/r02_pytest_z.log
ts variants
for v in noz dev ts1; do
  for dbg in 0 1 2; do
    RDB_GRAM_DEBUG=$dbg RDB_GRAM_IMPL=slots timeout 120 python tools/bench_gram.py 16000000 5 --lib build/var_$v/librosdyn_b200.so 2>&1 | sed "s/^/$v /"
  done
done > gpurun_out/r02_slot_variants.log 2>&1
cat gpurun_out/r02_slot_variants.log
ts done
