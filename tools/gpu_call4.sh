#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts variants
for v in dev unpaired; do
  RDB_GRAM_IMPL=ring timeout 120 python tools/bench_gram.py 16000000 5 --lib build/var_$v/librosdyn_b200.so 2>&1 | sed "s/^/$v /"
done > gpurun_out/r02_ring_variants2.log 2>&1
cat gpurun_out/r02_ring_variants2.log
ts tests; RDB_GRAM_IMPL=ring timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -x -q -k "gram and not 1e6" > gpurun_out/r02_pytest_ring2.log 2>&1; tail -3 gpurun_out/r02_pytest_ring2.log
ts ncu
RDB_GRAM_IMPL=ring timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_ring_kernel -s 3 -c 1 -f -o gpurun_out/r02_ring_v2 \
  python tools/bench_gram.py 2000000 1 --lib build/var_dev/librosdyn_b200.so > gpurun_out/r02_ring_v2.log 2>&1
tail -2 gpurun_out/r02_ring_v2.log
ts done
