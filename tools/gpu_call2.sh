#!/bin/bash
# round-2 GPU call 2: ring kernel correctness + A/B against the slot kernel, datapath-sharing microbenchmark, bench.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts micro; timeout 120 build/dfma_halfwarp > gpurun_out/r02_micro_halfwarp.txt 2>&1; cat gpurun_out/r02_micro_halfwarp.txt
ts ring-tests; timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py tests/test_solve.py -m gpu -x -q \
   -k "gram or group_single or sharded or identification or eigen or cpp" > gpurun_out/r02_pytest_ring.log 2>&1; tail -4 gpurun_out/r02_pytest_ring.log
ts ab
for impl in ring slots; do
  RDB_GRAM_IMPL=$impl timeout 120 python tools/bench_gram.py 16000000 5 --lib build/var_dev/librosdyn_b200.so 2>&1 | sed "s/^/$impl /"
done > gpurun_out/r02_gram_ab.log 2>&1
cat gpurun_out/r02_gram_ab.log
ts bench; timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -c 1500 gpurun_out/r02_bench_b.json; tail -3 gpurun_out/r02_bench_b.err
ts done
