#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts check
timeout 200 python tools/variant_check.py /tmp/base.npz --lib build/var_base/librosdyn_b200.so
timeout 200 python tools/variant_check.py /tmp/new.npz && python tools/variant_check.py --compare /tmp/base.npz /tmp/new.npz 2>&1 | tee gpurun_out/r02_tail_check.log
ts tests
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py tests/test_fold.py -m gpu -x -q -k "gram or Gram or fold or identification or sharded or group" 2>&1 | tail -5 | tee gpurun_out/r02_tail_pytest.log
ts gram
{ timeout 150 python tools/bench_gram.py 64000000 8 --lib build/var_base/librosdyn_b200.so | sed "s/^/base /"
timeout 150 python tools/bench_gram.py 64000000 8 | sed "s/^/tail /"
} 2>&1 | tee gpurun_out/r02_tail_bench.log
ts done
