#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{ timeout 150 python tools/bench_gram.py 64000000 6 | sed "s/^/base /"
for k in 1 2 3 4 5; do timeout 150 python tools/bench_gram.py 64000000 6 --lib build/var_pace$k/librosdyn_b200.so | sed "s/^/pace$k /"; done
timeout 150 python tools/bench_ext.py c6 | sed "s/^/base /"
for k in 1 3 4; do timeout 150 python tools/bench_ext.py c6 --lib build/var_pace$k/librosdyn_b200.so | sed "s/^/pace$k /"; done
} 2>&1 | tee gpurun_out/r02_pace_bench.log
