#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/variant_check.py /tmp/base.npz --lib build/var_nopf/librosdyn_b200.so
timeout 200 python tools/variant_check.py /tmp/new.npz && python tools/variant_check.py --compare /tmp/base.npz /tmp/new.npz | grep -c True
{ timeout 150 python tools/bench_gram.py 64000000 8 --lib build/var_nopf/librosdyn_b200.so | sed "s/^/nopf /"
timeout 150 python tools/bench_gram.py 64000000 8 | sed "s/^/pf /"
for c in c6 c7; do timeout 150 python tools/bench_ext.py $c --lib build/var_nopf/librosdyn_b200.so | sed "s/^/nopf /"; timeout 150 python tools/bench_ext.py $c | sed "s/^/pf /"; done | grep ext
} 2>&1 | tee gpurun_out/r02_prefetch_bench.log
