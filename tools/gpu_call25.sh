#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts tests
timeout 600 python -m pytest tests -m gpu -x -q -k "extended or components or cross or ext" 2>&1 | tail -4
ts check
timeout 200 python tools/variant_check.py /tmp/base.npz --lib build/var_head/librosdyn_b200.so
timeout 200 python tools/variant_check.py /tmp/new.npz && python tools/variant_check.py --compare /tmp/base.npz /tmp/new.npz
ts bench
{ for c in c6 c7; do
timeout 150 python tools/bench_ext.py $c --lib build/var_head/librosdyn_b200.so | sed "s/^/head /"
timeout 150 python tools/bench_ext.py $c --lib build/var_e0/librosdyn_b200.so | sed "s/^/e0 /"
timeout 150 python tools/bench_ext.py $c | sed "s/^/e184 /"
timeout 150 python tools/bench_ext.py $c --lib build/var_e192/librosdyn_b200.so | sed "s/^/e192 /"
done; } 2>&1 | tee gpurun_out/r02_ext2_bench.log
ts done
