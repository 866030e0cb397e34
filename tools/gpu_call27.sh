#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{ for c in c6 c7; do
timeout 150 python tools/bench_ext.py $c --lib build/var_head/librosdyn_b200.so | sed "s/^/head /"
timeout 150 python tools/bench_ext.py $c --lib build/var_x0/librosdyn_b200.so | sed "s/^/x0 /"
timeout 150 python tools/bench_ext.py $c --lib build/var_x184/librosdyn_b200.so | sed "s/^/x184 /"
done; } 2>&1 | grep ext | tee gpurun_out/r02_ext4_bench.log
