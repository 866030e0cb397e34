#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ns in 32 128 512; do
  for c in c6 c7; do timeout 120 python tools/bench_ext.py $c --lib build/var_ns$ns/librosdyn_b200.so 2>&1 | sed "s/^/ns$ns /"; done
done | tee gpurun_out/r02_spin_variants.log
