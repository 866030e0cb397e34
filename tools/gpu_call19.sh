#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in t0 t4 t2; do for dbg in 0 1 2; do
  RDB_GRAM_DEBUG=$dbg timeout 120 python tools/bench_gram.py 32000000 6 --lib build/var_dev_$v/librosdyn_b200.so 2>&1 | sed "s/^/$v /"
done; done | tee gpurun_out/r02_tail_roles.log
