#!/bin/bash
cd "$(dirname "$0")/.."
for v in cur t1g8 t1g6; do
  for dbg in 0 2; do RDB_GRAM_DEBUG=$dbg timeout 120 python tools/bench_gram.py 16000000 5 --lib build/var_$v/librosdyn_b200.so 2>&1 | sed "s/^/$v /"; done
done | tee gpurun_out/r02_t1g8.log
