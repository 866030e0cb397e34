#!/bin/bash
cd "$(dirname "$0")/.."
for dbg in 2 0; do RDB_GRAM_DEBUG=$dbg timeout 200 python tools/bench_gen_scaling.py --lib build/var_cur/librosdyn_b200.so; done 2>&1 | tee gpurun_out/r02_gen_scaling.log
