#!/bin/bash
cd "$(dirname "$0")/.."
for v in kold knew km3 kold knew; do timeout 120 python tools/bench_kin.py --lib build/var_$v/librosdyn_b200.so; done 2>&1 | tee gpurun_out/r02_kin_variants.log
