#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 ./build/dfma_halfwarp 2>&1 | tee gpurun_out/r02_micro_datapath_sharing_v2.txt | tail -8
{ timeout 150 python tools/bench_gram.py 64000000 8 | sed "s/^/base /"
timeout 150 python tools/bench_gram.py 64000000 8 --lib build/var_genfirst/librosdyn_b200.so | sed "s/^/genfirst /"; } 2>&1 | tee gpurun_out/r02_genfirst_bench.log
