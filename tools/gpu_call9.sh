#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_ext_kernel -s 2 -c 1 -f -o gpurun_out/r02_ext_v1 python tools/bench_ext.py c6 > gpurun_out/r02_ext_v1.log 2>&1
tail -3 gpurun_out/r02_ext_v1.log
