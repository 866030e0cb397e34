#!/bin/bash
# round-2 GPU call 3: why is the ring kernel slow?  variants + ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts variants
for v in dev nofence eager nofence_eager; do
  RDB_GRAM_IMPL=ring timeout 120 python tools/bench_gram.py 16000000 5 --lib build/var_$v/librosdyn_b200.so 2>&1 | sed "s/^/$v /"
done > gpurun_out/r02_ring_variants.log 2>&1
cat gpurun_out/r02_ring_variants.log
ts ncu
RDB_GRAM_IMPL=ring timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_ring_kernel -s 3 -c 1 -f -o gpurun_out/r02_ring_v1 \
  python tools/bench_gram.py 2000000 1 --lib build/var_dev/librosdyn_b200.so > gpurun_out/r02_ring_v1.log 2>&1
tail -2 gpurun_out/r02_ring_v1.log
ts done
