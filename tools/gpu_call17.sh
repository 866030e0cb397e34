#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts check
timeout 200 python tools/variant_check.py /tmp/base.npz
for v in rp184 rp152 rp136; do
  timeout 200 python tools/variant_check.py /tmp/$v.npz --lib build/var_$v/librosdyn_b200.so && python tools/variant_check.py --compare /tmp/base.npz /tmp/$v.npz
done 2>&1 | tee gpurun_out/r02_repart_check.log
ts gram
{ timeout 150 python tools/bench_gram.py 64000000 8 | sed "s/^/base /"
for v in rp152 rp136; do timeout 150 python tools/bench_gram.py 64000000 8 --lib build/var_$v/librosdyn_b200.so | sed "s/^/$v /"; done
ts ext
for c in c6 c7; do
timeout 150 python tools/bench_ext.py $c | sed "s/^/base /"
timeout 150 python tools/bench_ext.py $c --lib build/var_rp184/librosdyn_b200.so | sed "s/^/rp184 /"
done
ts kin
timeout 150 python tools/bench_kin.py
timeout 150 python tools/bench_kin.py --lib build/var_rp136/librosdyn_b200.so
} 2>&1 | tee gpurun_out/r02_repart_bench.log
ts done
