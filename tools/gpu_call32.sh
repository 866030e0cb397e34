#!/bin/bash
cd "$(dirname "$0")/.."
timeout 200 python tools/variant_check.py /tmp/base.npz
timeout 200 python tools/variant_check.py /tmp/new.npz --lib build/var_after/librosdyn_b200.so && python tools/variant_check.py --compare /tmp/base.npz /tmp/new.npz | grep -c True
for c in c6 c7; do timeout 150 python tools/bench_ext.py $c | sed "s/^/cur /"; timeout 150 python tools/bench_ext.py $c --lib build/var_after/librosdyn_b200.so | sed "s/^/after /"; done | grep ext
