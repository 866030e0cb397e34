#!/bin/bash
cd "$(dirname "$0")/.."
cp rosdyn_b200/librosdyn_b200.so /tmp/new.so
for rep in 1 2; do
cp /tmp/new.so rosdyn_b200/librosdyn_b200.so
timeout 400 python bench.py --steps 5 --warmup 3 --workload materialise --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('new pinned', d['e2e']['value'], d['e2e']['link_GBps_per_gpu'])"
cp build/var_head/librosdyn_b200.so rosdyn_b200/librosdyn_b200.so
timeout 400 python bench.py --steps 5 --warmup 3 --workload materialise --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('old pinned', d['e2e']['value'], d['e2e']['link_GBps_per_gpu'])"
done
cp /tmp/new.so rosdyn_b200/librosdyn_b200.so
