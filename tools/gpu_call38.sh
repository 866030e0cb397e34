#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "host or eigen_record or facade or cpp" 2>&1 | tail -3
cp rosdyn_b200/librosdyn_b200.so /tmp/new.so
timeout 400 python bench.py --steps 5 --warmup 3 --workload materialise --pageable --no-cpu-baseline > gpurun_out/r02_bench_materialise_pageable.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_materialise_pageable.json').read().strip().splitlines()[-1]); print('new pageable', d['e2e'])"
cp build/var_head/librosdyn_b200.so rosdyn_b200/librosdyn_b200.so
timeout 400 python bench.py --steps 5 --warmup 3 --workload materialise --pageable --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('old pageable', d['e2e'])"
cp /tmp/new.so rosdyn_b200/librosdyn_b200.so
timeout 400 python bench.py --steps 5 --warmup 3 --workload materialise --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('new pinned', d['e2e']['value'], d['e2e']['link_GBps_per_gpu'])"
