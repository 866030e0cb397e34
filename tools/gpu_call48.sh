#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --pageable --no-cpu-baseline > gpurun_out/r02_bench_pageable.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_pageable.json').read().strip().splitlines()[-1]); print('gram', d['e2e']['value']/1e6, d['e2e']['frac_of_h2d'])"
timeout 300 python bench.py --steps 5 --warmup 3 --workload materialise --pageable --no-cpu-baseline > gpurun_out/r02_bench_materialise_pageable.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_materialise_pageable.json').read().strip().splitlines()[-1]); print('mat', d['e2e']['value']/1e6, d['e2e']['link_GBps_per_gpu'])"
timeout 600 python -m pytest tests -m gpu -q -k "host or pageable or sharded" 2>&1 | tail -2
