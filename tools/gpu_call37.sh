#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "host or identification or sharded" 2>&1 | tail -3
timeout 400 python bench.py --steps 5 --warmup 3 --pageable --no-cpu-baseline > gpurun_out/r02_bench_pageable.json 2>gpurun_out/r02_bench_pageable.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_pageable.json').read().strip().splitlines()[-1]); print('pageable', d['e2e'])"
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_pinned_check.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_pinned_check.json').read().strip().splitlines()[-1]); print('pinned', d['e2e']['value'], d['e2e']['frac_of_h2d'])"
