#!/bin/bash
cd "$(dirname "$0")/.."
for t in 2 4 8 12 16; do
  RDB_HOST_THREADS=$t timeout 300 python bench.py --steps 5 --warmup 3 --pageable --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('gram threads $t', round(d['e2e']['value']/1e6,1), 'M/s', round(d['e2e']['link_GBps_per_gpu'],1), 'GB/s')"
done
for t in 4 8 16; do
  RDB_HOST_THREADS=$t timeout 300 python bench.py --steps 5 --warmup 3 --workload materialise --pageable --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('mat threads $t', round(d['e2e']['value']/1e6,2), 'M/s', round(d['e2e']['link_GBps_per_gpu'],1), 'GB/s')"
done
