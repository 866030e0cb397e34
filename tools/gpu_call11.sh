#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts tests; timeout 400 python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "eigen or against_oracle_seeded or golden or host_buffer or cpp" > gpurun_out/r02_pytest_rec.log 2>&1; tail -4 gpurun_out/r02_pytest_rec.log
ts bench; timeout 300 python tools/bench_configs.py --configs h --min-seconds 0.5 2> gpurun_out/r02_cfg_h.err | tee gpurun_out/r02_cfg_h.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'][:90], '%.4g'%r['samples_per_s'], r.get('hbm_frac'), r.get('fp64_frac'), r['clocks']['sm_mhz'])"
tail -3 gpurun_out/r02_cfg_h.err
ts done
