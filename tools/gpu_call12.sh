#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts tests; timeout 500 python -m pytest tests/test_parity_gpu.py tests/test_wrench_jaclink.py tests/test_urdf.py tests/test_ik.py -m gpu -x -q > gpurun_out/r02_pytest_kin.log 2>&1; tail -4 gpurun_out/r02_pytest_kin.log
ts bench; timeout 300 python tools/bench_configs.py --configs 2,5 --min-seconds 1.0 2> gpurun_out/r02_cfg_25.err | tee gpurun_out/r02_cfg_25.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'][:90], '%.4g'%r['samples_per_s'], r.get('hbm_frac'), r['clocks']['sm_mhz'], r['clocks']['reasons'])"
tail -3 gpurun_out/r02_cfg_25.err
ts done
