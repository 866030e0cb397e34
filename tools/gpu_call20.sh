#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts latency
timeout 120 ./build/facade_check 2>&1 | tee gpurun_out/r02_latency_mapped.txt | tail -8
ts tests
timeout 900 python -m pytest tests -m gpu -x -q -k "small_host or host_buffer or cpp or eigen or facade or handle or threads" 2>&1 | tail -5 | tee gpurun_out/r02_mapped_pytest.log
ts done
