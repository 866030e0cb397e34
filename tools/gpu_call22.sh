#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 ./build/dmma_pacing 2>&1 | tee gpurun_out/r02_micro_dmma_pacing.txt
