#!/bin/bash
# 2 GPUs: NCCL group through the C-ABI (python + C++), handles on another device, torchrun bench through rdb_group_create_rank
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts tests; timeout 400 python -m pytest tests/test_round2_gpu.py -m gpu -x -q -k "group or handle_keeps or two_host_threads" > gpurun_out/r02_pytest_2gpu.log 2>&1; tail -4 gpurun_out/r02_pytest_2gpu.log
ts cpp; timeout 200 build/group_check 2 4000003 2>&1 | tee gpurun_out/r02_group_check_2gpu.txt
ts bench2; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -c 1200 gpurun_out/r02_bench_n2.json; tail -5 gpurun_out/r02_bench_n2.err
ts done
