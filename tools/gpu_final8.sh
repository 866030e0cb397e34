#!/bin/bash
# round-2 final 8-GPU evidence run: NCCL group (C++ and python), bench.py at N = 8 / 4 (torchrun, through rdb_group_create_rank), BASELINE configs 4 / 5 / headline on 8 GPUs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*"; }
nvidia-smi topo -m > $O/r02_topo_8gpu.txt 2>&1; nproc >> $O/r02_topo_8gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
ts group_check; timeout 300 build/group_check 8 16000003 2>&1 | tee $O/r02_group_check_8gpu.txt
ts pytest; timeout 400 python -m pytest tests/test_round2_gpu.py -m gpu -x -q -k "group or handle_keeps" > $O/r02_pytest_8gpu.log 2>&1; tail -3 $O/r02_pytest_8gpu.log
ts bench8; timeout 500 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r02_bench_n8.json 2> $O/r02_bench_n8.err; tail -c 300 $O/r02_bench_n8.json; tail -2 $O/r02_bench_n8.err
ts bench4; timeout 500 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > $O/r02_bench_n4.json 2> $O/r02_bench_n4.err; tail -c 300 $O/r02_bench_n4.json
ts bench8-chunk; RDB_HOST_CHUNK=1048576 timeout 500 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_n8_chunk1m.json 2> $O/r02_bench_n8_chunk1m.err; tail -c 300 $O/r02_bench_n8_chunk1m.json
ts configs8; timeout 600 $TR --nproc-per-node 8 --master-port 29524 tools/bench_configs.py --configs 4,5,h > $O/r02_configs_n8.jsonl 2> $O/r02_configs_n8.err; cat $O/r02_configs_n8.jsonl | cut -c 1-400; tail -2 $O/r02_configs_n8.err
ts done
