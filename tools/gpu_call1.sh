#!/bin/bash
# round-2 GPU call 1: full -m gpu suite, sustained bench, BASELINE-sized configs, kernel diagnostics, sanitizer logs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt 2>&1
nproc >> gpurun_out/r02_gpu.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -c 600 gpurun_out/r02_bench_a.json
python tools/bench_configs.py --configs 1,2,3,4,5,h > gpurun_out/r02_configs_a.jsonl 2> gpurun_out/r02_configs_a.err
tail -3 gpurun_out/r02_configs_a.err
for dbg in 0 1 2; do RDB_GRAM_DEBUG=$dbg python tools/bench_gram.py 16000000 5 --lib build/var_dev/librosdyn_b200.so; done > gpurun_out/r02_gram_diag_a.log 2>&1
cat gpurun_out/r02_gram_diag_a.log
# compute-sanitizer on the fused / cross-mode / IK / group kernels (small batches)
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r02_sanitizer_$tool.log python -m pytest tests/test_parity_gpu.py tests/test_components.py tests/test_ik.py tests/test_round2_gpu.py -m gpu -q -x \
     -k "(long_double_oracle and not 1e6) or gram_folded or extended_gram or components_gpu or gpu_ik_against_oracle or group_single or eigen_record" > gpurun_out/r02_sanitizer_$tool.pytest.log 2>&1
  tail -3 gpurun_out/r02_sanitizer_$tool.pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02_sanitizer_$tool.log | tail -3
done
