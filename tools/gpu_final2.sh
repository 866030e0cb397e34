#!/bin/bash
# round-2 closing check of the final tree on one GPU: full -m gpu suite, smoke, default bench line, reference arm, memcheck over the mapped host path
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts pytest; ( time timeout 900 python -m pytest tests -m gpu -q ) > $O/r02_pytest_gpu_final2.log 2>&1; tail -4 $O/r02_pytest_gpu_final2.log
ts smoke; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ts bench-default; timeout 400 python bench.py --steps 20 --warmup 3 > $O/r02_bench_default_final2.json 2> $O/r02_bench_default_final2.err; tail -c 300 $O/r02_bench_default_final2.json
ts bench-ref; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm_final2.json 2>/dev/null; tail -c 200 $O/r02_bench_reference_arm_final2.json
ts memcheck
timeout 600 compute-sanitizer --tool memcheck --log-file $O/r02_sanitizer_memcheck_mapped.log python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "small_host or cpp_eigen" > $O/r02_sanitizer_memcheck_mapped.pytest.log 2>&1
tail -2 $O/r02_sanitizer_memcheck_mapped.pytest.log; grep -E "ERROR SUMMARY" $O/r02_sanitizer_memcheck_mapped.log | tail -2
ts done
