#!/bin/bash
# round-2 closing check after the extended-model kernel change: full -m gpu suite, smoke, extended-model bench + ncu, sanitizers over its tests
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts pytest; ( time timeout 900 python -m pytest tests -m gpu -q ) > $O/r02_pytest_gpu_final3.log 2>&1; tail -4 $O/r02_pytest_gpu_final3.log
ts smoke; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ts ext; for c in c6 c7; do timeout 120 python tools/bench_ext.py $c; done > $O/r02_ext_bench.log 2>&1; cat $O/r02_ext_bench.log
ts ncu
bash tools/gpu_call30.sh
ts sanitizers
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file $O/r02_sanitizer_ext_$tool.log python -m pytest tests -m gpu -q -x -k "extended or components_gpu or cross" > $O/r02_sanitizer_ext_$tool.pytest.log 2>&1
  tail -1 $O/r02_sanitizer_ext_$tool.pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/r02_sanitizer_ext_$tool.log | tail -1
  grep -E "Hazard|hazard" $O/r02_sanitizer_ext_$tool.log | sed -E 's/.*(in|at) ([a-zA-Z_0-9:<>, ]+)\(.*/\2/' | sort | uniq -c | sort -rn | head -5
done
find $O -size +20M -exec rm -v {} \;
ts done
