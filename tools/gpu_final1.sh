#!/bin/bash
# round-2 final single-GPU evidence run: full -m gpu suite, bench lines, BASELINE-sized configs, ncu launch lists + full captures, sanitizers
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*"; }
ts pytest; ( time timeout 900 python -m pytest tests -m gpu -q ) > $O/r02_pytest_gpu_final.log 2>&1; tail -4 $O/r02_pytest_gpu_final.log
ts bench-default; timeout 400 python bench.py --steps 20 --warmup 3 > $O/r02_bench_default.json 2> $O/r02_bench_default.err; tail -c 300 $O/r02_bench_default.json
ts bench-mat; timeout 400 python bench.py --steps 20 --warmup 3 --workload materialise --no-cpu-baseline > $O/r02_bench_materialise.json 2> $O/r02_bench_materialise.err; tail -c 200 $O/r02_bench_materialise.json
ts bench-ref; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; tail -c 200 $O/r02_bench_reference_arm.json
ts bench-pageable; timeout 400 python bench.py --steps 5 --warmup 3 --pageable --no-cpu-baseline > $O/r02_bench_pageable.json 2> $O/r02_bench_pageable.err; tail -c 200 $O/r02_bench_pageable.json
ts configs; timeout 600 python tools/bench_configs.py --configs 1,2,3,4,5,h,g > $O/r02_configs.jsonl 2> $O/r02_configs.err; wc -l $O/r02_configs.jsonl; tail -2 $O/r02_configs.err
ts ext; for c in c6 c7; do timeout 120 python tools/bench_ext.py $c; done > $O/r02_ext_bench.log 2>&1; cat $O/r02_ext_bench.log
ts launches
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_gram.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02_launches_gram.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_materialise.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload materialise --samples 16000000 > $O/r02_launches_materialise.log 2>&1
ts ncu-full
# the reports (15 - 55 MB each with the source pages) are summarised HERE and deleted: gpurun_out/ travels back only below 64 MiB
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, note, regex, skip, command...
  local name=$1 note=$2 regex=$3 skip=$4; shift 4
  timeout 300 $NCU -k regex:$regex -s $skip -c 1 -o $O/$name "$@" > $O/$name.log 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep $O/${name}_ncu.txt "$note" > /dev/null 2>&1
  python tools/ncu_regions.py $O/$name.ncu-rep >> $O/${name}_ncu.txt 2>/dev/null
  ncu -i $O/$name.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum 2>/dev/null | tail -2 > $O/${name}_traffic.csv
  rm -f $O/$name.ncu-rep
  tail -3 $O/${name}_ncu.txt
}
cap r02_gram_fused_c6 "gram_fused_kernel<6,4,REV,0,4>: folded C6, zero mass column dropped (98 DMMA per 4 samples), 8 M samples" gram_fused_kernel 3 python tools/bench_gram.py 8000000 1
cap r02_gram_fused_c7 "gram_fused_kernel<7,3,REV,0,4>: C7, 3 slots, 4 generator warps (counter handshake), 143 DMMA per 4 samples, 8 M samples" gram_fused_kernel 7 python tools/bench_gram.py 8000000 1
cap r02_gram_ext_c6 "gram_ext_kernel<6,3,4,REV>: C6 + friction component on every joint, one pass (140 DMMA per 4 samples), 8 M samples" gram_ext_kernel 2 python tools/bench_ext.py c6
cap r02_kin_cfg2 "kin_kernel<7,CFG2,NP>: config 2 outputs, 16 M samples" kin_kernel 3 python tools/bench_kin.py
cap r02_dyn_regressor "dyn_kernel<7,REGRESSOR|TORQUE>: materialised regressor + torque (SoA planes), 4 M samples per launch" dyn_kernel 3 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload materialise --samples 8000000
ts sanitizers
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file $O/r02_sanitizer_$tool.log python -m pytest tests/test_parity_gpu.py tests/test_components.py tests/test_ik.py tests/test_round2_gpu.py -m gpu -q -x \
     -k "(long_double_oracle and not 1e6) or gram_folded or extended_gram or components_gpu or gpu_ik_against_oracle or group_single or eigen_record" > $O/r02_sanitizer_$tool.pytest.log 2>&1
  tail -2 $O/r02_sanitizer_$tool.pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/r02_sanitizer_$tool.log | tail -2
done
du -sh $O; find $O -size +20M -exec rm -v {} \;
ts done
