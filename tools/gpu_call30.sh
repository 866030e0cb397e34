#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for c in c6 c7; do
name=r02_gram_ext_${c}
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:gram_ext_kernel -s 2 -c 1 -o $O/$name python tools/bench_ext.py $c > $O/$name.log 2>&1
python tools/ncu_summary.py $O/$name.ncu-rep $O/${name}_ncu.txt "gram_ext_kernel: $c + friction component on every joint, one pass; rigid-body slots (zero mass column dropped) + 2 component columns per joint in a side buffer; 8 M samples" > /dev/null 2>&1
python tools/ncu_regions.py $O/$name.ncu-rep >> $O/${name}_ncu.txt 2>/dev/null
rm -f $O/$name.ncu-rep
grep -E "^generator|^mma|dmma.avg|pipe_fp64.avg|time_duration|registers_per_thread " $O/${name}_ncu.txt
done
