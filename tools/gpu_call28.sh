#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
name=r02_gram_ext3_c6
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:gram_ext_kernel -s 2 -c 1 -o $O/$name python tools/bench_ext.py c6 --lib build/var_e3/librosdyn_b200.so > $O/$name.log 2>&1
python tools/ncu_summary.py $O/$name.ncu-rep $O/${name}_ncu.txt "gram_ext_kernel (component fragments formed by the MMA warps from a side buffer), C6, 8 M samples" > /dev/null 2>&1
python tools/ncu_regions.py $O/$name.ncu-rep >> $O/${name}_ncu.txt 2>/dev/null
rm -f $O/$name.ncu-rep
grep -E "^generator|^mma|dmma.avg|pipe_fp64.avg|time_duration|bank_conflicts|local" $O/${name}_ncu.txt
