#!/bin/bash
# dev tool (run under gpurun): ncu --set full capture of the fused Gram kernel for C6, 2M samples.  usage: tools/prof_gram.sh <out-name> [lib] [dbg]
out=$1; lib=$2; dbg=${3:-0}
RDB_GRAM_DEBUG=$dbg ncu --set full --clock-control none --import-source on -k regex:gram_fused_kernel -s 3 -c 1 -f -o gpurun_out/$out \
  python tools/bench_gram.py 2000000 1 ${lib:+--lib $lib} > gpurun_out/$out.log 2>&1
tail -2 gpurun_out/$out.log
