#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
for n in 32768; do
  GF2B200_LIB=$PWD/gf2bv_b200/variants/libgf2b200_trace.so GF2B200_TRACE_FILE=$O/trace.bin timeout 120 python scripts/dev_bench.py $n 0 2 > /dev/null 2>&1
  echo "== trace n=$n" | tee -a $O/trace_r02g.txt
  python scripts/trace_forward.py $O/trace.bin | tee -a $O/trace_r02g.txt
  rm -f $O/trace.bin
done
