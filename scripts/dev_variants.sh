#!/bin/bash
# A/B of compile-time variants of libgf2b200 on ONE box: scripts/dev_variants.sh [n]
N=${1:-131072}
for so in gf2bv_b200/libgf2b200.so gf2bv_b200/variants/*.so gf2bv_b200/libgf2b200.so gf2bv_b200/variants/*.so; do
  echo -n "$(basename $so) "
  GF2B200_LIB=$PWD/$so python scripts/dev_bench.py $N 1 2 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],1), 'sweep', round(d['ms_sweep'],1), 'GB/s', round(d['sweep_GBs']), 'max', round(d['sweep_max_GBs']))"
done
