#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
run() { # label lib mode n reps
  echo -n "$1 $4 " | tee -a $O/ab_r02k.txt
  GF2B200_LIB=$PWD/$2 GF2B200_FORWARD=$3 timeout 90 python scripts/dev_bench.py $4 0 $5 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'max-panel ms', round(d['ms_sweep_max'],3), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6))" | tee -a $O/ab_r02k.txt
}
for rep in 1 2; do
  run pad6 gf2bv_b200/libgf2b200.so persist 131072 2
  for p in 10 12 14 16; do run pad$p gf2bv_b200/variants/libgf2b200_pad$p.so persist 131072 2; done
done
for p in 10 12 14 16; do run pad$p gf2bv_b200/variants/libgf2b200_pad$p.so persist 32768 4; done
run pad6 gf2bv_b200/libgf2b200.so persist 32768 4
