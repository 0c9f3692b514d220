#!/bin/bash
# round 2, GPU call V: panel geometry re-derived after the streaming loop (lean loop 275 -> 233 instructions) A/B
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02v.txt; }
run() { # label lib mode n reps
  echo -n "$1 $4 " | tee -a $O/ab_r02v.txt
  GF2B200_LIB=$PWD/$2 GF2B200_FORWARD=$3 timeout 120 python scripts/dev_bench.py $4 0 $5 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'max-panel ms', round(d['ms_sweep_max'],3), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6))" | tee -a $O/ab_r02v.txt
}
stamp parity
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_r02v.txt
stamp timing
V=gf2bv_b200/variants
for rep in 1 2; do
  run geo gf2bv_b200/libgf2b200.so persist 131072 2
  run prev $V/libgf2b200_prev.so persist 131072 2
  run pad8 $V/libgf2b200_pad8.so persist 131072 2
  run launches gf2bv_b200/libgf2b200.so launches 131072 2
done
for n in 32768 8192; do
  run geo gf2bv_b200/libgf2b200.so persist $n 4
  run prev $V/libgf2b200_prev.so persist $n 4
  run launches gf2bv_b200/libgf2b200.so launches $n 4
done
stamp trace
for n in 131072; do
  GF2B200_LIB=$PWD/$V/libgf2b200_trace.so GF2B200_TRACE_FILE=$O/trace.bin timeout 120 python scripts/dev_bench.py $n 0 1 > /dev/null 2>&1
  echo "== trace n=$n" | tee -a $O/trace_r02v.txt
  python scripts/trace_forward.py $O/trace.bin | tee -a $O/trace_r02v.txt
done
rm -f $O/trace.bin
stamp done
