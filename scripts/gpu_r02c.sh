#!/bin/bash
# round 2, GPU call C: persistent kernel variants (coefficient load path) + ncu of k_forward
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02c.txt; }
run() { # label lib mode n reps
  echo -n "$1 $4 " | tee -a $O/ab_r02c.txt
  GF2B200_LIB=$PWD/$2 GF2B200_FORWARD=$3 timeout 90 python scripts/dev_bench.py $4 0 $5 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'max-panel ms', round(d['ms_sweep_max'],3), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6))" | tee -a $O/ab_r02c.txt
}
stamp "parity (small)"
timeout 200 python -m pytest tests/test_gpu_solver.py -m gpu -x -q -k "random_dense or rank_deficient or forward_paths or 32768_matches" 2>&1 | tail -4 | tee $O/pytest_small_r02c.txt
stamp timing
for rep in 1 2; do
  run persist-cached gf2bv_b200/libgf2b200.so persist 131072 2
  run persist-ldcg gf2bv_b200/variants/libgf2b200_cf0.so persist 131072 2
  run persist-uncond gf2bv_b200/variants/libgf2b200_uncond.so persist 131072 2
  run launches gf2bv_b200/libgf2b200.so launches 131072 2
done
for n in 32768 8192; do
  run persist-cached gf2bv_b200/libgf2b200.so persist $n 4
  run persist-uncond gf2bv_b200/variants/libgf2b200_uncond.so persist $n 4
  run launches gf2bv_b200/libgf2b200.so launches $n 4
done
stamp "ncu k_forward n=65536"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_forward -c 1 \
    -o $O/forward_r02c python scripts/dev_bench.py 65536 0 1 > $O/ncu_forward_r02c.log 2>&1
stamp done
