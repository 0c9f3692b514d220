#!/bin/bash
# round 2, GPU call W: programmatic dependent launch in the launch chain, size-based choice of the forward path
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02w.txt; }
run() { # label lib mode n reps
  echo -n "$1 $4 " | tee -a $O/ab_r02w.txt
  GF2B200_LIB=$PWD/$2 GF2B200_FORWARD=$3 timeout 120 python scripts/dev_bench.py $4 0 $5 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'max-panel ms', round(d['ms_sweep_max'],3), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6), 'one_kernel', d['forward_kernel_launches'])" | tee -a $O/ab_r02w.txt
}
stamp "parity, launch chain forced"
GF2B200_FORWARD=launches timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3 | sed 's/^/launches: /' | tee $O/pytest_r02w.txt
stamp "parity, default"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $O/pytest_r02w.txt
stamp timing
V=gf2bv_b200/variants
for rep in 1 2; do
  run auto gf2bv_b200/libgf2b200.so auto 131072 2
  run persist gf2bv_b200/libgf2b200.so persist 131072 2
  run nopdl $V/libgf2b200_prev.so launches 131072 2
done
for n in 65536 32768 8192; do
  run auto gf2bv_b200/libgf2b200.so auto $n 4
  run launches gf2bv_b200/libgf2b200.so launches $n 4
  run nopdl $V/libgf2b200_prev.so launches $n 4
done
stamp done
