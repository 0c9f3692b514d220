#!/bin/bash
# round 2, GPU call S: lean streaming units (no predicates, 32-bit shared addresses) -- A/B on one box
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02s.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv | tee $O/smi_r02s.txt
run() { # label lib n reps
  echo -n "$1 $3 " | tee -a $O/ab_r02s.txt
  GF2B200_LIB=$PWD/$2 timeout 120 python scripts/dev_bench.py $3 0 $4 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'max-panel ms', round(d['ms_sweep_max'],3), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6))" | tee -a $O/ab_r02s.txt
}
stamp parity
for v in lean1 lean1f lean2; do
  GF2B200_LIB=$PWD/gf2bv_b200/variants/libgf2b200_$v.so timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -x -q -k "random_dense or rank_deficient or synthetic_device or sparse or 32768_matches" 2>&1 | tail -2 | sed "s/^/$v: /" | tee -a $O/pytest_r02s.txt
done
stamp timing
for rep in 1 2; do
  for v in nolean lean1 lean1f lean2; do run $v gf2bv_b200/variants/libgf2b200_$v.so 131072 2; done
done
for v in nolean lean1 lean1f lean2; do run $v gf2bv_b200/variants/libgf2b200_$v.so 32768 4; done
for v in nolean lean1f; do run $v gf2bv_b200/variants/libgf2b200_$v.so 8192 4; done
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv | tee -a $O/smi_r02s.txt
stamp done
