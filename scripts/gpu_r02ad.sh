#!/bin/bash
# round 2, GPU call AD: k_sweep_apply (the launch chain's sweep applies the next panel in its tail) -- parity + A/B
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02ad.txt; }
run() { # label env mode n reps
  echo -n "$1 $4 " | tee -a $O/ab_r02ad.txt
  env $2 GF2B200_FORWARD=$3 timeout 120 python scripts/dev_bench.py $4 0 $5 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6), 'launches', d['kernel_launches'])" | tee -a $O/ab_r02ad.txt
}
stamp "parity, launch chain forced (k_sweep_apply)"
GF2B200_FORWARD=launches timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3 | sed 's/^/launches: /' | tee $O/pytest_r02ad.txt
stamp "parity, default"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $O/pytest_r02ad.txt
stamp timing
for rep in 1 2; do
  run tail-apply X=1 auto 131072 2
  run no-tail-apply GF2B200_NO_TAIL_APPLY=1 auto 131072 2
done
for n in 65536 32768 8192; do
  run tail-apply X=1 launches $n 4
  run no-tail-apply GF2B200_NO_TAIL_APPLY=1 launches $n 4
  run k_forward X=1 persist $n 4
done
stamp done
