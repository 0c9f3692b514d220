#!/bin/bash
# round 2, GPU call B: the persistent forward kernel -- parity suite, then timing against the per-panel launch chain
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02b.txt; }
stamp "small parity first (a hang here must not cost the whole call)"
timeout 120 python -m pytest tests/test_gpu_solver.py -m gpu -x -q -k "random_dense or rank_deficient" 2>&1 | tail -4 | tee $O/pytest_small_r02b.txt
stamp "pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu_r02b.txt
stamp "timing"
for rep in 1 2; do for mode in persist launches; do
  echo -n "$mode 131072 " | tee -a $O/ab_r02b.txt
  GF2B200_FORWARD=$mode timeout 90 python scripts/dev_bench.py 131072 0 2 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],1), 'fwd', round(d['ms_forward'],1), 'sweep/kernel', round(d['ms_sweep'],1), 'max-panel ms', round(d['ms_sweep_max'],3), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6))" | tee -a $O/ab_r02b.txt
done; done
for mode in persist launches; do
  echo -n "$mode 32768 " | tee -a $O/ab_r02b.txt
  GF2B200_FORWARD=$mode timeout 60 python scripts/dev_bench.py 32768 0 4 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2))" | tee -a $O/ab_r02b.txt
  echo -n "$mode 8192 " | tee -a $O/ab_r02b.txt
  GF2B200_FORWARD=$mode timeout 60 python scripts/dev_bench.py 8192 0 4 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2))" | tee -a $O/ab_r02b.txt
done
stamp "api timing (MT19937)"
timeout 120 python scripts/dev_api.py 2>&1 | tail -5 | tee $O/api_r02b.txt
stamp done
