#!/bin/bash
# round 2, GPU call D: full parity suite on the new build (persistent forward + batched kernel basis),
# per-CTA trace of k_forward, kernel-basis timing
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02d.txt; }
stamp "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -12 | tee $O/pytest_gpu_r02d.txt
stamp "trace"
for n in 32768 131072; do
  GF2B200_LIB=$PWD/gf2bv_b200/variants/libgf2b200_trace.so GF2B200_TRACE_FILE=$O/trace_$n.bin timeout 120 python scripts/dev_bench.py $n 0 2 2>&1 | grep ms_total | tail -1 | cut -c1-200
  python scripts/trace_forward.py $O/trace_$n.bin | tee $O/trace_$n.txt
  rm -f $O/trace_$n.bin
done
stamp "kernel basis timing"
timeout 300 python scripts/dev_basis.py 32768 4096 2>&1 | tail -3 | tee $O/basis_r02d.txt
timeout 300 python scripts/dev_basis.py 131072 1000 2>&1 | tail -3 | tee -a $O/basis_r02d.txt
stamp "timing default build"
for n in 131072 32768; do for mode in persist launches; do
  echo -n "$mode $n " | tee -a $O/ab_r02d.txt
  GF2B200_FORWARD=$mode timeout 90 python scripts/dev_bench.py $n 0 3 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'max-panel ms', round(d['ms_sweep_max'],3))" | tee -a $O/ab_r02d.txt
done; done
stamp done
