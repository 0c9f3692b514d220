#!/bin/bash
# One GPU-box call on the DEFAULT build plus compile-time variants under
# gf2bv_b200/variants/: parity suite + smoke on the default, A/B of every variant at
# n=131072 and 32768, bench line, one full ncu capture of k_sweep, a partial launch list.
#   scripts/gpu_ab2.sh <tag> [launch-list count]
set -u
R=${1:-ab2}
NL=${2:-600}
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_$R.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/smi_$R.txt; nproc >> $O/smi_$R.txt
stamp "pytest -m gpu (default build)"
timeout 240 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} 2>&1 | tail -6 | tee $O/pytest_gpu_$R.txt
stamp "smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke_$R.txt
stamp "A/B n=131072"
LIBS=${AB_LIBS:-"gf2bv_b200/libgf2b200.so $(ls gf2bv_b200/variants/*.so)"}
for rep in $(seq 1 ${AB_REPS:-2}); do for so in $LIBS; do
  echo -n "$(basename $so) " | tee -a $O/ab_$R.txt
  GF2B200_LIB=$PWD/$so timeout 60 python scripts/dev_bench.py 131072 1 2 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],1), 'sweep', round(d['ms_sweep'],1), 'GB/s', round(d['sweep_GBs']), 'max', round(d['sweep_max_GBs']))" | tee -a $O/ab_$R.txt
done; done
stamp "A/B n=32768"
for so in $LIBS; do
  echo -n "32768 $(basename $so) " | tee -a $O/ab_$R.txt
  GF2B200_LIB=$PWD/$so timeout 40 python scripts/dev_bench.py 32768 0 4 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2))" | tee -a $O/ab_$R.txt
done
stamp "bench.py default"
timeout 150 python bench.py --steps 3 --warmup 3 2> $O/bench_$R.err | tee $O/bench_$R.json
stamp "ncu full sweep"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 20 -c 3 \
    -o $O/sweep_$R python scripts/dev_bench.py 131072 0 1 > $O/ncu_full_$R.log 2>&1
[ "$NL" -gt 0 ] && stamp "ncu launch list (first $NL launches)"
[ "$NL" -gt 0 ] && timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c $NL --csv \
    --log-file $O/launches_$R.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > $O/bench_under_ncu_$R.log 2>&1
stamp done
ls -la $O
