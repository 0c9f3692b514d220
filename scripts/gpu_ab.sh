#!/bin/bash
# One GPU-box call: parity suite + A/B + bench + ncu for a compile-time variant of
# libgf2b200 (built HERE beforehand, it travels with the snapshot).
#   scripts/gpu_ab.sh <variant .so> <tag>
set -u
V=$PWD/$1
R=${2:-ab}
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_$R.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/smi_$R.txt; nproc >> $O/smi_$R.txt
stamp "pytest -m gpu on $1"
GF2B200_LIB=$V timeout 220 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu_$R.txt
stamp "A/B n=131072"
for so in gf2bv_b200/libgf2b200.so $1 gf2bv_b200/libgf2b200.so $1; do
  echo -n "$(basename $so) " | tee -a $O/ab_$R.txt
  GF2B200_LIB=$PWD/$so timeout 60 python scripts/dev_bench.py 131072 1 2 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],1), 'sweep', round(d['ms_sweep'],1), 'GB/s', round(d['sweep_GBs']), 'max', round(d['sweep_max_GBs']))" | tee -a $O/ab_$R.txt
done
stamp "bench.py variant"
GF2B200_LIB=$V timeout 150 python bench.py --steps 3 --warmup 3 2> $O/bench_$R.err | tee $O/bench_$R.json
stamp "ncu full sweep, variant"
GF2B200_LIB=$V timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 20 -c 3 \
    -o $O/sweep_$R python scripts/dev_bench.py 131072 0 1 > $O/ncu_full_$R.log 2>&1
stamp "ncu launch list, variant"
GF2B200_LIB=$V timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 6290 --csv \
    --log-file $O/launches_$R.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > $O/bench_under_ncu_$R.log 2>&1
stamp "bench.py default"
timeout 150 python bench.py --steps 3 --warmup 3 2> $O/bench_default_$R.err | tee $O/bench_default_$R.json
stamp "A/B n=32768"
for so in gf2bv_b200/libgf2b200.so $1; do
  echo -n "32768 $(basename $so) " | tee -a $O/ab_$R.txt
  GF2B200_LIB=$PWD/$so timeout 40 python scripts/dev_bench.py 32768 0 3 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2))" | tee -a $O/ab_$R.txt
done
stamp "bench.py --impl reference"
timeout 90 python bench.py --impl reference --steps 2 --warmup 1 2>> $O/bench_$R.err | tee $O/bench_reference_$R.json
stamp done
ls -la $O
