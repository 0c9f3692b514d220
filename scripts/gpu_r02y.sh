#!/bin/bash
# round 2, GPU call Y: the bench exactly as the driver runs it (both arms, N = 1) + the ncu evidence of the same build
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02y.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/smi_r02y.txt; nproc >> $O/smi_r02y.txt
run() { # label lib mode n reps
  echo -n "$1 $4 " | tee -a $O/ab_r02y.txt
  GF2B200_LIB=$PWD/$2 GF2B200_FORWARD=$3 timeout 120 python scripts/dev_bench.py $4 0 $5 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6), 'one_kernel', d['forward_kernel_launches'])" | tee -a $O/ab_r02y.txt
}
stamp "A/B tail select"
for rep in 1 2; do
  run auto gf2bv_b200/libgf2b200.so auto 131072 2
  run tailsel gf2bv_b200/variants/libgf2b200_tailsel.so auto 131072 2
done
stamp "reference arm"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> $O/bench_ref_r02y.err | tee $O/bench_ref_r02y.json | cut -c1-300
stamp "b200 arm"
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 2> $O/bench_r02y.err | tee $O/bench_r02y.json | cut -c1-400
stamp "ncu launch list (first 900 launches of the bench command)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_r02y.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify --no-extra > $O/bench_under_ncu_r02y.log 2>&1
stamp "ncu full k_sweep n=131072 (3 launches from panel 20)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweepE -s 20 -c 3 \
    -o $O/sweep_r02y python scripts/dev_bench.py 131072 0 1 > $O/ncu_sweep_r02y.log 2>&1
stamp "ncu full k_forward n=32768"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_forward -c 1 \
    -o $O/forward_r02y python scripts/dev_bench.py 32768 0 1 > $O/ncu_forward_r02y.log 2>&1
stamp done
tail -3 $O/bench_r02y.err
