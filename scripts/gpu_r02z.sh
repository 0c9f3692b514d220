#!/bin/bash
# round 2, GPU call Z: same-box table (lean units / Gauss-Jordan search on and off, both forward paths) + ncu --set full of k_sweep
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02z.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/smi_r02z.txt
run() { # label lib mode n reps
  echo -n "$1 $4 " | tee -a $O/ab_r02z.txt
  GF2B200_LIB=$PWD/$2 GF2B200_FORWARD=$3 timeout 120 python scripts/dev_bench.py $4 0 $5 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6), 'one_kernel', d['forward_kernel_launches'])" | tee -a $O/ab_r02z.txt
}
B=gf2bv_b200/variants/libgf2b200_base.so; L=gf2bv_b200/libgf2b200.so
stamp table
for n in 131072 65536 32768 8192; do
  reps=4; [ $n = 131072 ] && reps=2
  run "chain-round1-loop" $B launches $n $reps
  run "chain-lean" $L launches $n $reps
  run "k_forward-first" $B persist $n $reps
  run "k_forward-lean+gj" $L persist $n $reps
done
run "chain-round1-loop" $B launches 131072 2
run "chain-lean" $L launches 131072 2
stamp "ncu full k_sweep n=131072 (3 launches from panel 20)"
timeout 600 ncu --set full --clock-control none --import-source on -k k_sweep -s 20 -c 3 \
    -o $O/sweep_r02z python scripts/dev_bench.py 131072 0 1 > $O/ncu_sweep_r02z.log 2>&1
tail -3 $O/ncu_sweep_r02z.log
stamp done
