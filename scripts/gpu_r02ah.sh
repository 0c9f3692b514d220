#!/bin/bash
# round 2, GPU call AH: ncu evidence of the final build's headline kernel (k_sweep_apply): launch list + one full capture
set -u
O=gpurun_out; mkdir -p $O
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_r02ah.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify --no-extra > $O/bench_under_ncu_r02ah.log 2>&1
echo "launch list rc $?"
timeout 150 ncu --set full --clock-control none --import-source on -k k_sweep_apply -s 20 -c 3 \
    -o $O/sweep_r02ah python scripts/dev_bench.py 131072 0 1 > $O/ncu_sweep_r02ah.log 2>&1
echo "full capture rc $?"; ls -la $O/sweep_r02ah.ncu-rep
