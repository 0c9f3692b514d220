#!/bin/bash
# round 2, GPU call I: the bench exactly as the driver runs it (both arms), N = 1
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02i.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/smi_r02i.txt; nproc >> $O/smi_r02i.txt
stamp "reference arm"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> $O/bench_ref_r02i.err | tee $O/bench_ref_r02i.json
stamp "b200 arm"
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 2> $O/bench_r02i.err | tee $O/bench_r02i.json
stamp done
tail -5 $O/bench_r02i.err
