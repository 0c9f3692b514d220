#!/bin/bash
# round 2, 8-GPU call AC: the shipped build -- parity of the NCCL/IPC path (incl. the batched sharded kernel basis) + the n = 524288 bench line
set -u
G=${1:-8}
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02ac_$G.txt; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
stamp "dist_check parity"
timeout 600 $TR --master-port 29511 scripts/dist_check.py 4096 16384 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" | tail -25 | tee $O/dist_check_r02ac_$G.txt
stamp "bench n=524288"
timeout 900 $TR --master-port 29513 bench.py --gpus $G --steps 2 --warmup 1 2>>$O/bench_r02ac_$G.err | tail -1 | tee $O/bench_big_r02ac_$G.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('n=524288 ms/step', round(d['ms_per_step'],1), 'sweep share', round(r['sweep_share_of_step'],3), 'frac', round(r['frac'],3), 'e2e', d['e2e'].get('ms_per_step'), 'dist_parity', d.get('dist_parity',{}).get('equal'))" | tee -a $O/bench_r02ac_$G.txt
stamp done
grep -v "^W1017\|^\[W\|^$\|\*\*\*\|OMP_NUM" $O/bench_r02ac_$G.err | tail -5
