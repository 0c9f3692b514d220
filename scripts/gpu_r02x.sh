#!/bin/bash
# round 2, multi-GPU call X: lean units in k_sweep_dist, the sharded batched kernel basis over NCCL
set -u
G=${1:-2}
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02x_$G.txt; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
stamp "dist_check parity"
timeout 600 $TR --master-port 29511 scripts/dist_check.py 4096 5000 16384 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -25 | tee $O/dist_check_r02x_$G.txt
stamp "pytest multiproc"
timeout 600 python -m pytest tests/test_gpu_multiproc.py -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_r02x_$G.txt
stamp "bench n=131072"
timeout 300 $TR --master-port 29512 bench.py --gpus $G --steps 3 --warmup 2 --size 131072 --no-e2e 2>>$O/bench_r02x_$G.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('n=131072 ms/step', round(d['ms_per_step'],1), 'sweep share', round(r['sweep_share_of_step'],3), 'frac', round(r['frac'],3), 'dist_parity', d.get('dist_parity',{}).get('equal'))" | tee -a $O/bench_r02x_$G.txt
stamp "bench n=524288"
timeout 900 $TR --master-port 29513 bench.py --gpus $G --steps 1 --warmup 1 2>>$O/bench_r02x_$G.err | tail -1 | tee $O/bench_big_r02x_$G.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('n=524288 ms/step', round(d['ms_per_step'],1), 'sweep share', round(r['sweep_share_of_step'],3), 'frac', round(r['frac'],3), 'e2e', d['e2e'].get('ms_per_step'))" | tee -a $O/bench_r02x_$G.txt
stamp done
tail -5 $O/bench_r02x_$G.err
