#!/bin/bash
# round 2, multi-GPU call: the sharded look-ahead (publish + election inside the sweep) on real GPUs
set -u
G=${1:-2}
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02m_$G.txt; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
stamp "dist_check parity"
timeout 600 $TR --master-port 29511 scripts/dist_check.py 4096 5000 16384 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -25 | tee $O/dist_check_r02m_$G.txt
stamp "bench n=131072 look-ahead on/off"
for mode in on off; do
  if [ $mode = off ]; then export GF2B200_NO_DIST_LOOKAHEAD=1; else unset GF2B200_NO_DIST_LOOKAHEAD; fi
  echo "lookahead $mode" | tee -a $O/bench_r02m_$G.txt
  timeout 300 $TR --master-port 29512 bench.py --gpus $G --steps 3 --warmup 2 --size 131072 --no-e2e 2>>$O/bench_r02m_$G.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step', round(d['ms_per_step'],1), 'sweep share', round(r['sweep_share_of_step'],3), 'frac', round(r['frac'],3), 'dist_parity', d.get('dist_parity',{}).get('equal'))" | tee -a $O/bench_r02m_$G.txt
done
unset GF2B200_NO_DIST_LOOKAHEAD
stamp "bench n=524288"
timeout 900 $TR --master-port 29513 bench.py --gpus $G --steps 1 --warmup 1 2>>$O/bench_r02m_$G.err | tail -1 | tee $O/bench_big_r02m_$G.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('n=524288 ms/step', round(d['ms_per_step'],1), 'sweep share', round(r['sweep_share_of_step'],3), 'frac', round(r['frac'],3), 'e2e', d['e2e'].get('ms_per_step'))" | tee -a $O/bench_r02m_$G.txt
stamp done
tail -5 $O/bench_r02m_$G.err
