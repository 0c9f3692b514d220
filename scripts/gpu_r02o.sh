#!/bin/bash
set -u
G=${1:-8}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
port=29520
for pad in 24 6 40; do
  port=$((port+1))
  GF2B200_DIST_PAD=$pad timeout 600 $TR --master-port $port bench.py --gpus $G --steps 2 --warmup 1 --no-e2e --no-dist-parity 2>>$O/bench_r02o_$G.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('pad $pad n=524288 ms/step', round(d['ms_per_step'],1), 'sweep share', round(r['sweep_share_of_step'],3), 'frac', round(r['frac'],3))" | tee -a $O/bench_r02o_$G.txt
done
grep -v "^W1017\|^\[W\|^$\|\*\*\*\|OMP_NUM" $O/bench_r02o_$G.err | tail -5
