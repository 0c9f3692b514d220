#!/bin/bash
# round 2, GPU call A: l1tex probe under ncu + A/B of the committed-but-untimed switches
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/smi_r02a.txt; nproc >> $O/smi_r02a.txt
timeout 300 ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__lsu_writeback_active.sum,l1tex__lsuin_requests.sum \
  --clock-control none --csv --log-file $O/l1tex_probe_r02a.csv build/l1tex_probe > $O/l1tex_probe_r02a.log 2>&1
for rep in 1 2; do for so in gf2bv_b200/libgf2b200.so $(ls gf2bv_b200/variants/*.so); do
  echo -n "$(basename $so) " | tee -a $O/ab_r02a.txt
  GF2B200_LIB=$PWD/$so timeout 60 python scripts/dev_bench.py 131072 1 2 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],1), 'sweep', round(d['ms_sweep'],1), 'GB/s', round(d['sweep_GBs']), 'max', round(d['sweep_max_GBs']))" | tee -a $O/ab_r02a.txt
done; done
echo -n "carveout72 " | tee -a $O/ab_r02a.txt
GF2B200_CARVEOUT=72 timeout 60 python scripts/dev_bench.py 131072 0 2 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],1))" | tee -a $O/ab_r02a.txt
for so in gf2bv_b200/libgf2b200.so $(ls gf2bv_b200/variants/*.so); do
  echo -n "32768 $(basename $so) " | tee -a $O/ab_r02a.txt
  GF2B200_LIB=$PWD/$so timeout 40 python scripts/dev_bench.py 32768 0 4 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2))" | tee -a $O/ab_r02a.txt
done
echo -n "32768 carveout72 " | tee -a $O/ab_r02a.txt
GF2B200_CARVEOUT=72 timeout 40 python scripts/dev_bench.py 32768 0 4 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2))" | tee -a $O/ab_r02a.txt
