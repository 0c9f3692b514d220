#!/bin/bash
# round 2, GPU call AA: look-ahead search at the start of the panel (before the tables and unit 0) -- parity, A/B, trace
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02aa.txt; }
run() { # label lib mode n reps
  echo -n "$1 $4 " | tee -a $O/ab_r02aa.txt
  GF2B200_LIB=$PWD/$2 GF2B200_FORWARD=$3 timeout 120 python scripts/dev_bench.py $4 0 $5 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6), 'one_kernel', d['forward_kernel_launches'])" | tee -a $O/ab_r02aa.txt
}
stamp parity
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_r02aa.txt
GF2B200_FORWARD=persist timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -x -q -k "synthetic or forward_paths" 2>&1 | tail -2 | sed 's/^/persist forced: /' | tee -a $O/pytest_r02aa.txt
stamp timing
V=gf2bv_b200/variants
for n in 32768 16384 8192 65536; do
  run early gf2bv_b200/libgf2b200.so persist $n 4
  run prev $V/libgf2b200_prev.so persist $n 4
done
run early gf2bv_b200/libgf2b200.so persist 131072 2
run prev $V/libgf2b200_prev.so persist 131072 2
stamp trace
GF2B200_LIB=$PWD/$V/libgf2b200_trace.so GF2B200_TRACE_FILE=$O/trace.bin timeout 120 python scripts/dev_bench.py 32768 0 1 > /dev/null 2>&1
echo "== trace n=32768" | tee -a $O/trace_r02aa.txt
python scripts/trace_forward.py $O/trace.bin | tee -a $O/trace_r02aa.txt
rm -f $O/trace.bin
stamp api
timeout 120 python scripts/dev_api.py 2>&1 | grep -E "m4ri_solve mode|device stats|pack only|LinearSystem" | tail -8 | tee $O/api_r02aa.txt
stamp done
