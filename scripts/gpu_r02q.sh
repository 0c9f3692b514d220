#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
run() { # label lib mode n reps
  echo -n "$1 $4 " | tee -a $O/ab_r02q.txt
  GF2B200_LIB=$PWD/$2 GF2B200_FORWARD=$3 timeout 90 python scripts/dev_bench.py $4 0 $5 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'max-panel ms', round(d['ms_sweep_max'],3), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6))" | tee -a $O/ab_r02q.txt
}
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest_r02q.txt
run persist gf2bv_b200/libgf2b200.so persist 131072 2
run persist gf2bv_b200/libgf2b200.so persist 32768 4
timeout 120 python scripts/dev_api.py 2>&1 | grep -E "m4ri_solve mode 0|device stats|pack only|LinearSystem" | tail -6 | tee $O/api_r02q.txt
GF2B200_FORWARD=launches timeout 120 python scripts/dev_api.py 2>&1 | grep -E "m4ri_solve mode 0|device stats" | tail -2 | sed 's/^/launches: /' | tee -a $O/api_r02q.txt
GF2B200_LIB=$PWD/gf2bv_b200/variants/libgf2b200_trace.so GF2B200_TRACE_FILE=$O/trace.bin timeout 120 python scripts/dev_api.py > /dev/null 2>&1
echo "== trace MT19937 (20000 x 19968)" | tee -a $O/trace_r02q.txt
python scripts/trace_forward.py $O/trace.bin | tee -a $O/trace_r02q.txt
rm -f $O/trace.bin
