#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
run() { # label lib mode n reps
  echo -n "$1 $4 " | tee -a $O/ab_r02p.txt
  GF2B200_LIB=$PWD/$2 GF2B200_FORWARD=$3 timeout 90 python scripts/dev_bench.py $4 0 $5 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'max-panel ms', round(d['ms_sweep_max'],3), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6))" | tee -a $O/ab_r02p.txt
}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest_r02p.txt
for rep in 1 2; do
  run persist gf2bv_b200/libgf2b200.so persist 131072 2
  run launches gf2bv_b200/libgf2b200.so launches 131072 2
done
for n in 32768 8192; do
  run persist gf2bv_b200/libgf2b200.so persist $n 4
  run launches gf2bv_b200/libgf2b200.so launches $n 4
done
for n in 32768; do
  GF2B200_LIB=$PWD/gf2bv_b200/variants/libgf2b200_trace.so GF2B200_TRACE_FILE=$O/trace.bin timeout 120 python scripts/dev_bench.py $n 0 2 > /dev/null 2>&1
  echo "== trace n=$n" | tee -a $O/trace_r02p.txt
  python scripts/trace_forward.py $O/trace.bin | tee -a $O/trace_r02p.txt
  rm -f $O/trace.bin
done
timeout 120 python scripts/dev_api.py 2>&1 | grep -E "m4ri_solve mode 0|device stats|pack only|LinearSystem" | tail -6 | tee $O/api_r02p.txt
GF2B200_LIB=$PWD/gf2bv_b200/variants/libgf2b200_trace.so GF2B200_TRACE_FILE=$O/trace.bin timeout 120 python scripts/dev_api.py > /dev/null 2>&1
echo "== trace MT19937 (20000 x 19968)" | tee -a $O/trace_r02p.txt
python scripts/trace_forward.py $O/trace.bin | tee -a $O/trace_r02p.txt
rm -f $O/trace.bin
