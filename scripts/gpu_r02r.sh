#!/bin/bash
# round 2, GPU call R: sparse sweep (list units) -- parity, timing, MT19937 trace; ncu launch list of the
# bench command and one full capture of k_forward at n = 131072 (the shipped build)
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02r.txt; }
run() { # label lib mode n reps
  echo -n "$1 $4 " | tee -a $O/ab_r02r.txt
  GF2B200_LIB=$PWD/$2 GF2B200_FORWARD=$3 timeout 90 python scripts/dev_bench.py $4 0 $5 2>&1 | grep ms_total | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_total'],2), 'fwd', round(d['ms_forward'],2), 'max-panel ms', round(d['ms_sweep_max'],3), 'GB/s whole', round(d['sweep_bytes']/d['ms_forward']/1e6))" | tee -a $O/ab_r02r.txt
}
stamp parity
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest_r02r.txt
stamp timing
run persist gf2bv_b200/libgf2b200.so persist 131072 2
run persist gf2bv_b200/libgf2b200.so persist 32768 4
timeout 120 python scripts/dev_api.py 2>&1 | grep -E "m4ri_solve mode|device stats|pack only|LinearSystem" | tail -7 | tee $O/api_r02r.txt
GF2B200_LIB=$PWD/gf2bv_b200/variants/libgf2b200_trace.so GF2B200_TRACE_FILE=$O/trace.bin timeout 120 python scripts/dev_api.py > /dev/null 2>&1
echo "== trace MT19937 (20000 x 19968)" | tee -a $O/trace_r02r.txt
python scripts/trace_forward.py $O/trace.bin | tee -a $O/trace_r02r.txt
rm -f $O/trace.bin
stamp "ncu launch list of the bench command"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02r.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify --no-extra > $O/bench_under_ncu_r02r.log 2>&1
stamp "ncu full k_forward n=131072"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_forward -c 1 \
    -o $O/forward_r02r python scripts/dev_bench.py 131072 0 1 > $O/ncu_forward_r02r.log 2>&1
stamp done
