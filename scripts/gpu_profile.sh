#!/bin/bash
# Run on the GPU box (gpurun): GPU parity tests, bench line, ncu launch list and
# one full ncu capture of the sweep kernel.  Outputs land in gpurun_out/.
set -u
R=${1:-r01}
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.build()' 2>&1 | tail -3
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$R.txt
nproc >> gpurun_out/smi_$R.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$R.txt
timeout 900 python bench.py --steps 3 --warmup 3 2> gpurun_out/bench_$R.err | tee gpurun_out/bench_$R.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>> gpurun_out/bench_$R.err | tee gpurun_out/bench_reference_$R.json
# launch list of one bench step: the first solve of `bench.py --steps 1 --warmup 0`
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6290 --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu \
    > gpurun_out/bench_under_ncu_$R.log 2>&1
# full capture of three early (largest) sweeps
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 20 -c 3 \
    -o gpurun_out/sweep_$R python scripts/dev_bench.py 131072 0 1 > gpurun_out/ncu_full_$R.log 2>&1
# one-GPU run of the multi-GPU size (denominator of the 8-GPU speed-up)
if [ "${2:-}" = "big" ]; then
  timeout 600 python bench.py --size 524288 --steps 1 --warmup 1 --no-e2e --no-cpu 2>> gpurun_out/bench_$R.err \
      | tee gpurun_out/bench_1gpu_524288_$R.json
fi
ls -la gpurun_out
