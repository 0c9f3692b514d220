#!/bin/bash
# round 2, GPU call AE: the final build -- racecheck after the syncwarp fix, smoke(), the whole GPU suite, both bench arms as the driver runs them
set -u
O=gpurun_out; mkdir -p $O
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline_r02ae.txt; }
stamp racecheck
timeout 300 compute-sanitizer --tool racecheck --log-file $O/sanitizer_racecheck_r02ae.log python scripts/sanitize_probe.py 2>&1 | tail -1
grep -E "RACECHECK SUMMARY|Race reported" $O/sanitizer_racecheck_r02ae.log | sort | uniq -c | head -5
stamp smoke
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke_r02ae.txt
stamp pytest
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_r02ae.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/smi_r02ae.txt; nproc >> $O/smi_r02ae.txt
stamp "reference arm"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> $O/bench_ref_r02ae.err | tee $O/bench_ref_r02ae.json | cut -c1-200
stamp "b200 arm"
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 2> $O/bench_r02ae.err | tee $O/bench_r02ae.json | cut -c1-300
stamp done
tail -3 $O/bench_r02ae.err
