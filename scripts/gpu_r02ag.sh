#!/bin/bash
# round 2, GPU call AG: racecheck after the slow-path header fix (hdr_slow), twice
set -u
O=gpurun_out; mkdir -p $O
for i in 1 2; do
  T0=$(date +%s)
  timeout 200 compute-sanitizer --tool racecheck --log-file $O/sanitizer_racecheck_r02ag_$i.log python scripts/sanitize_probe.py 2>&1 | tail -2 | tee -a $O/sanitizer_racecheck_r02ag.out
  echo "racecheck run $i: $(( $(date +%s) - T0 )) s" | tee -a $O/sanitizer_racecheck_r02ag.out
  grep -E "RACECHECK SUMMARY" $O/sanitizer_racecheck_r02ag_$i.log | tee -a $O/sanitizer_racecheck_r02ag.out
done
