#!/bin/bash
# round 2, GPU call AF: racecheck again, with the k_forward watchdog at 600 s (the tool slows the kernels ~100x) and progress notes
set -u
O=gpurun_out; mkdir -p $O
GF2B200_LIB=$PWD/gf2bv_b200/variants/libgf2b200_longwait.so timeout 330 compute-sanitizer --tool racecheck --log-file $O/sanitizer_racecheck_r02af.log python scripts/sanitize_probe.py 2>&1 | tail -12 | tee $O/sanitizer_racecheck_r02af.out
grep -E "RACECHECK SUMMARY|Race reported" $O/sanitizer_racecheck_r02af.log | sort | uniq -c | head -5
