#!/bin/bash
# Hazard check of the kernels WITHOUT a GPU: the GPU parity suite against the
# --racecheck build of the CPU-emulated kernels (tests/cpu_emu/emu_racecheck.cpp: every
# shared/device memory access of every GPU thread is checked against the GPU's own
# synchronisation -- __syncthreads, warp collectives, mbarrier waits, atomics, kernel
# boundaries).  The GPU-side counterpart is scripts/sanitize.sh (compute-sanitizer).
# What it covers: the launch chain (k_select / k_apply / k_sweep, forced here with
# GF2B200_FORWARD=launches GF2B200_NO_TAIL_APPLY=1), the loopback-sharded kernels, back-substitution and
# kernel basis.  NOT covered: k_forward and k_sweep_apply -- their CTAs hand data over through
# release / acquire flags inside one launch, which the checker does not model (it reports every
# such hand-over as a hazard); those protocols are argued in DESIGN.md 3a and exercised on the GPU.
#   scripts/emu_racecheck.sh [strip words: 8|16] [extra pytest args]
# Prints "EMU-RACECHECK: N hazard(s)"; exit status 1 if N > 0.
set -u
SWORDS=${1:-8}; shift || true
LIB=$(python tests/cpu_emu/build_emu.py --strip-words $SWORDS --racecheck | tail -1)
LOG=$(mktemp)
GF2B200_FORWARD=launches GF2B200_NO_TAIL_APPLY=1 GF2B200_TEST_EMULATION=1 GF2B200_LIB=$LIB python -m pytest tests/test_gpu_solver.py tests/test_gpu_sharded.py tests/test_gpu_api.py -m gpu -x -q -p no:cacheprovider \
  -k "not 32768 and not mt19937 and not 8192 and not 1025-3000 and not 2000-1500 and not 4099 and not 5000 and not 4096 and not 0.001 and not 2100 and not bignull" "$@" 2>&1 | tee $LOG | grep -v "^EMU-RACECHECK hazard" | tail -5
grep "^EMU-RACECHECK hazard" $LOG | head -20
N=$(grep -c "^EMU-RACECHECK hazard" $LOG); rm -f $LOG
[ "$N" -eq 0 ]
