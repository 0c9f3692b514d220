#!/bin/bash
# Memory check of the kernels WITHOUT a GPU: the GPU parity suite against the
# AddressSanitizer build of the CPU-emulated kernels (tests/cpu_emu).  Device
# allocations are heap blocks and shared memory is static storage there, so an
# out-of-bounds access of a kernel is reported like any heap/global overflow.
# The GPU-side counterpart is scripts/sanitize.sh (compute-sanitizer memcheck/racecheck).
# (UBSan: `python tests/cpu_emu/build_emu.py --ubsan`, LD_PRELOAD=$(gcc -print-file-name=libubsan.so);
#  parity subset and tests/cpu_emu/fuzz_emu.py are clean under both.)
#   scripts/emu_asan.sh [strip words: 8|16] [extra pytest args]
set -eu
SWORDS=${1:-8}; shift || true
LIB=$(python tests/cpu_emu/build_emu.py --strip-words $SWORDS --asan)
LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 GF2B200_TEST_EMULATION=1 GF2B200_LIB=$LIB \
python -m pytest tests/test_gpu_solver.py tests/test_gpu_sharded.py tests/test_gpu_api.py -m gpu -x -q -p no:cacheprovider \
  -k "not 32768 and not mt19937 and not 8192 and not 1025-3000 and not 2000-1500 and not 4099 and not 5000 and not 4096 and not 0.001 and not 2100 and not bignull" "$@"
