#!/bin/bash
# Build compile-time variants of libgf2b200 into gf2bv_b200/variants/ (they travel to
# the GPU box with the snapshot) for an A/B with scripts/gpu_ab2.sh.
#   scripts/build_variants.sh name:"-DFLAG=.. -DFLAG2=.." [name:"flags" ...]
# Without arguments: the switches that are still in the source (the losers of the round-1 / round-2 A/Bs are gone).
set -eu
mkdir -p gf2bv_b200/variants
if [ $# -eq 0 ]; then
  set -- "trace:-DPERSIST_TRACE=1" "nolean:-DSWEEP_LEAN_UNITS=0" "nogj:-DPERSIST_GJ_SEARCH=0" "pad4:-DSWEEP_SEL_PAD=4" "pad8:-DSWEEP_SEL_PAD=8" \
         "notile:-DSWEEP_EARLY_TILE=0" "s128:-DGF2_STRIP_WORDS=16"
fi
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags \
       -I include -o gf2bv_b200/variants/libgf2b200_$name.so gf2bv_b200/csrc/gf2b200.cu -ldl &
done
wait
ls -la gf2bv_b200/variants/
