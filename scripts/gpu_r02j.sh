#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
GF2B200_LIB=$PWD/gf2bv_b200/variants/libgf2b200_trace.so GF2B200_TRACE_FILE=$O/trace_131072.bin timeout 120 python scripts/dev_bench.py 131072 0 2 > /dev/null 2>&1
python - <<'PY'
import numpy as np
raw=np.fromfile("gpurun_out/trace_131072.bin",dtype=np.uint64); nw,G=int(raw[0]),int(raw[1])
tr=raw[2+nw+2:].astype(np.int64).reshape(nw,G,8)
# keep panels 0..767 only, as float32 relative to panel start, slots 2 (units done) and 4 (apply done) and 5
sl=tr[:768]
t0=sl[:,:,0].min(axis=1)
out=np.stack([(sl[:,:,k]-t0[:,None]).astype(np.float32) for k in (0,1,2,3,4,5)],axis=2)
np.save("gpurun_out/trace_131072_head.npy", out)
PY
rm -f $O/trace_131072.bin
