#!/usr/bin/env python
"""Developer timing probe: python scripts/dev_bench.py N [profile] [reps]"""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
from gf2bv_b200 import _shim
n = int(sys.argv[1]); prof = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = _shim.Context(0); ctx.set_profile(bool(prof))
s = ctx.system(n, n)
for it in range(reps):
    s.generate(1)
    t0 = time.time(); s.eliminate(); t1 = time.time()
    st = s.stats()
    W = n ** 3 / 3
    st["wall_s"] = t1 - t0
    st["bitops_per_s"] = W / (st["ms_total"] / 1e3)
    if st["ms_sweep"]:
        st["sweep_GBs"] = st["sweep_bytes"] / st["ms_sweep"] / 1e6
        st["sweep_max_GBs"] = st["sweep_bytes_max"] / st["ms_sweep_max"] / 1e6
    print(json.dumps(st))
r = s.result(0)
print("status", r.status, "rank", r.rank, "bad", s.check_synthetic(1, r.origin))
