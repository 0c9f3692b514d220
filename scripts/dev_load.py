#!/usr/bin/env python
"""Developer probe: time gf2b200_system_load_host / gf2b200_solve phases."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from gf2bv_b200 import _shim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
nw = n // 64
tA = torch.empty((n, nw), dtype=torch.int64, pin_memory=True)
tb = torch.zeros((n // 64 + 1,), dtype=torch.int64, pin_memory=True)
A = tA.numpy().view(np.uint64); b = tb.numpy().view(np.uint64)
t0 = time.perf_counter(); _shim.synth_host(A, b, 0, n, 1); print("synth_host s", time.perf_counter() - t0)
ctx = _shim.Context(0)
t0 = time.perf_counter(); s = ctx.system(n, n); torch.cuda.synchronize(); print("system_create s", time.perf_counter() - t0)
for it in range(3):
    t0 = time.perf_counter(); s.load_host(A, b); t1 = time.perf_counter()
    s.eliminate(); t2 = time.perf_counter(); r = s.result(0); t3 = time.perf_counter()
    print(f"load {t1-t0:.4f} eliminate {t2-t1:.4f} result {t3-t2:.4f}")
t0 = time.perf_counter(); s.close(); print("destroy s", time.perf_counter() - t0)
for it in range(2):
    t0 = time.perf_counter(); r = ctx.solve(A, b, n, 0); print("solve s", time.perf_counter() - t0)
Ap = np.array(A)  # pageable copy
t0 = time.perf_counter(); r = ctx.solve(Ap, np.array(b), n, 0); print("solve pageable s", time.perf_counter() - t0)
