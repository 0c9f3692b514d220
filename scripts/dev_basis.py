#!/usr/bin/env python
"""Developer timing probe of the batched kernel basis: python scripts/dev_basis.py n nullity"""
import sys, time, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from gf2bv_b200 import _shim
from test_gpu_solver import _dup_rows_system
n, d = int(sys.argv[1]), int(sys.argv[2])
A, b = _dup_rows_system(n, n - d, 11)
ctx = _shim.Context(0)
s = ctx.system(n, n)
for it in range(2):
    s.load_host(A, b); s.eliminate()
    t0 = time.perf_counter(); r = s.result(1); dt = time.perf_counter() - t0
    st = s.stats()
    print(json.dumps({"n": n, "nullity": int(r.basis.shape[0]), "rank": int(r.rank), "result_wall_ms": dt * 1e3,
                      "ms_basis_solve": st["ms_basis_solve"], "ms_basis_output": st["ms_basis_output"],
                      "basis_sweep_bytes": st["basis_sweep_bytes"], "basis_panels": st["basis_panels"],
                      "sweep_GBs": st["basis_sweep_bytes"] / st["ms_basis_solve"] / 1e6 if st["ms_basis_solve"] else None,
                      "ms_eliminate": st["ms_total"]}))
# spot check: A v = 0 for a few basis vectors
from oracle import residual
bad = sum(residual(A, None, n, r.basis[i].copy()) for i in (0, len(r.basis) // 2, len(r.basis) - 1))
print("basis residual rows", bad)
