#!/usr/bin/env python
"""Developer probe: the API path (m4ri_solve through the extension) on the MT19937
fixture (BASELINE config 2): pack / solve breakdown, CPU oracle beside it."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import oracle
from conftest import load_sparse_eqs, GOLDEN
from gf2bv_b200 import _internal, _shim
import gf2bv_b200 as gf2bv

eqs, cols, digest, state = load_sparse_eqs(GOLDEN / "mt19937_bs32.npz")
print("system", len(eqs), "x", cols)
t0 = time.perf_counter(); a, b = _internal._pack_probe(eqs, cols); t_pack = time.perf_counter() - t0
print(f"pack only (host, digit-wise): {t_pack*1e3:.1f} ms for {len(a)/1e6:.1f} MB")
for it in range(4):
    t0 = time.perf_counter(); sol = _internal.m4ri_solve(eqs, cols, 0); dt = time.perf_counter() - t0
    print(f"m4ri_solve mode 0: {dt*1e3:.1f} ms")
t0 = time.perf_counter(); sp = _internal.m4ri_solve(eqs, cols, 1); dt = time.perf_counter() - t0
print(f"m4ri_solve mode 1: {dt*1e3:.1f} ms  dim {sp.dimension}")
lin = gf2bv.LinearSystem([32] * 624)
t0 = time.perf_counter(); s1 = lin.solve_one(eqs); dt = time.perf_counter() - t0
print(f"LinearSystem.solve_one: {dt*1e3:.1f} ms  ok={s1 == state}")
A = np.frombuffer(a, dtype=np.uint64).reshape(len(eqs), -1); B = np.frombuffer(b, dtype=np.uint64)
ctx = _shim.Context(0)
for it in range(3):
    t0 = time.perf_counter(); r = ctx.solve(A, B, cols, 0); dt = time.perf_counter() - t0
    print(f"gf2b200_solve on packed (pageable numpy) buffers: {dt*1e3:.1f} ms")
s = ctx.system(len(eqs), cols); s.load_host(A, B); s.eliminate(); print("device stats", {k: round(v, 2) for k, v in s.stats().items() if k.startswith("ms_")})
t0 = time.perf_counter(); want = oracle.solve_packed(A, B, cols, 0); dt = time.perf_counter() - t0
print(f"CPU oracle (Four-Russians port, {oracle.threads()} threads) on the packed system: {dt*1e3:.1f} ms; equal={np.array_equal(want.origin, r.origin)}")
t0 = time.perf_counter(); want = oracle.solve_packed(A, B, cols, 0, tier='schoolbook'); dt = time.perf_counter() - t0
print(f"CPU oracle schoolbook: {dt*1e3:.1f} ms")
