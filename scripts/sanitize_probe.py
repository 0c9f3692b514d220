#!/usr/bin/env python
"""Small solves for compute-sanitizer (memcheck / racecheck): single GPU, loopback
shards, kernel basis, rank-deficient and inconsistent systems; checked vs the oracle."""
import random, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import oracle
from gf2bv_b200 import _shim
from test_gpu_solver import _rand_system

rnd = random.Random(4)
def note(*a):
    print(*a, flush=True)

single, sh3 = _shim.Context(0), _shim.Context(0, shards=3)
for (m, n, cap) in [(70, 64, None), (300, 257, None), (1100, 1030, None), (2100, 2050, None), (700, 640, 300), (64, 1100, 20)]:
    for consistent in (True, False):
        A, b = _rand_system(rnd, m, n, rank_cap=cap, consistent=consistent)
        note('single + 3 shards', m, n, cap, consistent)
        want = oracle.solve_packed(A, b, n, 1)
        got = single.solve(A, b, n, 1)
        assert got.status == want.status and got.rank == want.rank
        if want.status == 0:
            assert np.array_equal(got.origin, want.origin) and np.array_equal(got.basis, want.basis)
        g3 = sh3.solve(A, b, n, 0)
        assert g3.status == want.status and g3.rank == want.rank
        if want.status == 0:
            assert np.array_equal(g3.origin, want.origin)
note('synthetic 3000')
s = single.system(3000, 3000); s.generate(1); s.eliminate(); r = s.result(0)
assert s.check_synthetic(1, r.origin) == 0
# round 2: the launch chain with k_sweep_apply, sparse panels (candidate lists + list sweep), rows loaded in
# blocks, the sharded kernel basis
import os
from test_gpu_solver import _near_triangular
os.environ["GF2B200_FORWARD"] = "launches"
chain = _shim.Context(0)
del os.environ["GF2B200_FORWARD"]
for (m, n, cap) in [(2100, 2050, None), (1500, 1400, 600)]:
    note('chain / shards mode 1 / blocks', m, n, cap)
    A, b = _rand_system(rnd, m, n, rank_cap=cap, consistent=True)
    want = oracle.solve_packed(A, b, n, 1)
    for c in (chain, sh3):
        got = c.solve(A, b, n, 1)
        assert got.rank == want.rank and np.array_equal(got.origin, want.origin) and np.array_equal(got.basis, want.basis)
    sb = single.system(m, n)
    sb.load_host_blocks(A, b, [(r0, min(256, m - r0)) for r0 in range(0, m, 256)][::-1])
    sb.eliminate()
    gb = sb.result(0)
    assert gb.rank == want.rank and np.array_equal(gb.origin, want.origin)
note('sparse near-triangular')
A, b = _near_triangular(0, 3000, 2500, 3)
want = oracle.solve_packed(A, b, 2500, 0)
for c in (single, chain):
    got = c.solve(A, b, 2500, 0)
    assert got.rank == want.rank and np.array_equal(got.origin, want.origin)
print("sanitize probe ok")
