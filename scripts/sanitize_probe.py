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
single, sh3 = _shim.Context(0), _shim.Context(0, shards=3)
for (m, n, cap) in [(70, 64, None), (300, 257, None), (1100, 1030, None), (2100, 2050, None), (700, 640, 300), (64, 1100, 20)]:
    for consistent in (True, False):
        A, b = _rand_system(rnd, m, n, rank_cap=cap, consistent=consistent)
        want = oracle.solve_packed(A, b, n, 1)
        got = single.solve(A, b, n, 1)
        assert got.status == want.status and got.rank == want.rank
        if want.status == 0:
            assert np.array_equal(got.origin, want.origin) and np.array_equal(got.basis, want.basis)
        g3 = sh3.solve(A, b, n, 0)
        assert g3.status == want.status and g3.rank == want.rank
        if want.status == 0:
            assert np.array_equal(g3.origin, want.origin)
s = single.system(3000, 3000); s.generate(1); s.eliminate(); r = s.result(0)
assert s.check_synthetic(1, r.origin) == 0
print("sanitize probe ok")
