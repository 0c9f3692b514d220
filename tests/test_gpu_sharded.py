"""The row-sharded elimination on ONE GPU through loopback contexts:
`world` shards in one process, exchanges as device copies, otherwise the same
kernels and control flow as the NCCL path.  The answer must be bit-identical to
the oracle and to the unsharded solve (rank, pivot columns, particular solution):
the result is independent of which rows are chosen as pivots."""
import random

import numpy as np
import pytest

import oracle
from gf2bv_b200 import _shim
from test_gpu_solver import _rand_system

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def single():
    return _shim.Context(0)


@pytest.fixture(scope="module", params=[2, 3, 4, 8])
def sharded(request):
    return _shim.Context(0, shards=request.param)


def _check(got, want):
    assert got.status == want.status
    assert got.rank == want.rank
    if want.status == 0:
        assert np.array_equal(got.pivcols, want.pivcols)
        assert np.array_equal(got.origin, want.origin)


@pytest.mark.parametrize("m,n", [(1, 1), (5, 3), (3, 7), (64, 64), (130, 127), (300, 257), (1000, 513),
                                 (1500, 1024), (2100, 2050), (1025, 3000), (4099, 2500)])
def test_sharded_random_dense(sharded, single, m, n):
    rnd = random.Random(m * 7 + n)
    for consistent in (True, False):
        A, b = _rand_system(rnd, m, n, consistent=consistent)
        want = oracle.solve_packed(A, b, n, 0)
        _check(sharded.solve(A, b, n, 0), want)
        _check(single.solve(A, b, n, 0), want)


@pytest.mark.parametrize("m,n,cap", [(100, 100, 10), (300, 200, 64), (300, 200, 65), (700, 640, 300),
                                     (2000, 1500, 700), (64, 4096, 20), (4096, 64, 20), (1111, 999, 1)])
def test_sharded_rank_deficient(sharded, m, n, cap):
    rnd = random.Random(cap * 31 + m)
    for consistent in (True, False):
        A, b = _rand_system(rnd, m, n, rank_cap=cap, consistent=consistent)
        want = oracle.solve_packed(A, b, n, 0)
        got = sharded.solve(A, b, n, 0)
        _check(got, want)
        if consistent:
            assert got.status == 0 and oracle.residual(A, b, n, got.origin) == 0


def test_sharded_pivots_concentrated_in_one_shard(sharded):
    # all the rank lives in the LAST rows: other shards contribute no candidates
    rnd = random.Random(5)
    m, n = 1600, 400
    A, b = _rand_system(rnd, m, n, consistent=True)
    A[: m - 450] = 0
    b[:] = 0
    _check(sharded.solve(A, None, n, 0), oracle.solve_packed(A, None, n, 0))
    A2, b2 = _rand_system(rnd, m, n, consistent=True)
    A2[: m - 450] = 0  # zero rows keep their b bits: inconsistent unless those bits are 0
    _check(sharded.solve(A2, b2, n, 0), oracle.solve_packed(A2, b2, n, 0))


@pytest.mark.parametrize("m,n,cap", [
    (70, 64, None), (300, 200, 190), (700, 640, 620), (2200, 2110, None), (2300, 2100, 2070),
    # (large nullities: minutes on the emulated kernels, instant on the GPU)
    pytest.param(1300, 2100, 900, id="bignull-1300-2100-900"), pytest.param(64, 4096, 20, id="bignull-64-4096-20"),
    pytest.param(1111, 999, 1, id="bignull-1111-999-1")])
def test_sharded_kernel_basis(sharded, m, n, cap):
    """mode 1 on a sharded system: the blocked multi-right-hand-side triangular solve, every shard on
    its own echelon rows, one exchange per backward panel (basis_sharded, gf2b200.cu);
    basis values AND order (M4RI's sigma order) equal the oracle's, and a mode-0
    result taken afterwards is still the particular solution"""
    rnd = random.Random(cap or 0 + m)
    A, b = _rand_system(rnd, m, n, rank_cap=cap, consistent=True)
    want = oracle.solve_packed(A, b, n, 1)
    ss = sharded.system(m, n)
    ss.load_host(A, b)
    ss.eliminate()
    got = ss.result(1)
    _check(got, want)
    assert got.basis.shape == want.basis.shape and np.array_equal(got.basis, want.basis)
    _check(ss.result(0), want)
    got2 = ss.result(1)
    assert np.array_equal(got2.basis, want.basis) and np.array_equal(got2.origin, want.origin)


def test_sharded_kernel_basis_stats(sharded):
    """the sharded kernel basis is ONE blocked solve: a backward sweep per panel with pivots, not a
    back-substitution per free column"""
    rnd = random.Random(21)
    m, n, cap = 900, 1400, 500
    A, b = _rand_system(rnd, m, n, rank_cap=cap, consistent=True)
    ss = sharded.system(m, n)
    ss.load_host(A, b)
    ss.eliminate()
    got = ss.result(1)
    want = oracle.solve_packed(A, b, n, 1)
    _check(got, want)
    assert np.array_equal(got.basis, want.basis)
    st = ss.stats()
    assert 0 < st["basis_panels"] <= (n + 63) // 64 and st["basis_sweep_bytes"] > 0


def test_sharded_kernel_basis_homogeneous(sharded):
    A, _ = _rand_system(random.Random(9), 660, 700, consistent=True)
    want = oracle.solve_packed(A, None, 700, 1)
    got = sharded.solve(A, None, 700, 1)
    _check(got, want)
    assert np.array_equal(got.basis, want.basis)


@pytest.mark.parametrize("n,seed", [(4096, 1), (5000, 2), (8192, 3)])
def test_sharded_synthetic_equals_single(sharded, single, n, seed):
    ss = sharded.system(n, n)
    ss.generate(seed)
    ss.eliminate()
    got = ss.result(0)
    s1 = single.system(n, n)
    s1.generate(seed)
    s1.eliminate()
    ref = s1.result(0)
    _check(got, ref)
    assert got.status == 0 and ss.check_synthetic(seed, got.origin) == 0
    st = ss.stats()
    assert st["rank"] == ref.rank and st["exchange_bytes"] > 0 and st["m_local"] == n
    A, b, _ = oracle.synth(n, n, seed)
    assert np.array_equal(got.origin, oracle.solve_packed(A, b, n, 0).origin)
