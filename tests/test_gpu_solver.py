"""Parity of the CUDA path (through the C-ABI, libgf2b200.so) against the CPU oracle.

Everything here needs a B200: run with ``pytest -m gpu``.  Bit-exact is the bar
(integer/bit work): rank, pivot columns, particular solution and kernel basis
(values AND order) must equal the oracle's.
"""
import hashlib
import random

import numpy as np
import pytest

import oracle
from gf2bv_b200 import _shim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _shim.Context(0)


def _rand_system(rnd, m, n, rank_cap=None, consistent=None, density=0.5):
    """Random m x n system as packed words; rank_cap limits the row space."""
    nw = (n + 63) // 64
    rng = np.random.default_rng(rnd.getrandbits(32))
    if rank_cap is None:
        bits = rng.random((m, n)) < density
    else:
        basis = rng.random((rank_cap, n)) < 0.5
        comb = rng.random((m, rank_cap)) < 0.5
        bits = (comb.astype(np.uint8) @ basis.astype(np.uint8)) % 2 == 1
    pad = np.zeros((m, nw * 64), dtype=bool)
    pad[:, :n] = bits
    A = np.packbits(pad, axis=1, bitorder="little").view(np.uint64).reshape(m, nw).copy()
    if consistent:
        x = rng.random(n) < 0.5
        bv = (bits.astype(np.int64) @ x.astype(np.int64)) % 2 == 1
    else:
        bv = rng.random(m) < 0.5
    bp = np.zeros(((m + 63) // 64) * 64, dtype=bool)
    bp[:m] = bv
    b = np.packbits(bp, bitorder="little").view(np.uint64).copy()
    return A, b


def _assert_same(got, want, mode):
    assert got.status == want.status
    assert got.rank == want.rank
    if want.status == 1:
        return
    assert np.array_equal(got.pivcols, want.pivcols)
    assert np.array_equal(got.origin, want.origin)
    if mode == 1:
        assert got.basis.shape == want.basis.shape
        assert np.array_equal(got.basis, want.basis)


# includes the layout's edges: n = 960 / 1024 (b word last of a strip / first of the next),
# n = 1984 / 2048 (b word inside / just past a 32-word super-panel of the back-substitution)
SHAPES = [(1, 1), (4, 4), (5, 3), (3, 7), (64, 64), (65, 64), (64, 65), (130, 127), (200, 128),
          (128, 200), (300, 257), (1000, 513), (965, 960), (1500, 1024), (1990, 1984), (2050, 2048),
          (2100, 2050), (1025, 3000)]


@pytest.mark.parametrize("m,n", SHAPES)
def test_random_dense_matches_oracle(ctx, m, n):
    rnd = random.Random(m * 100003 + n)
    for consistent in (True, False):
        A, b = _rand_system(rnd, m, n, consistent=consistent)
        for mode in (0, 1):
            want = oracle.solve_packed(A, b, n, mode, tier="schoolbook" if m * n < 300000 else "m4rm")
            got = ctx.solve(A, b, n, mode)
            _assert_same(got, want, mode)


@pytest.mark.parametrize("m,n,cap", [(100, 100, 10), (300, 200, 64), (300, 200, 65), (700, 640, 300),
                                     (2000, 1500, 700), (64, 4096, 20), (4096, 64, 20), (1111, 999, 1)])
def test_rank_deficient_matches_oracle(ctx, m, n, cap):
    rnd = random.Random(cap * 7919 + m)
    for consistent in (True, False):
        A, b = _rand_system(rnd, m, n, rank_cap=cap, consistent=consistent)
        for mode in (0, 1):
            want = oracle.solve_packed(A, b, n, mode)
            got = ctx.solve(A, b, n, mode)
            _assert_same(got, want, mode)
            if consistent:
                assert got.status == 0 and oracle.residual(A, b, n, got.origin) == 0


@pytest.mark.parametrize("density", [0.001, 0.01, 0.1])
def test_sparse_and_zero_rows(ctx, density):
    rnd = random.Random(int(density * 1e6))
    m, n = 1800, 1200
    A, b = _rand_system(rnd, m, n, consistent=True, density=density)
    A[::3] = 0  # zero rows (the reference pads with them, __init__.py:235-237)
    b[:] = 0
    for mode in (0, 1):
        _assert_same(ctx.solve(A, b, n, mode), oracle.solve_packed(A, b, n, mode), mode)
    # homogeneous through the NULL-b entry
    got = ctx.solve(A, None, n, 1)
    _assert_same(got, oracle.solve_packed(A, None, n, 1), 1)
    assert not got.origin.any()


def _near_triangular(seed, m, n, extra):
    """sparse, near-triangular, rows shuffled (the MT19937 class of systems): row i has a
    one near column i*n/m and `extra` more anywhere -- every panel of k_forward takes the all-rows
    search, collects candidate lists and sweeps the far strips over the list (sparse_sweep)"""
    rng = np.random.default_rng(seed)
    nw = (n + 63) // 64
    A = np.zeros((m, nw), dtype=np.uint64)
    rows = np.repeat(np.arange(m), extra + 1)
    cols = np.concatenate([np.minimum(n - 1, np.arange(m) * n // m)[:, None],
                           rng.integers(0, n, size=(m, extra))], axis=1).ravel()
    np.bitwise_xor.at(A, (rows, cols // 64), np.uint64(1) << (cols % 64).astype(np.uint64))
    A = A[rng.permutation(m)]
    x = rng.integers(0, 2, size=n).astype(np.int64)
    bits = np.unpackbits(A.view(np.uint8), axis=1, bitorder="little")[:, :n]
    bp = np.zeros(((m + 63) // 64) * 64, dtype=bool)
    bp[:m] = (bits.astype(np.int64) @ x) % 2 == 1
    return A, np.packbits(bp, bitorder="little").view(np.uint64).copy()


@pytest.mark.parametrize("seed,m,n,extra", [(0, 6000, 5000, 3), (1, 9000, 4000, 2), (2, 5000, 5000, 1), (3, 3000, 2500, 3)])
def test_sparse_near_triangular_matches_oracle(ctx, seed, m, n, extra):
    A, b = _near_triangular(seed, m, n, extra)
    for mode in (0, 1):
        _assert_same(ctx.solve(A, b, n, mode), oracle.solve_packed(A, b, n, mode), mode)


def test_load_in_blocks_any_order(ctx):
    """gf2b200_system_load_begin / _rows / _end: blocks in any order give the system load_host gives;
    also through 3 loopback shards (a block may straddle shards)"""
    rnd = random.Random(5)
    m, n = 1500, 1100
    A, b = _rand_system(rnd, m, n, rank_cap=800, consistent=True)
    want = oracle.solve_packed(A, b, n, 1)
    blocks = [(r, min(192, m - r)) for r in range(0, m, 192)]
    rnd.shuffle(blocks)
    for c in (ctx, _shim.Context(0, shards=3)):
        s = c.system(m, n)
        s.load_host_blocks(A, b, blocks)
        s.eliminate()
        _assert_same(s.result(1), want, 1)
        s.load_host_blocks(A, None, blocks[::-1])
        s.eliminate()
        _assert_same(s.result(1), oracle.solve_packed(A, None, n, 1), 1)


def test_all_zero_and_identity(ctx):
    n = 200
    nw = (n + 63) // 64
    A = np.zeros((n, nw), dtype=np.uint64)
    got = ctx.solve(A, None, n, 1)
    assert got.rank == 0 and got.basis.shape == (n, nw)
    _assert_same(got, oracle.solve_packed(A, None, n, 1), 1)
    for i in range(n):
        A[i, i >> 6] = np.uint64(1 << (i & 63))
    b = np.array([0xDEADBEEFCAFEF00D, 0x0123456789ABCDEF, 0xFFFFFFFFFFFFFFFF, 0x55], dtype=np.uint64)
    got = ctx.solve(A, b, n, 1)
    assert got.rank == n and got.basis.shape[0] == 0
    want = b.copy()
    want[-1] &= np.uint64((1 << (n & 63)) - 1)
    assert np.array_equal(got.origin, want)


def test_bits_above_cols_ignored_and_stride(ctx):
    # reference ignores bits >= cols (_internal.c:45,48); stride64 > ceil(n/64) is legal
    rnd = random.Random(99)
    m, n = 150, 100
    A, b = _rand_system(rnd, m, 128, consistent=False)
    wide = np.zeros((m, 5), dtype=np.uint64)
    wide[:, :2] = A
    wide[:, 2:] = np.uint64(0xFFFFFFFFFFFFFFFF)
    want = oracle.solve_packed(A, b, n, 1)
    _assert_same(ctx.solve(wide, b, n, 1), want, 1)


def test_mt19937_seed3142_golden(ctx, golden_mt32):
    # reference examples/mt.py:21,38: the deterministic golden vector
    eqs, cols, digest, state = golden_mt32
    A, b = oracle.pack_equations(eqs, cols)
    got = ctx.solve(A, b, cols, 1)
    assert got.status == 0 and got.rank == cols and got.basis.shape[0] == 0
    sol = oracle.words_to_int(got.origin)
    assert hashlib.sha256(sol.to_bytes(2496, "little")).hexdigest() == digest
    assert tuple((sol >> (32 * i)) & 0xFFFFFFFF for i in range(624)) == state


def test_mt19937_bs17(ctx, golden_mt17):
    eqs, cols, digest, _ = golden_mt17
    A, b = oracle.pack_equations(eqs, cols)
    got = ctx.solve(A, b, cols, 0)
    sol = oracle.words_to_int(got.origin)
    assert hashlib.sha256(sol.to_bytes(2496, "little")).hexdigest() == digest


@pytest.mark.parametrize("n,seed", [(1024, 1), (4096, 1), (4096, 2), (8192, 3), (5000, 4)])
def test_synthetic_device_generated_matches_oracle(ctx, n, seed):
    """The on-device generator + solve against the oracle's generator + solve."""
    sysm = ctx.system(n, n)
    sysm.generate(seed)
    sysm.eliminate()
    got = sysm.result(1)
    A, b, xstar = oracle.synth(n, n, seed)
    want = oracle.solve_packed(A, b, n, 1)
    _assert_same(got, want, 1)
    assert got.status == 0
    assert oracle.residual(A, b, n, got.origin) == 0
    assert sysm.check_synthetic(seed, got.origin) == 0
    bad = got.origin.copy()
    bad[0] ^= np.uint64(1)
    assert sysm.check_synthetic(seed, bad) > 0
    st = sysm.stats()
    assert st["rank"] == want.rank and st["sweep_launches"] > 0


def test_synthetic_32768_properties(ctx):
    """BASELINE config 3 size: residual recomputed from the seed + idempotence."""
    n, seed = 32768, 1
    sysm = ctx.system(n, n)
    sysm.generate(seed)
    sysm.eliminate()
    got = sysm.result(0)
    assert got.status == 0 and n - 8 <= got.rank <= n
    assert sysm.check_synthetic(seed, got.origin) == 0
    # free variables are zero
    free = np.ones(n, dtype=bool)
    free[got.pivcols] = False
    bits = np.unpackbits(got.origin.view(np.uint8), bitorder="little")[:n]
    assert not bits[free].any()
    # same system again -> identical answer
    sysm.generate(seed)
    sysm.eliminate()
    again = sysm.result(0)
    assert np.array_equal(again.origin, got.origin) and again.rank == got.rank


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_synthetic_32768_matches_oracle(ctx, seed):
    """BASELINE configs[2] bit for bit: rank, pivot columns (the column rank profile) and the
    particular solution against the oracle's Four-Russians tier (seconds on the host cores)."""
    n = 32768
    sysm = ctx.system(n, n)
    sysm.generate(seed)
    sysm.eliminate()
    got = sysm.result(0)
    A, b, _ = oracle.synth(n, n, seed)
    want = oracle.solve_packed(A, b, n, 0, tier="m4rm")
    _assert_same(got, want, 0)
    assert sysm.check_synthetic(seed, got.origin) == 0


def test_forward_paths_agree(ctx):
    """The one-kernel forward elimination (k_forward) and the per-panel launch chain
    (GF2B200_FORWARD=launches) are two schedules of the same arithmetic: same echelon
    form, same answer.  The second context is created with the switch set."""
    import os

    rnd = random.Random(4242)
    old = os.environ.get("GF2B200_FORWARD")
    os.environ["GF2B200_FORWARD"] = "launches"
    try:
        ctx2 = _shim.Context(0)
    finally:
        if old is None:
            del os.environ["GF2B200_FORWARD"]
        else:
            os.environ["GF2B200_FORWARD"] = old
    shapes = [(3000, 2500, None), (2500, 3000, None), (5000, 4099, 3000), (1200, 1200, 5)]
    if os.environ.get("GF2B200_TEST_EMULATION") == "1":
        shapes = [(2500, 2100, None), (1200, 1200, 5)]  # the emulated kernels are ~1000 x slower
    for (m, n, cap) in shapes:
        A, b = _rand_system(rnd, m, n, rank_cap=cap, consistent=True)
        g1, g2 = ctx.solve(A, b, n, 1), ctx2.solve(A, b, n, 1)
        _assert_same(g1, g2, 1)
        # the chain without k_sweep_apply (the sweep's tail applying the next panel): plain k_apply + k_sweep
        os.environ["GF2B200_NO_TAIL_APPLY"] = "1"
        try:
            g3 = ctx2.solve(A, b, n, 1)
        finally:
            del os.environ["GF2B200_NO_TAIL_APPLY"]
        _assert_same(g1, g3, 1)
    n = 2048 if os.environ.get("GF2B200_TEST_EMULATION") == "1" else 16384
    outs = []
    for c in (ctx, ctx2):
        sysm = c.system(n, n)
        sysm.generate(5)
        sysm.eliminate()
        outs.append((sysm.result(0), sysm.stats()))
        sysm.close()
    _assert_same(outs[0][0], outs[1][0], 0)
    assert outs[1][1]["forward_kernel_launches"] == 0
    ctx2.close()
    if outs[0][1]["forward_kernel_launches"] != 1:
        pytest.skip("this build has no k_forward (128-byte strips): both contexts ran the launch chain")


def _dup_rows_system(n, distinct, seed):
    """n x n system whose rows distinct.. repeat rows 0..: rank ~ distinct, nullity n - distinct,
    consistent (b repeats too).  Built by the library's host-side workload generator."""
    nw = (n + 63) // 64
    Ah = np.empty((distinct, nw), dtype=np.uint64)
    bh = np.zeros((distinct + 63) // 64 + 1, dtype=np.uint64)
    _shim.synth_host(Ah, bh, 0, n, seed)
    reps = -(-n // distinct)
    A = np.concatenate([Ah] * reps)[:n].copy()
    bits = np.unpackbits(bh.view(np.uint8), bitorder="little")[:distinct]
    bb = np.concatenate([bits] * reps)[:n]
    pad = np.zeros(((n + 63) // 64) * 64, dtype=np.uint8)
    pad[:n] = bb
    return A, np.packbits(pad, bitorder="little").view(np.uint64).copy()


@pytest.mark.parametrize("n,nullity", [pytest.param(4096, 1500, id="bignull-4096-1500"),
                                       pytest.param(32768, 4096, id="bignull-32768-4096")])
def test_kernel_basis_large_nullity(ctx, n, nullity):
    """the kernel basis as ONE blocked multi-right-hand-side triangular
    solve (reference: mzd_trsm_upper_left over all n - r right-hand sides, _internal.c:330-348).
    Values AND sigma order against the oracle."""
    A, b = _dup_rows_system(n, n - nullity, 11)
    want = oracle.solve_packed(A, b, n, 1, tier="m4rm")
    sysm = ctx.system(n, n)
    sysm.load_host(A, b)
    sysm.eliminate()
    got = sysm.result(1)
    st = sysm.stats()
    sysm.close()
    assert want.rank == n - nullity and got.basis.shape == (nullity, (n + 63) // 64)
    _assert_same(got, want, 1)
    assert st["basis_panels"] > 0 and st["ms_basis_solve"] > 0
    print(f"kernel basis n={n} nullity={nullity}: solve {st['ms_basis_solve']:.2f} ms, output {st['ms_basis_output']:.2f} ms, "
          f"sweep GB/s {st['basis_sweep_bytes'] / st['ms_basis_solve'] / 1e6:.0f}")
