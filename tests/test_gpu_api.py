"""The reference-facing API on the GPU: `_internal.m4ri_solve` and
`LinearSystem.solve_one/solve_all` (the plugin surface, reference
gf2bv/__init__.py:229-277 + _internal.c:359-502) against the oracle and the
committed golden fixtures.  Bit-exact, including kernel-basis order and the
solve_all enumeration order."""
import hashlib
import random
import threading

import pytest

import oracle
import gf2bv_b200 as gf2bv
from gf2bv_b200 import _internal

pytestmark = pytest.mark.gpu


def _ints(xs):
    return [int(x, 16) for x in xs]


def _pad(eqs, cols):
    return eqs + [0] * max(0, cols - len(eqs))


def _same_space(got, want):
    if want is None:
        assert got is None
        return
    assert type(got) is _internal.AffineSpace
    assert got.dimension == want.dimension
    assert got.origin == want.origin
    assert got.basis == tuple(want.basis)


def test_readme_4x1(golden_small):
    # README.md:32-44 through the public API; expected order from SURVEY.md A.5
    lin = gf2bv.LinearSystem([1, 1, 1, 1])
    a, b, c, d = lin.gens()
    zeros = [a ^ b ^ c ^ 1, b ^ d, a ^ c ^ 1]
    g = golden_small["readme_4x1"]
    assert lin.get_eqs(zeros) == _ints(g["eqs"])
    assert list(lin.solve_all(zeros)) == [tuple(s) for s in g["solve_all"]]
    assert lin.solve_one(zeros) == tuple(g["solve_one"])
    sp = lin.solve_raw_space(zeros)
    assert sp.dimension == 1 and sp.origin == 1 and sp.basis == (5,)
    assert lin.solve_raw_one(zeros) == 1


@pytest.mark.parametrize("name", ["simple_linear", "simple_affine"])
def test_simple_examples(golden_small, name):
    # examples/simple.py: 128 columns, nullity 3; values/order = oracle (unpinned by the reference)
    g = golden_small[name]
    lin = gf2bv.LinearSystem(g["sizes"])
    eqs = _ints(g["eqs"])
    sp = _internal.m4ri_solve(_pad(eqs, g["cols"]), g["cols"], 1)
    assert sp.dimension == 3
    assert sp.origin == int(g["origin"], 16) and list(sp.basis) == _ints(g["basis"])
    assert [list(s) for s in lin.solve_all(eqs)] == g["solve_all"]
    assert list(lin.solve_one(eqs)) == g["solve_all"][0]
    # every enumerated solution satisfies every equation (the reference's own property, simple.py:18)
    for raw in sp:
        point = (raw << 1) | 1
        assert all(((e & point).bit_count() & 1) == 0 for e in eqs)


@pytest.mark.parametrize("name", ["lfsr_galois", "lfsr_fibonacci", "xoshiro"])
def test_unique_solution_examples(golden_small, name):
    # examples/lfsr.py:20, xoshiro.py:16: unique solution == the generator's initial state
    g = golden_small[name]
    lin = gf2bv.LinearSystem(g["sizes"])
    eqs = _ints(g["eqs"])
    assert [list(s) for s in lin.solve_all(eqs)] == g["solve_all"]
    assert list(lin.solve_one(eqs)) == g["solve_all"][0]


def test_unsat(golden_small):
    g = golden_small["unsat"]
    lin = gf2bv.LinearSystem(g["sizes"])
    eqs = _ints(g["eqs"])
    assert lin.solve_one(eqs) is None and list(lin.solve_all(eqs)) == []
    assert _internal.m4ri_solve(eqs, g["cols"], 0) is None
    assert _internal.m4ri_solve(eqs, g["cols"], 1) is None


def test_mt19937_state_recovery(golden_mt32):
    # examples/mt.py:21,38 (seed 3142): sol == random.Random(3142) state
    eqs, cols, digest, state = golden_mt32
    lin = gf2bv.LinearSystem([32] * 624)
    assert lin.solve_one(eqs) == state
    raw = _internal.m4ri_solve(eqs, cols, 0)
    assert hashlib.sha256(raw.to_bytes(2496, "little")).hexdigest() == digest
    assert state == tuple(random.Random(3142).getstate()[1][:-1])
    assert list(lin.solve_all(eqs)) == [state]


def test_random_systems_match_oracle_through_m4ri_solve():
    rnd = random.Random(11)
    for trial in range(60):
        cols = rnd.choice([1, 2, 5, 17, 63, 64, 65, 100, 128, 130, 200, 321, 700])
        rows = cols + rnd.choice([0, 0, 1, 5, 40])
        base = [rnd.getrandbits(cols + 1) & ~1 for _ in range(rnd.randint(0, cols))]
        eqs = []
        for _ in range(rows):
            v = 0
            for bv in base:
                if rnd.random() < 0.5:
                    v ^= bv
            eqs.append(v)
        if rnd.random() < 0.6:
            x = rnd.getrandbits(cols)
            eqs = [(e & ~1) | (bin((e >> 1) & x).count("1") & 1) for e in eqs]
        else:
            eqs = [e ^ (rnd.random() < 0.05) for e in eqs]
        if trial % 7 == 0:
            eqs = [-e if rnd.random() < 0.3 else e | (rnd.getrandbits(9) << (cols + 1)) for e in eqs]
        want0 = oracle.m4ri_solve(eqs, cols, 0)
        assert _internal.m4ri_solve(eqs, cols, 0) == want0
        got = _internal.m4ri_solve(eqs, cols, 1)
        want = oracle.m4ri_solve(eqs, cols, 1)
        _same_space(got, want)
        if want is not None and want.dimension <= 8:
            assert list(got) == list(want)
            assert [got.get(i) for i in range(1 << want.dimension)] == \
                   [want.get(i) for i in range(1 << want.dimension)]


def test_streamed_pack_matches_oracle(monkeypatch):
    """large systems are packed by worker threads in blocks that stream to the GPU while the rest is
    still being packed (gf2b200_system_load_begin / _rows / _end); the thresholds are lowered so the
    path runs on small systems: full rank, rank-deficient, inconsistent, homogeneous, both modes"""
    monkeypatch.setenv("GF2B200_PACK_MIN_WORDS", "64")
    monkeypatch.setenv("GF2B200_PACK_BLOCK_BYTES", "4096")
    rnd = random.Random(77)
    for cols, rows, cap in [(300, 700, None), (257, 1000, 100), (640, 641, None), (130, 900, 0)]:
        basis = [rnd.getrandbits(cols + 1) & ~1 for _ in range(cols if cap is None else cap)]
        eqs = []
        for _ in range(rows):
            v = 0
            for bv in rnd.sample(basis, min(len(basis), 7)):
                v ^= bv
            eqs.append(v)
        x = rnd.getrandbits(cols)
        consistent = [(e & ~1) | (bin((e >> 1) & x).count("1") & 1) for e in eqs]
        noisy = [e ^ (rnd.random() < 0.02) for e in consistent]
        for sysm in (consistent, noisy, eqs):
            for mode in (0, 1):
                got, want = _internal.m4ri_solve(sysm, cols, mode), oracle.m4ri_solve(sysm, cols, mode)
                if mode == 0:
                    assert got == want
                else:
                    _same_space(got, want)


def test_dimension_too_large_error_carries_space():
    lin = gf2bv.LinearSystem([40])
    (v,) = lin.gens()
    zeros = [v[0] ^ v[1], v[39] ^ 1]
    with pytest.raises(gf2bv.DimensionTooLargeError) as ei:
        list(lin.solve_all(zeros))
    sp = ei.value.space
    want = oracle.m4ri_solve(_pad(lin.get_eqs(zeros), 40), 40, 1)
    assert sp.dimension == 38
    _same_space(sp, want)
    assert sp.get(5) == want.get(5)
    assert lin.solve_one(zeros) == (1 << 39,)


def test_large_nullity_uses_slow_iterator():
    cols = 100
    eqs = _pad([0b110, 1 | (1 << 100)], cols)
    sp = _internal.m4ri_solve(eqs, cols, 1)
    want = oracle.m4ri_solve(eqs, cols, 1)
    _same_space(sp, want)
    assert sp.dimension == 98
    it = iter(sp)
    assert type(it) is _internal.AffineSpaceIteratorSlow
    wi = iter(want)
    assert [next(it) for _ in range(40)] == [next(wi) for _ in range(40)]


def test_quadratic_system_small():
    # x0*x1 = 1, x0 ^ x2 = 0  over 3 unknowns -> x = (1,1,1)
    q = gf2bv.QuadraticSystem([3])
    (x,) = q.gens()
    zeros = [q.mul_bit(x[0], x[1]) ^ 1, x[0] ^ x[2]]
    zeros += q.bit_assert(x[0], 1)
    sols = list(q.solve_all(zeros))
    assert (0b111,) in sols
    for s in sols:
        assert q.evaluate(x[0] ^ x[2], s) == 0 and (s[0] & 1) == 1 and (s[0] >> 1) & 1 == 1
    assert q.solve_one(zeros) == sols[0]


def _lfsr_step(state, nbits, tapmask):
    """One step of a Fibonacci LFSR on an int or a BitVec: shift right, feedback into the top bit."""
    if isinstance(state, int):
        fb = bin(state & tapmask).count("1") & 1
        return (state >> 1) | (fb << (nbits - 1))
    fb = (state & tapmask).sum()
    return (state >> 1) ^ (fb.broadcast(0, nbits) & (1 << (nbits - 1)))


@pytest.mark.parametrize("nbits,nout,seed,nullity", [(24, 600, 4, 2), (24, 600, 1, 26), (40, 1700, 3, 3),
                                                     pytest.param(128, 9000, 1, 0, id="full128")])
def test_filter_generator_by_linearisation(nbits, nout, seed, nullity):
    """The workload class of the reference's examples/nlfsr.py (QuadraticSystem feeding the
    solve path with n + n(n-1)/2 columns; 128 bits -> 8256 columns as in nlfsr.py:45): a
    quadratic filter on an LFSR state, recovered through solve_all / solve_one.  State, taps
    and filter are seeded, so the answer is pinned; the raw solution space must equal the
    oracle's (origin, kernel basis values and order)."""
    rnd = random.Random(seed)
    tapmask = rnd.getrandbits(nbits) | 1 | (1 << (nbits - 1))
    pick = rnd.sample(range(nbits), 4)
    init = rnd.getrandbits(nbits) | 1
    q = gf2bv.QuadraticSystem([nbits])
    (x,) = q.gens()
    st, sym, zeros, outs = init, x, [], []
    for _ in range(nout):
        st, sym = _lfsr_step(st, nbits, tapmask), _lfsr_step(sym, nbits, tapmask)
        a, b, c, d = [(st >> i) & 1 for i in pick]
        A, B, C, D = [sym[i] for i in pick]
        outs.append((a & b) ^ (c & d) ^ a ^ c)
        zeros.append(q.mul_bit(A, B) ^ q.mul_bit(C, D) ^ A ^ C ^ outs[-1])
    cols = nbits + nbits * (nbits - 1) // 2
    eqs = q.get_eqs(zeros)
    eqs = eqs + [0] * max(0, cols - len(eqs))
    got, want = _internal.m4ri_solve(eqs, cols, 1), oracle.m4ri_solve(eqs, cols, 1)
    assert got.dimension == want.dimension == nullity
    assert got.origin == want.origin and tuple(got.basis) == tuple(want.basis)
    if nullity > 16:
        with pytest.raises(gf2bv.DimensionTooLargeError) as ei:
            list(q.solve_all(zeros))
        assert ei.value.space.dimension == nullity
        return
    sols = list(q.solve_all(zeros))
    assert (init,) in sols and q.solve_one(zeros) == sols[0]
    for (sol,) in sols:  # every solution that survives the monomial filter reproduces the outputs
        st = sol
        for o in outs[:300]:
            st = _lfsr_step(st, nbits, tapmask)
            a, b, c, d = [(st >> i) & 1 for i in pick]
            assert ((a & b) ^ (c & d) ^ a ^ c) == o


def test_concurrent_calls_from_threads():
    # the reference releases the GIL around the solve (_internal.c:429); concurrent callers are legal
    rnd = random.Random(3)
    jobs = []
    for _ in range(6):
        cols = rnd.choice([64, 130, 257])
        eqs = [rnd.getrandbits(cols + 1) for _ in range(cols + 3)]
        jobs.append((eqs, cols, oracle.m4ri_solve(eqs, cols, 0)))
    out = [None] * len(jobs)

    def run(i):
        out[i] = _internal.m4ri_solve(jobs[i][0], jobs[i][1], 0)

    ts = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert out == [j[2] for j in jobs]
