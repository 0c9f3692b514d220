"""The oracle against the reference's golden vectors and against itself (CPU only)."""
import hashlib
import random

import numpy as np
import pytest

import oracle


def _ints(xs):
    return [int(x, 16) for x in xs]


def _pad(eqs, cols):
    return eqs + [0] * max(0, cols - len(eqs))


@pytest.mark.parametrize("tier", ["schoolbook", "m4rm"])
def test_readme_4x1_known_answer(golden_small, tier):
    g = golden_small["readme_4x1"]
    sp = oracle.m4ri_solve(_pad(_ints(g["eqs"]), 4), 4, 1, tier=tier)
    # origin (1,0,0,0), basis a^c, order (1,0,0,0) then (0,0,1,0)
    assert sp.origin == 1 and sp.basis == (5,)
    assert list(sp) == [1, 4]
    assert sp.get(0) == 1 and sp.get(1) == 4


@pytest.mark.parametrize("name", ["lfsr_galois", "lfsr_fibonacci", "xoshiro"])
@pytest.mark.parametrize("tier", ["schoolbook", "m4rm"])
def test_reference_unique_solution_examples(golden_small, name, tier):
    g = golden_small[name]
    cols = g["cols"]
    sp = oracle.m4ri_solve(_pad(_ints(g["eqs"]), cols), cols, 1, tier=tier)
    assert sp.dimension == 0
    sol, s = [], sp.origin
    for size in g["sizes"]:
        sol.append(s & ((1 << size) - 1))
        s >>= size
    assert [sol] == g["solve_all"]


@pytest.mark.parametrize("name", ["simple_linear", "simple_affine"])
def test_simple_examples_match_bigint(golden_small, name):
    g = golden_small[name]
    eqs = _pad(_ints(g["eqs"]), g["cols"])
    for tier in ("schoolbook", "m4rm"):
        sp = oracle.m4ri_solve(eqs, g["cols"], 1, tier=tier)
        assert sp.origin == int(g["origin"], 16)
        assert list(sp.basis) == _ints(g["basis"])
        assert sp.dimension == 3


def test_unsat(golden_small):
    g = golden_small["unsat"]
    assert oracle.m4ri_solve(_ints(g["eqs"]), g["cols"], 0) is None
    assert oracle.m4ri_solve(_ints(g["eqs"]), g["cols"], 1) is None


def test_mt19937_seed3142_golden(golden_mt32):
    # reference examples/mt.py:21,38 -- the deterministic golden vector
    eqs, cols, digest, state = golden_mt32
    assert (len(eqs), cols) == (20000, 19968)
    sol = oracle.m4ri_solve(eqs, cols, 0)
    assert hashlib.sha256(sol.to_bytes(2496, "little")).hexdigest() == digest
    assert digest == "2f79f22d6342e077883a3d624be1db6c2d630489255ec20affda7018a148304f"
    assert tuple((sol >> (32 * i)) & 0xFFFFFFFF for i in range(624)) == state
    assert state == tuple(random.Random(3142).getstate()[1][:-1])


def test_mt19937_bs17(golden_mt17):
    eqs, cols, digest, state = golden_mt17
    sp = oracle.m4ri_solve(eqs, cols, 1)
    assert sp.dimension == 0
    assert hashlib.sha256(sp.origin.to_bytes(2496, "little")).hexdigest() == digest


def test_tiers_agree_random():
    rnd = random.Random(7)
    for _ in range(120):
        cols = rnd.choice([1, 2, 5, 17, 63, 64, 65, 100, 128, 130, 200, 321])
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
        for mode in (0, 1):
            a = oracle.solve_bigint(eqs, cols, mode)
            b = oracle.m4ri_solve(eqs, cols, mode, tier="schoolbook")
            c = oracle.m4ri_solve(eqs, cols, mode, tier="m4rm")
            if a is None:
                assert b is None and c is None
            elif mode == 0:
                assert a == b == c
            else:
                assert a.origin == b.origin == c.origin
                assert a.basis == b.basis == c.basis


def test_sigma_order_example():
    # n=4, pivots {1,3} -> sigma [1,3,2,0], free order [2,0]
    eqs = [0b00100, 0b10000, 0, 0]  # x1 = 0, x3 = 0
    sp = oracle.m4ri_solve(eqs, 4, 1)
    assert sp.basis == (1 << 2, 1 << 0)


def test_iterators_and_get():
    sp = oracle.OracleAffineSpace(0b1000, (0b0001, 0b0010, 0b0100))
    assert list(sp) == [8, 9, 11, 10, 14, 15, 13, 12]  # Gray order (_internal.c:101-122)
    assert [sp.get(i) for i in range(8)] == [8, 9, 10, 11, 12, 13, 14, 15]
    big = oracle.OracleAffineSpace(0, tuple(1 << i for i in range(65)))
    it = iter(big)  # dimension > 64 -> binary counter (_internal.c:63-91,185)
    assert [next(it) for _ in range(5)] == [0, 1, 2, 3, 4]
    assert list(oracle.OracleAffineSpace(5, ())) == [5]


def test_argument_errors():
    with pytest.raises(TypeError):
        oracle.m4ri_solve((1, 2), 2, 0)
    with pytest.raises(ValueError):
        oracle.m4ri_solve([1, 2], 0, 0)
    with pytest.raises(ValueError):
        oracle.m4ri_solve([1, 2], 2, 2)
    with pytest.raises(ValueError):
        oracle.m4ri_solve([1], 2, 0)
    with pytest.raises(TypeError):
        oracle.m4ri_solve([1, "x"], 2, 0)


def test_high_bits_and_sign_ignored():
    # bits above cols ignored (_internal.c:45,48); digits read by magnitude (:10,14)
    assert oracle.m4ri_solve([0b011 | (1 << 40), 0b101], 2, 0) == 0b11
    assert oracle.m4ri_solve([-0b011, 0b101], 2, 0) == 0b11


def test_synth_dense_consistent_and_residual():
    for n in (64, 257, 1024):
        A, b, x = oracle.synth(n, n, 1)
        s = oracle.solve_packed(A, b, n, 1)
        assert s.status == 0 and n - 3 <= s.rank <= n
        assert oracle.residual(A, b, n, s.origin) == 0
        s2 = oracle.solve_packed(A, b, n, 1, tier="schoolbook")
        assert np.array_equal(s.origin, s2.origin) and np.array_equal(s.basis, s2.basis)
        for v in s.basis:
            assert oracle.residual(A, None, n, v) == 0
