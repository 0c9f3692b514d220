"""Host-side logic of the drop-in (CPU only): the `_internal` extension's codec,
AffineSpace/iterators and argument checks, the Python LinearSystem layer, and the
C-ABI library's exported symbols.  No device compute happens here."""
import ctypes
import pickle
import random
import re
from pathlib import Path

import numpy as np
import pytest

import oracle
import gf2bv_b200 as gf2bv
from gf2bv_b200 import _internal, _shim

ROOT = Path(__file__).resolve().parents[1]


def _ints(xs):
    return [int(x, 16) for x in xs]


# ---------------------------------------------------------------- C-ABI surface
def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "gf2b200.h").read_text()
    declared = set(re.findall(r"\b(gf2b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = ctypes.CDLL(str(_shim.LIB_PATH))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_shim.EXPORTS) == declared
    assert lib.gf2b200_abi_version() == 3


def test_ctypes_mirrors_match_header_structs():
    """The ctypes Structures in _shim.py must list the header's struct fields in order."""
    header = (ROOT / "include" / "gf2b200.h").read_text()

    def fields(struct_name):
        body = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\} " + struct_name + ";", header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        return [re.split(r"[\s\*]+", decl.strip())[-1] for decl in body.split(";") if decl.strip()]

    assert fields("gf2b200_result") == [n for n, _ in _shim.CResult._fields_]
    assert fields("gf2b200_stats") == [n for n, _ in _shim.CStats._fields_]


def test_extension_surface_matches_reference_names():
    # gf2bv/__init__.py:8-16 imports these seven; _internal.c:829-831 adds the types
    for name in ("AffineSpace", "eqs_to_sage_mat_helper", "m4ri_solve", "mul_bit_quad", "to_bits",
                 "tuple_where", "xor_tuple", "AffineSpaceIterator", "AffineSpaceIteratorSlow"):
        assert hasattr(_internal, name), name
    with pytest.raises(RuntimeError):
        _internal.eqs_to_sage_mat_helper([1], 1)


# ---------------------------------------------------------------- equation codec
def test_pack_codec_matches_oracle_restatement():
    rnd = random.Random(1)
    for cols in (1, 2, 29, 30, 31, 59, 60, 63, 64, 65, 89, 90, 128, 150, 1000):
        eqs = [rnd.getrandbits(cols + 1) for _ in range(40)]
        eqs += [0, 1, -rnd.getrandbits(cols + 1), rnd.getrandbits(cols + 70), (1 << (cols + 1)) - 1,
                1 << cols, 1 << (cols + 1), rnd.getrandbits(10)]
        a, b = _internal._pack_probe(eqs, cols)
        nw = (cols + 63) // 64
        A = np.frombuffer(a, dtype=np.uint64).reshape(len(eqs), nw)
        B = np.frombuffer(b, dtype=np.uint64)
        wantA, wantB = oracle.pack_equations(eqs, cols)
        assert np.array_equal(A, wantA), cols
        assert np.array_equal(B, wantB), cols


def test_pack_codec_block_boundaries_and_sparse_rows():
    """the packer converts 32 digits (960 bits) -> 15 words at a time and skips all-zero blocks: widths around the
    block and word boundaries, sparse rows (one or two bits anywhere), rows longer than cols"""
    rnd = random.Random(3)
    for cols in (959, 960, 961, 1023, 1024, 1025, 1919, 1920, 1921, 2879, 2880, 2881, 4096, 5000, 19968):
        eqs = []
        for _ in range(30):
            kind = rnd.randrange(4)
            if kind == 0:
                e = rnd.getrandbits(cols + 1)
            elif kind == 1:
                e = (1 << rnd.randrange(cols + 1)) | (1 << rnd.randrange(cols + 1)) | rnd.getrandbits(1)
            elif kind == 2:
                e = rnd.getrandbits(cols + 1 + rnd.choice([1, 29, 30, 31, 959, 960, 961]))
            else:
                e = rnd.getrandbits(rnd.randrange(1, cols + 2))
            eqs.append(e * rnd.choice([1, 1, -1]))
        eqs += [0, 1, (1 << (cols + 1)) - 1, 1 << cols, 1 << (cols + 1), (1 << 960), (1 << 959) | 1, 1 << 961]
        a, b = _internal._pack_probe(eqs, cols)
        A = np.frombuffer(a, dtype=np.uint64).reshape(len(eqs), -1)
        B = np.frombuffer(b, dtype=np.uint64)
        wantA, wantB = oracle.pack_equations(eqs, cols)
        assert np.array_equal(A, wantA), cols
        assert np.array_equal(B, wantB), cols


def test_pack_codec_threaded_path():
    """>= 2^21 words: rows are packed by several threads without the GIL (64-row blocks)."""
    rnd = random.Random(2)
    rows, cols = 70001, 3000
    eqs = [rnd.getrandbits(rnd.choice([1, cols // 2, cols + 1, cols + 40])) * rnd.choice([1, -1]) for _ in range(rows)]
    a, b = _internal._pack_probe(eqs, cols)
    A = np.frombuffer(a, dtype=np.uint64).reshape(rows, -1)
    B = np.frombuffer(b, dtype=np.uint64)
    wantA, wantB = oracle.pack_equations(eqs, cols)
    assert np.array_equal(A, wantA) and np.array_equal(B, wantB)


def test_pack_codec_rejects_non_ints():
    with pytest.raises(TypeError):
        _internal._pack_probe([1, 2.0], 4)


# ---------------------------------------------------------------- m4ri_solve argument checks
def test_m4ri_solve_argument_errors():
    # same checks, same order, same exception types as _internal.c:363-395,405-410
    with pytest.raises(TypeError, match="3 arguments"):
        _internal.m4ri_solve([1, 2], 2)
    with pytest.raises(TypeError, match="must be a list"):
        _internal.m4ri_solve((1, 2), 2, 0)
    with pytest.raises(ValueError, match="positive"):
        _internal.m4ri_solve([1, 2], 0, 0)
    with pytest.raises(ValueError, match="positive"):
        _internal.m4ri_solve([1, 2], -3, 0)
    with pytest.raises(ValueError, match="Invalid mode"):
        _internal.m4ri_solve([1, 2], 2, 2)
    with pytest.raises(ValueError, match="pad with zeros"):
        _internal.m4ri_solve([1], 2, 0)
    with pytest.raises(TypeError, match="must be integers"):
        _internal.m4ri_solve([1, "x"], 2, 0)
    with pytest.raises(TypeError):
        _internal.m4ri_solve([1, 2], "2", 0)


def test_no_cpu_fallback_without_device():
    if _shim.lib().gf2b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        _internal.m4ri_solve([0b011, 0b101], 2, 0)
    with pytest.raises(RuntimeError):
        gf2bv.LinearSystem([2]).solve_one([3])
    with pytest.raises(_shim.Gf2b200Error):
        _shim.Context(0)


# ---------------------------------------------------------------- AffineSpace
def test_affine_space_gray_iteration_and_get():
    sp = _internal._make_affine_space(0b1000, (0b0001, 0b0010, 0b0100), 4)
    assert type(sp) is _internal.AffineSpace
    assert sp.dimension == 3 and sp.origin == 8 and sp.basis == (1, 2, 4)
    it = iter(sp)
    assert type(it) is _internal.AffineSpaceIterator
    assert list(it) == [8, 9, 11, 10, 14, 15, 13, 12]
    assert [sp.get(i) for i in range(8)] == [8, 9, 10, 11, 12, 13, 14, 15]
    assert sp.get(8) == 8 and sp.get(-3) == sp.get(3)  # bits >= dimension ignored; magnitude only
    with pytest.raises(TypeError):
        sp.get()
    with pytest.raises(TypeError):
        sp.get("1")


@pytest.mark.parametrize("dim", [0, 1, 2, 5, 10, 63, 64, 65, 70, 130])
def test_affine_space_matches_oracle_enumeration(dim):
    rnd = random.Random(dim)
    cols = max(dim, 1) + 37
    origin = rnd.getrandbits(cols)
    basis = tuple(rnd.getrandbits(cols) | (1 << i) for i in range(dim))
    sp = _internal._make_affine_space(origin, basis, cols)
    want = oracle.OracleAffineSpace(origin, basis)
    assert sp.dimension == dim and sp.origin == origin and sp.basis == basis
    it = iter(sp)
    assert type(it) is (_internal.AffineSpaceIterator if dim <= 64 else _internal.AffineSpaceIteratorSlow)
    n = min(2 ** dim + 3, 1500)
    got = []
    for v in it:
        got.append(v)
        if len(got) >= n:
            break
    wit = iter(want)
    exp = []
    for v in wit:
        exp.append(v)
        if len(exp) >= n:
            break
    assert got == exp
    if dim <= 10:
        assert len(got) == 2 ** dim and len(set(got)) == 2 ** dim
    for i in (0, 1, 2, 3, 77, (1 << dim) - 1 if dim else 0, rnd.getrandbits(dim + 5)):
        assert sp.get(i) == want.get(i)


def test_iterator_keeps_space_alive():
    it = iter(_internal._make_affine_space(3, (4,), 3))
    import gc

    gc.collect()
    assert list(it) == [3, 7]
    assert list(it) == []  # exhausted iterators stay exhausted


# ---------------------------------------------------------------- tuple helpers
def test_tuple_helpers():
    assert _internal.to_bits(6, 0b10110) == (False, True, True, False, True, False)
    assert _internal.to_bits(3, -5) == (True, False, True)
    assert _internal.to_bits(0, 9) == ()
    assert _internal.to_bits(70, 1 << 65)[65] is True
    with pytest.raises(ValueError):
        _internal.to_bits(-1, 1)
    with pytest.raises(TypeError):
        _internal.to_bits(2, "a")
    assert _internal.xor_tuple((1, 2, 1 << 80), (3, 4, 1)) == (2, 6, (1 << 80) | 1)
    with pytest.raises(ValueError):
        _internal.xor_tuple((1,), (1, 2))
    with pytest.raises(TypeError):
        _internal.xor_tuple([1], (1,))
    with pytest.raises(TypeError):
        _internal.xor_tuple((1,), ("a",))
    cond = _internal.to_bits(3, 0b101)
    out = _internal.tuple_where(cond, (10, 11, 12), 0)
    assert out == (10, 0, 12) and out is cond  # the reference writes into cond (:672-675)
    assert _internal.tuple_where(_internal.to_bits(2, 1), 7, (8, 9)) == (7, 9)
    with pytest.raises(ValueError):
        _internal.tuple_where(_internal.to_bits(2, 1), (1, 2, 3), 0)


def test_mul_bit_quad_matches_definition():
    rnd = random.Random(5)
    for n in (1, 2, 5, 9):
        basis = [1 << i for i in range(1 + n + n * (n - 1) // 2)]
        for _ in range(20):
            a, b, v = rnd.getrandbits(n), rnd.getrandbits(n), rnd.getrandbits(n + 1)
            want, mi = v, 1 + n
            for i in range(n):
                for j in range(i):
                    if ((a >> i) & (b >> j) ^ (a >> j) & (b >> i)) & 1:
                        want |= basis[mi]
                    mi += 1
            assert _internal.mul_bit_quad(n, a, b, v, basis) == want
    with pytest.raises(ValueError):
        _internal.mul_bit_quad(3, 1, 1, 0, [1, 2])
    with pytest.raises(ValueError):
        _internal.mul_bit_quad(0, 1, 1, 0, [1])


# ---------------------------------------------------------------- Python layer vs the reference's own outputs
def test_python_layer_matches_reference_fixture(golden_small):
    g = golden_small["pylayer"]  # produced by the reference's gf2bv/__init__.py (tests/golden/make_golden.py)
    lin = gf2bv.LinearSystem(g["sizes"])
    x, y, z = lin.gens()
    zeros = [x ^ 5, 0, (y >> 1) ^ (y << 2) ^ 0x11, 7, z.rotl(3) ^ z, 0, (z & 0x0F) ^ 3, y[2] ^ x[0]]
    assert lin.get_eqs(zeros) == _ints(g["get_eqs"])
    for s, want in g["convert_sol"].items():
        assert list(lin._convert_sol(int(s, 16))) == want
    for s, want in g["evaluate"]:
        assert lin.evaluate((y >> 1) ^ (y << 2) ^ 0x11, lin._convert_sol(int(s, 16))) == want


def test_bitvec_algebra():
    lin = gf2bv.LinearSystem([4, 4])
    x, y = lin.gens()
    assert len(x) == 4 and x._bits == (2, 4, 8, 16) and y._bits == (32, 64, 128, 256)
    assert (x ^ y)._bits == (34, 68, 136, 272)
    assert (x ^ 0b0101)._bits == (3, 4, 9, 16) and (0b0101 ^ x)._bits == (3, 4, 9, 16)
    assert (x >> 1)._bits == (4, 8, 16, 0) and (x << 1)._bits == (0, 2, 4, 8) and (x >> 0) is x
    assert x.rotl(1)._bits == (16, 2, 4, 8) and x.rotr(1)._bits == (4, 8, 16, 2)
    assert (x & 0b0110)._bits == (0, 4, 8, 0) and (x & 0xF) is x
    assert (x | 0b0001)._bits == (1, 4, 8, 16) and (x | 0xF)._bits == (True,) * 4
    assert (x % 4)._bits == (2, 4, 0, 0)
    with pytest.raises(ValueError):
        x % 3
    with pytest.raises(ValueError):
        x ^ y.zeroext(1)
    assert x.sum()._bits == (30,) and x[1]._bits == (4,) and x[1:3]._bits == (4, 8)
    assert x.zeroext(2)._bits == (2, 4, 8, 16, 0, 0) and x.signext(1)._bits[-1] == 16
    assert x.broadcast(0, 3)._bits == (2, 2, 2) and x.dup(2)._bits == x._bits * 2
    assert x.concat(y)._bits == x._bits + y._bits and x.lshift_ext(1)._bits == (0,) + x._bits
    lo = (x & 0b0011) | (y & 0b1100)
    assert lo._bits == (2, 4, 128, 256)
    with pytest.raises(ValueError):
        x | y
    # evaluate: x = 0b1010, y = 0b0110 -> raw solution int
    raw = 0b1010 | (0b0110 << 4)
    assert x.evaluate(raw) == 0b1010 and (x ^ y ^ 1).evaluate(raw) == (0b1010 ^ 0b0110 ^ 1)
    assert lin.evaluate(x ^ y, (0b1010, 0b0110)) == 0b1100


def test_linear_system_plumbing_without_device(monkeypatch):
    """_solve_internal's pre-processing (gf2bv/__init__.py:229-240), with the extension
    call intercepted -- this checks the host logic only, not a solver."""
    from gf2bv_b200 import system

    calls = []
    monkeypatch.setattr(system, "m4ri_solve", lambda eqs, cols, mode: calls.append((list(eqs), cols, mode)) or None)
    lin = gf2bv.LinearSystem([2, 1])
    a, b = lin.gens()
    assert lin.solve_one([a ^ 1, 0, b]) is None
    assert calls[-1] == ([3, 4, 8], 3, 0)          # literal 0 dropped, three equations = cols, no padding
    assert list(lin.solve_all([a[0] ^ b])) == []
    assert calls[-1] == ([2 ^ 8, 0, 0], 3, 1)      # padded with zero rows to rows >= cols
    n = len(calls)
    assert lin.solve_one([a ^ a ^ 1]) is None      # literal 1 -> unsat without calling the extension
    assert len(calls) == n
    assert pickle.loads(pickle.dumps(lin))._sizes == [2, 1]
    q = pickle.loads(pickle.dumps(gf2bv.QuadraticSystem([3])))
    assert q._lin_size == 3 and q._quad_size == 3 and q._cols == 6 and len(q.gens()) == 1


def test_solve_all_dimension_guard_and_filter(monkeypatch):
    from gf2bv_b200 import system

    space = _internal._make_affine_space(0, tuple(1 << i for i in range(5)), 5)
    monkeypatch.setattr(system, "m4ri_solve", lambda eqs, cols, mode: space)
    lin = gf2bv.LinearSystem([5])
    (v,) = lin.gens()
    with pytest.raises(gf2bv.DimensionTooLargeError) as ei:
        list(lin.solve_all([v[0] ^ v[0]], max_dimension=4))
    assert ei.value.space is space and ei.value.space.dimension == 5
    sols = list(lin.solve_all([v[0] ^ v[0]], max_dimension=5))
    assert len(sols) == 32 and sols[:4] == [(0,), (1,), (3,), (2,)]


def test_quadratic_system_host_logic():
    q = gf2bv.QuadraticSystem([2, 2])
    x, y = q.gens()
    for a in (x[0], x[0] ^ y[1], x[1] ^ 1, y[0] ^ x[1] ^ 1):
        for b in (y[1], x[1] ^ y[0], y[0] ^ 1):
            assert q.mul_bit(a, b)._bits[0] == q._mul_bit_slow(a._bits[0], b._bits[0])
    with pytest.raises(ValueError):
        q.mul_bit(x, y)
    # monomial numbering: (i, j<i) row by row after the 4 linear unknowns
    assert q.mul_bit(x[1], x[0])._bits[0] == 1 << 5
    assert q.mul_bit(y[1], y[0])._bits[0] == 1 << 10
    assert q.mul_bit(x[0], x[0])._bits[0] == x[0]._bits[0]
    # convert_sol filters raw solutions whose monomial bits contradict the linear bits
    lin_bits = 0b0111
    quad = 0
    k = 0
    for i in range(4):
        for j in range(i):
            quad |= (((lin_bits >> i) & (lin_bits >> j)) & 1) << k
            k += 1
    assert q.convert_sol(lin_bits | (quad << 4)) == (0b11, 0b01)
    assert q.convert_sol(lin_bits | ((quad ^ 1) << 4)) is None
    zs = q.bit_assert(x[0] ^ y[0], 1)
    assert len(zs) == 1 + 4 and zs[0] == (x[0] ^ y[0] ^ 1)._bits[0]
    assert q.evaluate(x ^ y, (0b10, 0b11)) == 0b01
