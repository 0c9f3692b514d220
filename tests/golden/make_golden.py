#!/usr/bin/env python
"""Generate tests/golden/*.{json,npz} by IMPORTING THE REFERENCE'S PYTHON LAYER.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference's C extension cannot be built here (M4RI absent), so
``gf2bv._internal`` is replaced by pure-Python stand-ins for the tuple helpers
(restating _internal.c:504-676) and ``m4ri_solve`` is routed to the CPU oracle.
What the fixtures therefore pin:

* inputs  : the exact equation lists the REFERENCE'S OWN Python code
            (BitVec algebra, crypto models, get_eqs) produces for its examples;
* outputs : for unique-solution systems, the value the reference's example
            asserts demand (examples/mt.py:38, lfsr.py:20, xoshiro.py:16) -- the
            script re-runs those asserts with the oracle as the solver;
            ``get_eqs`` / ``_convert_sol`` outputs come from reference code itself;
* for underdetermined systems (README 4x1, examples/simple.py) the outputs are the
  ORACLE's (M4RI semantics) -- labelled "unpinned" in the file.
"""
from __future__ import annotations

import hashlib
import json
import random
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402

OUT = Path(__file__).resolve().parent


# ---- stand-in for gf2bv._internal (tuple helpers only) --------------------
def to_bits(n, a):
    a = abs(a)
    return tuple(bool((a >> i) & 1) for i in range(n))


def xor_tuple(a, b):
    if len(a) != len(b):
        raise ValueError("The length of a and b is not equal")
    return tuple(x ^ y for x, y in zip(a, b))


def tuple_where(cond, a, b):
    al, bl = isinstance(a, tuple), isinstance(b, tuple)
    return tuple((a[i] if al else a) if c else (b[i] if bl else b) for i, c in enumerate(cond))


def mul_bit_quad(n, a, b, v, basis):
    ab, bb = to_bits(n, a), to_bits(n, b)
    mi = 1 + n
    for i in range(n):
        for j in range(i):
            if (ab[i] & bb[j]) ^ (ab[j] & bb[i]):
                v |= basis[mi]
            mi += 1
    return v


stub = types.ModuleType("gf2bv._internal")
stub.AffineSpace = oracle.OracleAffineSpace
stub.eqs_to_sage_mat_helper = lambda *a: (_ for _ in ()).throw(RuntimeError("no libgd"))
stub.m4ri_solve = oracle.m4ri_solve
stub.mul_bit_quad = mul_bit_quad
stub.to_bits = to_bits
stub.tuple_where = tuple_where
stub.xor_tuple = xor_tuple
sys.modules["gf2bv._internal"] = stub
sys.path.insert(0, "/root/reference")
import gf2bv  # noqa: E402  (the reference's own gf2bv/__init__.py)
from gf2bv import LinearSystem  # noqa: E402
from gf2bv.crypto.lfsr import FibonacciLFSR, GaloisLFSR  # noqa: E402
from gf2bv.crypto.mt import MT19937  # noqa: E402
from gf2bv.crypto.xoshiro import Xoshiro256starstar  # noqa: E402

assert gf2bv.__file__.startswith("/root/reference")


def hexl(xs):
    return [hex(x) for x in xs]


def sparse_pack(eqs):
    idx, off = [], [0]
    for e in eqs:
        while e:
            low = e & -e
            idx.append(low.bit_length() - 1)
            e ^= low
        off.append(len(idx))
    return np.array(idx, dtype=np.uint16), np.array(off, dtype=np.uint32)


small = {}

# ---- README.md:29-45 (4 x 1-bit) ------------------------------------------
lin = LinearSystem([1, 1, 1, 1])
a, b, c, d = lin.gens()
zeros = [a ^ b ^ c ^ 1, b ^ d, a ^ c ^ 1]
eqs = lin.get_eqs(zeros)
sp = oracle.solve_bigint(eqs + [0] * (4 - len(eqs)), 4, 1)
small["readme_4x1"] = {
    "sizes": [1, 1, 1, 1], "cols": 4, "eqs": hexl(eqs),
    "pinned": "unpinned (oracle, SURVEY A.5)",
    "origin": hex(sp.origin), "basis": hexl(sp.basis),
    "solve_all": [list(lin.convert_sol(s)) for s in sp],
    "solve_one": list(lin.convert_sol(sp.origin)),
}

# ---- examples/simple.py:30-37 (linear) ------------------------------------
sys.path.insert(0, "/root/reference/examples")
import simple as ex_simple  # noqa: E402

lin = LinearSystem((64, 64))
xs, ys = lin.gens()
zeros = list(ex_simple.magic(xs, ys))
eqs = lin.get_eqs(zeros)
padded = eqs + [0] * max(0, 128 - len(eqs))
sp = oracle.solve_bigint(padded, 128, 1)
sols = [lin.convert_sol(s) for s in sp]
assert all(ex_simple.magic(*s) == (0, 0, 0) for s in sols)  # examples/simple.py:16-18
small["simple_linear"] = {
    "sizes": [64, 64], "cols": 128, "eqs": hexl(eqs),
    "pinned": "property only in the reference (simple.py:18,23,27); values = oracle",
    "origin": hex(sp.origin), "basis": hexl(sp.basis),
    "solve_all": [list(s) for s in sols],
}

# ---- examples/simple.py:40-50 (affine), seeded ----------------------------
rnd = random.Random(20261017)
inp = rnd.getrandbits(64), rnd.getrandbits(64)
z = ex_simple.magic(*inp)
lin = LinearSystem((64, 64))
xs, ys = lin.gens()
z1s, z2s, z3s = ex_simple.magic(xs, ys)
zeros = [z1s ^ z[0], z2s ^ z[1], z3s ^ z[2]]
eqs = lin.get_eqs(zeros)
padded = eqs + [0] * max(0, 128 - len(eqs))
sp = oracle.solve_bigint(padded, 128, 1)
sols = [lin.convert_sol(s) for s in sp]
assert all(ex_simple.magic(*s) == z for s in sols)
assert inp in sols
small["simple_affine"] = {
    "sizes": [64, 64], "cols": 128, "eqs": hexl(eqs), "target": list(z), "input": list(inp),
    "pinned": "property only in the reference; values = oracle",
    "origin": hex(sp.origin), "basis": hexl(sp.basis),
    "solve_all": [list(s) for s in sols],
}

# ---- examples/lfsr.py (unique solution), seeded ---------------------------
for name, cls, mask in (("lfsr_galois", GaloisLFSR, 0x5C2B76970103D4EEFCD4A2C681CC400D),
                        ("lfsr_fibonacci", FibonacciLFSR, 0x6D6AC812F52A212D5A0B9F3117801FD5)):
    init_st = rnd.getrandbits(128)
    l1 = cls(128, mask, init_st)
    out = [l1() for _ in range(256)]
    lin = LinearSystem([128])
    (sym,) = lin.gens()
    l2 = cls(128, mask, sym)
    zeros = [l2() ^ o for o in out]
    eqs = lin.get_eqs(zeros)
    got = list(lin.solve_all(zeros))
    assert got == [(init_st,)]  # examples/lfsr.py:18-20
    small[name] = {"sizes": [128], "cols": 128, "eqs": hexl(eqs),
                   "pinned": "reference assert examples/lfsr.py:20 (unique)",
                   "solve_all": [[init_st]]}

# ---- examples/xoshiro.py (unique solution), seeded ------------------------
st = [rnd.getrandbits(64) for _ in range(4)]
xos = Xoshiro256starstar(list(st))
out = [xos() for _ in range(10)]
lin = LinearSystem([64] * 4)
xos2 = Xoshiro256starstar(lin.gens())
zeros = [xos2.step() ^ Xoshiro256starstar.untemper(o) for o in out]
eqs = lin.get_eqs(zeros)
got = list(lin.solve_all(zeros))
assert got == [tuple(st)]
small["xoshiro"] = {"sizes": [64] * 4, "cols": 256, "eqs": hexl(eqs),
                    "pinned": "reference assert examples/xoshiro.py:13-16 (unique)",
                    "solve_all": [list(st)]}

# ---- python-layer I/O produced by reference code itself -------------------
lin = LinearSystem([3, 5, 8])
x, y, zv = lin.gens()
zeros = [x ^ 5, 0, (y >> 1) ^ (y << 2) ^ 0x11, 7, zv.rotl(3) ^ zv, 0, (zv & 0x0F) ^ 3, y[2] ^ x[0]]
small["pylayer"] = {
    "sizes": [3, 5, 8],
    "get_eqs": hexl(lin.get_eqs(zeros)),
    "zeros_desc": "x^5, 0, (y>>1)^(y<<2)^0x11, 7, rotl(z,3)^z, 0, (z&0x0F)^3, y[2]^x[0]",
    "convert_sol": {hex(s): list(lin._convert_sol(s)) for s in (0, 0xFFFF, 0xA5C3, 0x1234, 0x8001)},
    "evaluate": [[hex(s), lin.evaluate((y >> 1) ^ (y << 2) ^ 0x11, lin._convert_sol(s))]
                 for s in (0, 0xFFFF, 0xA5C3)],
}

# unsat + literal-1 early-out cases (gf2bv/__init__.py:231-233)
lin = LinearSystem([2])
(v,) = lin.gens()
small["unsat"] = {"sizes": [2], "cols": 2,
                  "eqs": hexl(lin.get_eqs([v ^ 1, v ^ 2])), "solve_one": None}

(OUT / "small_systems.json").write_text(json.dumps(small, indent=1))

# ---- examples/mt.py:19-45, bs=32, seed 3142 (the deterministic golden) ----
for bs, samples in ((32, None), (17, None)):
    rand = random.Random(3142)
    st = tuple(rand.getstate()[1][:-1])
    effective_bs = ((bs - 1) & bs) or bs
    ns = 624 * 32 // effective_bs if samples is None else samples
    out = [rand.getrandbits(bs) for _ in range(ns)]
    lin = LinearSystem([32] * 624)
    mt = lin.gens()
    rng = MT19937(mt)
    zeros = [rng.getrandbits(bs) ^ o for o in out] + [mt[0] ^ 0x80000000]
    eqs = lin.get_eqs(zeros)
    cols = 19968
    if len(eqs) < cols:
        eqs = eqs + [0] * (cols - len(eqs))
    sol = oracle.m4ri_solve(eqs, cols, 0)
    assert lin.convert_sol(sol) == st  # examples/mt.py:38
    digest = hashlib.sha256(sol.to_bytes(2496, "little")).hexdigest()
    if bs == 32:
        assert digest == "2f79f22d6342e077883a3d624be1db6c2d630489255ec20affda7018a148304f"
    idx, off = sparse_pack(eqs)
    np.savez_compressed(OUT / f"mt19937_bs{bs}.npz", idx=idx, off=off,
                        cols=np.int64(cols), sha256=np.array(digest),
                        state=np.array(st, dtype=np.uint32))
    print("mt bs", bs, len(eqs), "x", cols, digest)

print("wrote", sorted(p.name for p in OUT.iterdir()))
