"""CPU oracle for the gf2bv hot path -- TEST INFRASTRUCTURE, never the product.

Restates, on the CPU, what the reference computes on
``LinearSystem.solve_one/solve_all -> _internal.m4ri_solve -> M4RI``
(reference ``gf2bv/_internal.c:359-502``; semantics in ``SURVEY.md`` Appendix A).

Three independent statements of the same semantics live here and are checked
against each other by ``tests/test_oracle.py``:

* ``solve_bigint``      -- pure-Python big-int Gauss-Jordan (small cases only),
* ``gf2_oracle.c`` tier 1 ``gf2o_solve_schoolbook`` -- the spec in C,
* ``gf2_oracle.c`` tier 2 ``gf2o_solve_m4rm``       -- blocked Four-Russians +
  OpenMP port, the timed CPU baseline (``cpu_baseline.kind == "port"``).

PARITY PINNING: M4RI (the third-party library holding the reference's
arithmetic, setup.py:14-17) is absent from this image, so there is no
``oracle/_ref`` build.  Unique-solution systems are pinned by the reference's own
example asserts (fixtures in ``tests/golden``); for underdetermined systems the
reference has no golden vectors -> "parity unpinned" there, pinned here to M4RI's
documented semantics (free variables 0, kernel basis in sigma order).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this package.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
from pathlib import Path
from typing import Iterator, Optional

import numpy as np

_HERE = Path(__file__).resolve().parent


# --------------------------------------------------------------------------
# build / load the C oracle (compiled on the host that runs it: -march=native)
# --------------------------------------------------------------------------
def _cpu_tag() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.sha1(line.encode()).hexdigest()[:10]
    except OSError:
        pass
    return "generic"


def build(force: bool = False) -> Path:
    """Compile gf2_oracle.c for this host's CPU; returns the .so path."""
    out = _HERE / "_build" / f"libgf2oracle-{_cpu_tag()}.so"
    src = _HERE / "gf2_oracle.c"
    if force or not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        out.parent.mkdir(exist_ok=True)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        tmp = out.with_suffix(f".tmp{os.getpid()}.so")
        subprocess.check_call(
            [cc, "-O3", "-march=native", "-mtune=native", "-fopenmp", "-fPIC", "-std=c11",
             "-shared", "-o", str(tmp), str(src)]
        )
        os.replace(tmp, out)
    return out


class _Result(ctypes.Structure):
    _fields_ = [
        ("status", ctypes.c_int32),
        ("rank", ctypes.c_int64),
        ("kernel_dim", ctypes.c_int64),
        ("origin", ctypes.POINTER(ctypes.c_uint64)),
        ("basis", ctypes.POINTER(ctypes.c_uint64)),
        ("pivcols", ctypes.POINTER(ctypes.c_int64)),
        ("t_forward", ctypes.c_double),
        ("t_backward", ctypes.c_double),
    ]


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
        u64p = ctypes.POINTER(ctypes.c_uint64)
        for name in ("gf2o_solve_schoolbook", "gf2o_solve_m4rm"):
            fn = getattr(_lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [u64p, u64p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                           ctypes.c_int, ctypes.POINTER(_Result)]
        _lib.gf2o_result_free.argtypes = [ctypes.POINTER(_Result)]
        _lib.gf2o_result_free.restype = None
        _lib.gf2o_threads.restype = ctypes.c_int
        _lib.gf2o_set_threads.argtypes = [ctypes.c_int]
        _lib.gf2o_set_threads.restype = None
        _lib.gf2o_synth.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, u64p, u64p]
        _lib.gf2o_synth.restype = None
        _lib.gf2o_synth_xstar.argtypes = [ctypes.c_int64, ctypes.c_uint64, u64p]
        _lib.gf2o_synth_xstar.restype = None
        _lib.gf2o_residual.argtypes = [u64p, u64p, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_int64, u64p]
        _lib.gf2o_residual.restype = ctypes.c_int64
    return _lib


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.uint64 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


class Solution:
    """Packed result of one solve (words are little-endian 64-bit limbs)."""

    def __init__(self, status, rank, origin, basis, pivcols, t_forward=0.0, t_backward=0.0):
        self.status = status          # 0 ok, 1 inconsistent
        self.rank = rank
        self.origin = origin          # np.uint64[nw] or None
        self.basis = basis            # np.uint64[dim, nw] (mode 1) or None
        self.pivcols = pivcols
        self.t_forward = t_forward
        self.t_backward = t_backward


def solve_packed(A: np.ndarray, b: Optional[np.ndarray], n: int, mode: int = 0,
                 tier: str = "m4rm") -> Solution:
    """A: uint64[m, stride] row-major bit matrix; b: uint64[ceil(m/64)] packed bits."""
    assert A.ndim == 2
    m, stride = A.shape
    nw = (n + 63) // 64
    res = _Result()
    fn = lib().gf2o_solve_schoolbook if tier == "schoolbook" else lib().gf2o_solve_m4rm
    A = np.ascontiguousarray(A, dtype=np.uint64)
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.uint64)
    rc = fn(_p(A), _p(b), m, n, stride, mode, ctypes.byref(res))
    if rc < 0:
        raise MemoryError("oracle out of memory")
    try:
        if rc == 1:
            return Solution(1, int(res.rank), None, None, None, res.t_forward, res.t_backward)
        origin = np.ctypeslib.as_array(res.origin, shape=(max(nw, 1),)).copy()[:nw]
        piv = np.ctypeslib.as_array(res.pivcols, shape=(max(int(res.rank), 1),)).copy()[: int(res.rank)]
        basis = None
        if mode == 1:
            d = int(res.kernel_dim)
            if d:
                basis = np.ctypeslib.as_array(res.basis, shape=(d, nw)).copy()
            else:
                basis = np.zeros((0, nw), dtype=np.uint64)
        return Solution(0, int(res.rank), origin, basis, piv, res.t_forward, res.t_backward)
    finally:
        lib().gf2o_result_free(ctypes.byref(res))


def synth(m: int, n: int, seed: int = 1):
    """Dense synthetic system of SURVEY.md 8(d): returns (A[m,nw], b[ceil(m/64)], xstar[nw])."""
    nw = (n + 63) // 64
    A = np.empty((m, nw), dtype=np.uint64)
    b = np.zeros(((m + 63) // 64,), dtype=np.uint64)
    x = np.empty((nw,), dtype=np.uint64)
    lib().gf2o_synth(m, n, seed, _p(A), _p(b))
    lib().gf2o_synth_xstar(n, seed, _p(x))
    return A, b, x


def residual(A: np.ndarray, b: Optional[np.ndarray], n: int, x: np.ndarray) -> int:
    m, stride = A.shape
    return int(lib().gf2o_residual(_p(np.ascontiguousarray(A)), _p(b), m, n, stride,
                                   _p(np.ascontiguousarray(x, dtype=np.uint64))))


def threads() -> int:
    return int(lib().gf2o_threads())


def set_threads(n: int) -> None:
    """Fix the OpenMP thread count of the tier-2 port (bench.py: torchrun exports
    OMP_NUM_THREADS=1, which must not apply to the CPU baseline)."""
    lib().gf2o_set_threads(int(n))


# --------------------------------------------------------------------------
# Python-int <-> packed words (restates _internal.c:403-426 and :32-39)
# --------------------------------------------------------------------------
def pack_equations(eqs: list[int], cols: int):
    """eq bit 0 -> b, bit k (1..cols) -> column k-1; higher bits ignored; sign
    ignored (digits are read by magnitude, _internal.c:10,14,43-58)."""
    m = len(eqs)
    nw = (cols + 63) // 64
    A = np.zeros((m, max(nw, 1)), dtype=np.uint64)
    b = np.zeros(((m + 63) // 64 or 1,), dtype=np.uint64)
    mask = (1 << cols) - 1
    for i, e in enumerate(eqs):
        e = abs(int(e))
        if e & 1:
            b[i >> 6] |= np.uint64(1 << (i & 63))
        v = (e >> 1) & mask
        if v:
            A[i, :] = np.frombuffer(v.to_bytes(max(nw, 1) * 8, "little"), dtype="<u8")
    return A, b


def words_to_int(words: np.ndarray) -> int:
    return int.from_bytes(np.ascontiguousarray(words, dtype="<u8").tobytes(), "little")


# --------------------------------------------------------------------------
# m4ri_solve restated at the Python level (reference _internal.c:359-502)
# --------------------------------------------------------------------------
class OracleAffineSpace:
    """Mirror of _internal.AffineSpace (reference _internal.c:181-304)."""

    def __init__(self, origin: int, basis: tuple[int, ...]):
        self.origin = origin
        self.basis = basis

    @property
    def dimension(self) -> int:
        return len(self.basis)

    def get(self, i: int) -> int:
        # plain binary digits of i select basis rows (_internal.c:257-265)
        v = self.origin
        for j in range(self.dimension):
            if (i >> j) & 1:
                v ^= self.basis[j]
        return v

    def __iter__(self) -> Iterator[int]:
        d = self.dimension
        if d <= 64:
            # Gray-code walk (_internal.c:101-122)
            cur = self.origin
            idx = 0
            while True:
                yield cur
                x = idx ^ (idx >> 1)
                idx = (idx + 1) & ((1 << 64) - 1)
                y = idx ^ (idx >> 1)
                diff = ((x ^ y) & -(x ^ y)).bit_length() - 1 if (x ^ y) else 64
                if diff >= d or (d == 64 and idx == 0):
                    return
                cur ^= self.basis[diff]
        else:
            # little-endian binary counter (_internal.c:63-91)
            state = [0] * (d + 1)
            while not state[d]:
                v = self.origin
                for r in range(d):
                    if state[r]:
                        v ^= self.basis[r]
                sentinel = 1
                for r in range(d):
                    state[r] ^= 1
                    if state[r]:
                        sentinel = 0
                        break
                state[d] = sentinel
                yield v


def m4ri_solve(equations: list[int], cols: int, mode: int, tier: str = "m4rm"):
    """Restatement of reference m4ri_solve (_internal.c:359-502)."""
    if not isinstance(equations, list):
        raise TypeError("The first argument equations must be a list")
    if cols <= 0:
        raise ValueError("Number of columns must be positive")
    if mode not in (0, 1):
        raise ValueError("Invalid mode")
    if len(equations) < cols:
        raise ValueError("Number of rows must be greater than or equal to number of columns, try pad with zeros.")
    for e in equations:
        if not isinstance(e, int):
            raise TypeError("List items must be integers")
    A, b = pack_equations(equations, cols)
    sol = solve_packed(A, b, cols, mode, tier=tier)
    if sol.status == 1:
        return None
    origin = words_to_int(sol.origin)
    if mode == 0:
        return origin
    return OracleAffineSpace(origin, tuple(words_to_int(v) for v in sol.basis))


# --------------------------------------------------------------------------
# Independent pure-Python statement (big ints), small systems only
# --------------------------------------------------------------------------
def solve_bigint(equations: list[int], cols: int, mode: int):
    """Gauss-Jordan on Python ints: row int has bit 0 = constant, bit j+1 = a_j."""
    mask = (1 << (cols + 1)) - 1
    rows = [abs(e) & mask for e in equations]
    piv: list[int] = []
    r = 0
    for c in range(cols):
        bit = 1 << (c + 1)
        p = next((i for i in range(r, len(rows)) if rows[i] & bit), None)
        if p is None:
            continue
        rows[r], rows[p] = rows[p], rows[r]
        pr = rows[r]
        for i in range(len(rows)):
            if i != r and rows[i] & bit:
                rows[i] ^= pr
        piv.append(c)
        r += 1
    if any(rows[i] == 1 for i in range(r, len(rows))):
        return None
    origin = 0
    for j, c in enumerate(piv):
        if rows[j] & 1:
            origin |= 1 << c
    if mode == 0:
        return origin
    sigma = list(range(cols))
    for i, c in enumerate(piv):
        sigma[i], sigma[c] = sigma[c], sigma[i]
    basis = []
    for f in sigma[r:]:
        v = 1 << f
        for j, c in enumerate(piv):
            if (rows[j] >> (f + 1)) & 1:
                v |= 1 << c
        basis.append(v)
    return OracleAffineSpace(origin, tuple(basis))
