"""ctypes binding of ``libgf2b200.so`` (C-ABI in ``include/gf2b200.h``).

This is plumbing, not the product: it only marshals numpy buffers / raw device
pointers into the C entry points.  There is no CPU fallback -- if the library is
missing or no CUDA device is present, everything here raises.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path
from typing import Optional

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("GF2B200_LIB") or (_HERE / "libgf2b200.so"))

OK, INCONSISTENT = 0, 1

EXPORTS = [
    "gf2b200_abi_version", "gf2b200_device_count", "gf2b200_create", "gf2b200_nccl_unique_id",
    "gf2b200_create_dist", "gf2b200_destroy", "gf2b200_last_error", "gf2b200_set_stream",
    "gf2b200_set_profile", "gf2b200_solve", "gf2b200_result_free", "gf2b200_system_create",
    "gf2b200_system_destroy", "gf2b200_system_local_rows", "gf2b200_system_load_host",
    "gf2b200_system_load_device", "gf2b200_system_generate", "gf2b200_system_eliminate",
    "gf2b200_system_result", "gf2b200_system_stats", "gf2b200_system_check_synthetic",
    "gf2b200_host_alloc", "gf2b200_host_free", "gf2b200_create_shards", "gf2b200_synth_host",
    "gf2b200_system_load_begin", "gf2b200_system_load_rows", "gf2b200_system_load_end",
    "gf2b200_solve_open", "gf2b200_solve_close",
]


class Gf2b200Error(RuntimeError):
    pass


class CResult(ctypes.Structure):
    _fields_ = [
        ("status", ctypes.c_int32),
        ("rank", ctypes.c_int64),
        ("kernel_dim", ctypes.c_int64),
        ("origin", ctypes.POINTER(ctypes.c_uint64)),
        ("basis", ctypes.POINTER(ctypes.c_uint64)),
        ("pivcols", ctypes.POINTER(ctypes.c_int64)),
    ]


class CStats(ctypes.Structure):
    _fields_ = [
        ("ms_total", ctypes.c_double),
        ("ms_forward", ctypes.c_double),
        ("ms_backward", ctypes.c_double),
        ("ms_sweep", ctypes.c_double),
        ("sweep_bytes", ctypes.c_double),
        ("exchange_bytes", ctypes.c_double),
        ("sweep_launches", ctypes.c_int64),
        ("kernel_launches", ctypes.c_int64),
        ("panels", ctypes.c_int64),
        ("rank", ctypes.c_int64),
        ("m_local", ctypes.c_int64),
        ("ms_sweep_max", ctypes.c_double),
        ("sweep_bytes_max", ctypes.c_double),
        ("sweep_bytes_timed", ctypes.c_double),
        ("sweep_launches_timed", ctypes.c_int64),
        ("forward_kernel_launches", ctypes.c_int64),
        ("ms_basis_solve", ctypes.c_double),
        ("ms_basis_output", ctypes.c_double),
        ("basis_sweep_bytes", ctypes.c_double),
        ("basis_panels", ctypes.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib() -> ctypes.CDLL:
    """Load libgf2b200.so (built by ``__graft_entry__.build()``); raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise Gf2b200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
            " -- gf2bv_b200 has no CPU fallback")
    L = ctypes.CDLL(str(LIB_PATH), mode=ctypes.RTLD_GLOBAL)
    if hasattr(L, "gf2b200_emulated_build") and os.environ.get("GF2B200_TEST_EMULATION") != "1":
        # the kernel tests' CPU emulation build is test infrastructure, never a way to run without a GPU
        raise Gf2b200Error(f"{LIB_PATH} is the CPU emulation build of the kernel tests; refusing to use it -- "
                           "gf2bv_b200 has no CPU fallback")
    vp, u64p, i64 = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64
    L.gf2b200_abi_version.restype = ctypes.c_int
    L.gf2b200_device_count.restype = ctypes.c_int
    L.gf2b200_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
    L.gf2b200_nccl_unique_id.argtypes = [vp]
    L.gf2b200_create_dist.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
    L.gf2b200_create_shards.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int]
    L.gf2b200_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
    L.gf2b200_host_free.argtypes = [vp]
    L.gf2b200_host_free.restype = None
    L.gf2b200_synth_host.argtypes = [u64p, u64p, i64, i64, i64, ctypes.c_uint64]
    L.gf2b200_destroy.argtypes = [vp]
    L.gf2b200_destroy.restype = None
    L.gf2b200_last_error.argtypes = [vp]
    L.gf2b200_last_error.restype = ctypes.c_char_p
    L.gf2b200_set_stream.argtypes = [vp, vp]
    L.gf2b200_set_profile.argtypes = [vp, ctypes.c_int]
    L.gf2b200_solve.argtypes = [vp, u64p, u64p, i64, i64, i64, ctypes.c_int, ctypes.POINTER(CResult)]
    L.gf2b200_result_free.argtypes = [ctypes.POINTER(CResult)]
    L.gf2b200_result_free.restype = None
    L.gf2b200_system_create.argtypes = [vp, i64, i64, ctypes.POINTER(vp)]
    L.gf2b200_system_destroy.argtypes = [vp]
    L.gf2b200_system_destroy.restype = None
    L.gf2b200_system_local_rows.argtypes = [vp]
    L.gf2b200_system_local_rows.restype = i64
    L.gf2b200_system_load_host.argtypes = [vp, u64p, u64p, i64]
    L.gf2b200_system_load_device.argtypes = [vp, u64p, u64p, i64]
    L.gf2b200_system_load_begin.argtypes = [vp, i64]
    L.gf2b200_system_load_rows.argtypes = [vp, u64p, i64, i64]
    L.gf2b200_system_load_end.argtypes = [vp, u64p]
    L.gf2b200_solve_open.argtypes = [vp, i64, i64, ctypes.POINTER(vp)]
    L.gf2b200_solve_close.argtypes = [vp, vp, ctypes.c_int, ctypes.POINTER(CResult)]
    L.gf2b200_system_generate.argtypes = [vp, ctypes.c_uint64]
    L.gf2b200_system_eliminate.argtypes = [vp]
    L.gf2b200_system_result.argtypes = [vp, ctypes.c_int, ctypes.POINTER(CResult)]
    L.gf2b200_system_stats.argtypes = [vp, ctypes.POINTER(CStats)]
    L.gf2b200_system_check_synthetic.argtypes = [vp, ctypes.c_uint64, u64p, ctypes.POINTER(i64)]
    _lib = L
    return L


class PackedSolution:
    """status 0/1, rank, origin words, basis rows (mode 1), pivot columns."""

    def __init__(self, status, rank, origin, basis, pivcols):
        self.status, self.rank, self.origin, self.basis, self.pivcols = status, rank, origin, basis, pivcols


def _take_result(res: CResult, nw: int, mode: int) -> PackedSolution:
    try:
        if res.status == INCONSISTENT:
            return PackedSolution(1, int(res.rank), None, None, None)
        rank = int(res.rank)
        origin = np.ctypeslib.as_array(res.origin, shape=(nw,)).copy()
        piv = np.ctypeslib.as_array(res.pivcols, shape=(max(rank, 1),)).copy()[:rank]
        basis = None
        if mode == 1:
            d = int(res.kernel_dim)
            basis = (np.ctypeslib.as_array(res.basis, shape=(d, nw)).copy() if d
                     else np.zeros((0, nw), dtype=np.uint64))
        return PackedSolution(0, rank, origin, basis, piv)
    finally:
        lib().gf2b200_result_free(ctypes.byref(res))


class Context:
    """One solver context (device + stream).  ``rank/world/nccl_id`` select the
    row-sharded multi-GPU mode (one process per GPU)."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, nccl_id: Optional[bytes] = None,
                 shards: int = 0):
        L = lib()
        self._h = ctypes.c_void_p()
        if shards:
            # loopback: `shards` row shards on this one device (tests the sharded path on 1 GPU)
            rc = L.gf2b200_create_shards(ctypes.byref(self._h), device, shards)
            world = shards
        elif world > 1:
            buf = ctypes.create_string_buffer(nccl_id, 128)
            rc = L.gf2b200_create_dist(ctypes.byref(self._h), device, rank, world, buf)
        else:
            rc = L.gf2b200_create(ctypes.byref(self._h), device)
        if rc:
            raise Gf2b200Error(f"gf2b200_create failed ({rc}): {L.gf2b200_last_error(None).decode()}")
        self.device, self.rank, self.world = device, rank, world

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        rc = lib().gf2b200_nccl_unique_id(buf)
        if rc:
            raise Gf2b200Error(f"gf2b200_nccl_unique_id failed ({rc}): {lib().gf2b200_last_error(None).decode()}")
        return buf.raw

    def _check(self, rc: int, what: str):
        if rc:
            raise Gf2b200Error(f"{what} failed ({rc}): {lib().gf2b200_last_error(self._h).decode()}")

    def set_stream(self, cuda_stream: int):
        self._check(lib().gf2b200_set_stream(self._h, ctypes.c_void_p(cuda_stream)), "set_stream")

    def set_profile(self, on: bool):
        self._check(lib().gf2b200_set_profile(self._h, int(on)), "set_profile")

    def solve(self, A: np.ndarray, b: Optional[np.ndarray], n: int, mode: int = 0) -> PackedSolution:
        """Host-buffer solve of the m x n system (A: uint64[m, stride], b: packed bits)."""
        assert A.dtype == np.uint64 and A.ndim == 2 and A.flags.c_contiguous
        m, stride = A.shape
        res = CResult()
        bp = None
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.uint64)
            bp = b.ctypes.data
        rc = lib().gf2b200_solve(self._h, A.ctypes.data, bp, m, n, stride, mode, ctypes.byref(res))
        self._check(rc, "gf2b200_solve")
        return _take_result(res, (n + 63) // 64, mode)

    def solve_ptr(self, A_ptr: int, b_ptr: Optional[int], m: int, n: int, stride: int, mode: int = 0):
        res = CResult()
        rc = lib().gf2b200_solve(self._h, A_ptr, b_ptr, m, n, stride, mode, ctypes.byref(res))
        self._check(rc, "gf2b200_solve")
        return _take_result(res, (n + 63) // 64, mode)

    def system(self, m: int, n: int) -> "System":
        return System(self, m, n)

    def close(self):
        if self._h:
            lib().gf2b200_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class System:
    """A device-resident m x n system (rows sharded over ranks in dist mode)."""

    def __init__(self, ctx: Context, m: int, n: int):
        self.ctx, self.m, self.n = ctx, m, n
        self.nw = (n + 63) // 64
        self._h = ctypes.c_void_p()
        ctx._check(lib().gf2b200_system_create(ctx._h, m, n, ctypes.byref(self._h)), "system_create")

    @property
    def local_rows(self) -> int:
        return int(lib().gf2b200_system_local_rows(self._h))

    def load_host(self, A: np.ndarray, b: Optional[np.ndarray]):
        assert A.dtype == np.uint64 and A.ndim == 2 and A.flags.c_contiguous
        bp = None
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.uint64)
            bp = b.ctypes.data
        self.ctx._check(lib().gf2b200_system_load_host(self._h, A.ctypes.data, bp, A.shape[1]), "load_host")

    def load_host_blocks(self, A: np.ndarray, b: Optional[np.ndarray], order):
        """the streaming form of load_host (gf2b200_system_load_begin / _rows / _end): `order` lists
        (row0, nrows) blocks of the local rows, handed over in that order"""
        A = np.ascontiguousarray(A, dtype=np.uint64)
        bp = None
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.uint64)
            bp = b.ctypes.data
        L = lib()
        self.ctx._check(L.gf2b200_system_load_begin(self._h, A.shape[1]), "load_begin")
        for row0, nrows in order:
            self.ctx._check(L.gf2b200_system_load_rows(self._h, A[row0:].ctypes.data, row0, nrows), "load_rows")
        self.ctx._check(L.gf2b200_system_load_end(self._h, bp), "load_end")

    def load_host_ptr(self, A_ptr: int, b_ptr: Optional[int], stride: int):
        self.ctx._check(lib().gf2b200_system_load_host(self._h, A_ptr, b_ptr, stride), "load_host")

    def load_device_ptr(self, dA: int, db: Optional[int], stride: int):
        self.ctx._check(lib().gf2b200_system_load_device(self._h, dA, db, stride), "load_device")

    def generate(self, seed: int = 1):
        self.ctx._check(lib().gf2b200_system_generate(self._h, seed), "generate")

    def eliminate(self):
        self.ctx._check(lib().gf2b200_system_eliminate(self._h), "eliminate")

    def result(self, mode: int = 0) -> PackedSolution:
        res = CResult()
        self.ctx._check(lib().gf2b200_system_result(self._h, mode, ctypes.byref(res)), "system_result")
        return _take_result(res, self.nw, mode)

    def stats(self) -> dict:
        st = CStats()
        self.ctx._check(lib().gf2b200_system_stats(self._h, ctypes.byref(st)), "system_stats")
        return st.as_dict()

    def check_synthetic(self, seed: int, x: np.ndarray) -> int:
        x = np.ascontiguousarray(x, dtype=np.uint64)
        bad = ctypes.c_int64(-1)
        self.ctx._check(lib().gf2b200_system_check_synthetic(self._h, seed, x.ctypes.data, ctypes.byref(bad)),
                        "check_synthetic")
        return int(bad.value)

    def close(self):
        if self._h:
            lib().gf2b200_system_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_host(A: np.ndarray, b: np.ndarray, row0: int, n: int, seed: int = 1):
    """Fill A (uint64[nrows, ceil(n/64)]) and b (packed bits) with rows row0.. of the
    synthetic system (host memory; the e2e measurement's input)."""
    assert A.dtype == np.uint64 and A.flags.c_contiguous and A.shape[1] == (n + 63) // 64
    assert b.dtype == np.uint64 and b.size >= (A.shape[0] + 63) // 64
    rc = lib().gf2b200_synth_host(A.ctypes.data, b.ctypes.data, row0, A.shape[0], n, seed)
    if rc:
        raise Gf2b200Error(f"gf2b200_synth_host failed ({rc})")


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    """Process-wide context on the device named by GF2B200_DEVICE (default 0)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("GF2B200_DEVICE", "0")))
    return _default_ctx
