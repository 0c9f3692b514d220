"""gf2bv_b200 -- B200-native GF(2) linear-system solver behind gf2bv's LinearSystem API.

``import gf2bv_b200 as gf2bv`` gives the reference's public names
(gf2bv/__init__.py): ``BitVec``, ``LinearSystem``, ``QuadraticSystem``,
``DimensionTooLargeError``; ``gf2bv_b200._internal`` is the drop-in for the
reference's C extension (``m4ri_solve``, ``AffineSpace``, ...), whose solver is
``libgf2b200.so`` (hand-written sm_100a CUDA, C-ABI in ``include/gf2b200.h``).
There is no CPU fallback: solving without the built library or without a CUDA
device raises ``RuntimeError``.
"""
try:
    from . import _internal
except ImportError as exc:  # pragma: no cover - build step missing
    raise ImportError(
        "gf2bv_b200._internal is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
        "in the repository root (there is no pure-Python fallback)") from exc

from ._internal import AffineSpace
from .bitvec import BitVec
from .system import DimensionTooLargeError, LinearSystem, QuadraticSystem

__all__ = ["AffineSpace", "BitVec", "DimensionTooLargeError", "LinearSystem", "QuadraticSystem"]
