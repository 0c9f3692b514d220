"""Symbolic bit vectors over GF(2) -- the objects ``LinearSystem.gens()`` hands out.

API-compatible with the reference's ``gf2bv.BitVec`` (gf2bv/__init__.py:21-134):
a BitVec is a tuple of Python ints, one per bit, LSB first.  In each int, bit 0 is
the affine constant and bit ``k`` (k >= 1) is the coefficient of unknown ``k - 1``.
This layer only *constructs* systems; it runs before the solve path and stays on
the host.
"""
from __future__ import annotations

from functools import reduce
from operator import xor

from ._internal import to_bits, tuple_where, xor_tuple


def _is_constant(v: int) -> bool:
    return v == 0 or v == 1


class BitVec:
    __slots__ = ("_bits",)

    def __init__(self, bits: tuple[int, ...]):
        self._bits = bits  # symbolic bits, little-endian

    # -- container protocol -------------------------------------------------
    def __len__(self) -> int:
        return len(self._bits)

    def __getitem__(self, key):
        # a single index still yields a (1-bit) BitVec so bv[0] ^ bv is rejected
        picked = self._bits[key] if isinstance(key, slice) else (self._bits[key],)
        return BitVec(picked)

    # -- GF(2) addition -----------------------------------------------------
    def __xor__(self, other):
        if isinstance(other, BitVec):
            if len(other._bits) != len(self._bits):
                raise ValueError("Cannot mix bitvecs of different lengths")
            rhs = other._bits
        else:
            rhs = to_bits(len(self._bits), other)
        return BitVec(xor_tuple(self._bits, rhs))

    __rxor__ = __xor__
    __pow__ = __xor__  # `^` is exponentiation inside Sage's preparser

    # -- shifts / rotations (pure re-indexing) --------------------------------
    def __rshift__(self, n: int):
        return self if n == 0 else BitVec(self._bits[n:] + (0,) * n)

    def __lshift__(self, n: int):
        return self if n == 0 else BitVec((0,) * n + self._bits[:-n])

    def lshift_ext(self, n: int):
        return BitVec((0,) * n + self._bits)

    def rotr(self, n: int):
        return BitVec(self._bits[n:] + self._bits[:n])

    def rotl(self, n: int):
        return BitVec(self._bits[-n:] + self._bits[:-n])

    # -- masking with constants (linear) --------------------------------------
    def __and__(self, mask: int):
        keep = to_bits(len(self._bits), mask)
        if all(keep):
            return self
        return BitVec(tuple_where(keep, self._bits, 0))

    __rand__ = __and__

    def __or__(self, other):
        if not isinstance(other, BitVec):
            ones = to_bits(len(self._bits), other)
            if all(ones):
                return BitVec(ones)
            return BitVec(tuple_where(ones, 1, self._bits))
        short, long_ = (self._bits, other._bits) if len(self._bits) <= len(other._bits) else (other._bits, self._bits)
        merged = list(long_)
        for i, (p, q) in enumerate(zip(short, long_)):
            if not _is_constant(p) and not _is_constant(q):
                raise ValueError("Cannot compute logical or using bitvecs with non-zero bits")
            if p == 1 or q == 1:
                merged[i] = 1
            elif p == 0:
                merged[i] = q
            else:  # q == 0
                merged[i] = p
        return BitVec(tuple(merged))

    __ror__ = __or__

    def __mod__(self, n: int):
        if n & (n - 1):
            raise ValueError("modulo non-power-of-2 is not a linear operation")
        return self & (n - 1)

    # -- reshaping ------------------------------------------------------------
    def sum(self):
        return BitVec((reduce(xor, self._bits),))

    def zeroext(self, n: int):
        return BitVec(self._bits + (0,) * n)

    def signext(self, n: int):
        return BitVec(self._bits + (self._bits[-1],) * n)

    def broadcast(self, i: int, n: int):
        return BitVec((self._bits[i],) * n)

    def dup(self, n: int):
        return BitVec(self._bits * n)

    def concat(self, other: "BitVec"):
        return BitVec(self._bits + other._bits)

    # -- evaluation -----------------------------------------------------------
    def evaluate(self, s: int) -> int:
        """Value of this BitVec under the raw solution int ``s`` (bit j = x_j)."""
        point = (s << 1) | 1  # align with the equation encoding; constant term on
        value = 0
        for k, coeffs in enumerate(self._bits):
            if (coeffs & point).bit_count() & 1:
                value |= 1 << k
        return value
