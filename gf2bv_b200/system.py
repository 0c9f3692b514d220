"""``LinearSystem`` / ``QuadraticSystem``: the user-facing solve API.

Mirrors the reference's classes (gf2bv/__init__.py:146-408) -- same method names,
argument meaning, return values and error behaviour -- so a gf2bv user can write
``import gf2bv_b200 as gf2bv``.  The solve itself (``_solve_internal`` ->
``_internal.m4ri_solve``) runs on the B200 through libgf2b200.so; nothing here
falls back to a CPU solver.

Equation encoding (reference :151-159, _internal.c:411-425): an equation is a
Python int, bit 0 = constant term, bit k = coefficient of unknown k-1; ``zeros``
are expressions that must evaluate to 0.
"""
from __future__ import annotations

from collections.abc import Sequence
from typing import Optional

from ._internal import AffineSpace, eqs_to_sage_mat_helper, m4ri_solve, mul_bit_quad, to_bits
from .bitvec import BitVec

Zeros = Sequence  # of BitVec | int


class DimensionTooLargeError(Exception):
    """Raised by ``solve_all`` when the solution space has more than
    ``max_dimension`` dimensions; ``.space`` still gives access to it."""

    def __init__(self, message: str, space: AffineSpace):
        super().__init__(message)
        self.space = space


class LinearSystem:
    def __init__(self, sizes: list[int]):
        self._sizes = list(sizes)
        self._cols = sum(self._sizes)
        # _basis[0] is the constant term, _basis[k] is unknown k-1
        self._basis = [1 << i for i in range(self._cols + 1)]
        gens, at = [], 1
        for width in self._sizes:
            gens.append(BitVec(tuple(self._basis[at:at + width])))
            at += width
        self._vars = tuple(gens)

    def gens(self):
        return self._vars

    def __reduce__(self):
        return (self.__class__, (self._sizes,))

    # -- equations ------------------------------------------------------------
    def get_eqs(self, zeros: Zeros) -> list[int]:
        """Flatten ``zeros`` into equation ints, dropping literal 0 (= "0 == 0")."""
        flat: list[int] = []
        for z in zeros:
            if isinstance(z, BitVec):
                flat.extend(z._bits)
            else:
                flat.append(z)
        return [e for e in flat if e]

    def _solve_internal(self, zeros: Zeros, mode: int):
        eqs = self.get_eqs(zeros)
        if 1 in eqs:
            return None  # the literal equation 1 == 0
        short = self._cols - len(eqs)
        if short > 0:
            eqs.extend([0] * short)  # the extension wants rows >= cols
        return m4ri_solve(eqs, self._cols, mode)

    # -- solutions ------------------------------------------------------------
    def _convert_sol(self, s: int) -> tuple[int, ...]:
        parts = []
        for width in self._sizes:
            parts.append(s & ((1 << width) - 1))
            s >>= width
        assert s == 0, "Invalid solution"
        return tuple(parts)

    def convert_sol(self, s: int) -> Optional[tuple[int, ...]]:
        return self._convert_sol(s)

    def solve_raw_one(self, zeros: Zeros) -> Optional[int]:
        return self._solve_internal(zeros, 0)

    def solve_raw_space(self, zeros: Zeros) -> Optional[AffineSpace]:
        return self._solve_internal(zeros, 1)

    def solve_all(self, zeros: Zeros, *, max_dimension: int = 16):
        space = self.solve_raw_space(zeros)
        if space is None:
            return
        if space.dimension > max_dimension:
            raise DimensionTooLargeError(
                f"Solution space (dim {space.dimension}) is too large, try increase max_dimension "
                f"({max_dimension}) if you want (there will be 2**dim solutions)",
                space=space,
            )
        for raw in space:
            sol = self.convert_sol(raw)
            if sol is not None:
                yield sol

    def solve_one(self, zeros: Zeros):
        raw = self._solve_internal(zeros, 0)
        return None if raw is None else self.convert_sol(raw)

    def _join(self, sol: Sequence[int], sizes: Sequence[int]) -> int:
        s = 0
        for v, width in zip(reversed(sol), reversed(sizes)):
            s = (s << width) | v
        return s

    def evaluate(self, bv: BitVec, sol: tuple[int, ...]) -> int:
        """Value of ``bv`` under a solution tuple as returned by solve_one/solve_all."""
        return bv.evaluate(self._join(sol, self._sizes))

    # -- Sage interop (not on the solve path) -----------------------------------
    def get_sage_mat_slow(self, zeros: Zeros, *, tqdm=lambda x, desc: x):
        """(A, b) over GF(2) as Sage objects such that A x = b.  Needs SageMath."""
        from sage.all import GF, matrix, vector  # type: ignore

        eqs = self.get_eqs(zeros)
        rhs = vector(GF(2), [e & 1 for e in eqs])
        mat = matrix(GF(2), len(eqs), self._cols)
        for i, e in enumerate(tqdm(eqs, desc="Converting equations")):
            for j, bit in enumerate(to_bits(self._cols, e >> 1)):
                if bit:
                    mat[i, j] = 1
        return mat, rhs

    def get_sage_mat(self, zeros: Zeros):
        """The reference converts through libgd (`eqs_to_sage_mat_helper`, reference
        __init__.py:289-305); that bridge is out of scope here (the helper only raises
        RuntimeError, as the reference's does without libgd), so this is the slow conversion."""
        return self.get_sage_mat_slow(zeros)


class QuadraticSystem(LinearSystem):
    """Linearisation: every product x_i x_j (j < i) becomes one extra unknown."""

    def __init__(self, sizes: list[int]):
        n = sum(sizes)
        n_mono = n * (n - 1) // 2
        super().__init__(list(sizes) + [n_mono])
        self._quad_sizes = list(sizes)
        self._lin_size = n
        self._quad_size = n_mono
        self._const_lin_mask = (1 << (n + 1)) - 1

    def gens(self):
        return super().gens()[:-1]

    def __reduce__(self):
        return (self.__class__, (self._quad_sizes,))

    def _mul_bit(self, a: int, b: int) -> int:
        # x^2 = x over GF(2): constant and linear parts multiply bitwise
        low = (a & self._const_lin_mask) & b
        return mul_bit_quad(self._lin_size, a >> 1, b >> 1, low, self._basis)

    def _mul_bit_slow(self, a: int, b: int) -> int:
        n = self._lin_size
        out = (a & self._const_lin_mask) & b
        av, bv = to_bits(n, a >> 1), to_bits(n, b >> 1)
        mono = n + 1
        for i in range(n):
            for j in range(i):
                if (av[i] & bv[j]) ^ (av[j] & bv[i]):
                    out |= self._basis[mono]
                mono += 1
        return out

    def mul_bit(self, a: BitVec, b: BitVec) -> BitVec:
        if len(a) != 1 or len(b) != 1:
            raise ValueError("The inputs should be single bits")
        return BitVec((self._mul_bit(a._bits[0], b._bits[0]),))

    def _bit_assert(self, a: int, v: int):
        assert v in (0, 1), "Invalid bit"
        assert a not in (0, 1), "a should not be a constant"
        assert a >> self._lin_size == 0, "Not a linear term"
        zeros = [a ^ v]
        # a == v implies a * x == v * x for every unknown x
        for k in range(1, self._lin_size + 1):
            x = self._basis[k]
            if x == a:
                continue
            prod = self._mul_bit(a, x)
            zeros.append(prod ^ x if v else prod)
        return zeros

    def bit_assert(self, a: BitVec, v: int):
        if len(a) != 1:
            raise ValueError("The input should be a single bit")
        return self._bit_assert(a._bits[0], v)

    def _check_lin_match_quad(self, lin: int, quad: int) -> bool:
        n = self._lin_size
        assert lin >> n == 0, "Invalid linear part"
        for i in range(n):
            xi = (lin >> i) & 1
            for j in range(i):
                if (xi & (lin >> j) & 1) != (quad & 1):
                    return False
                quad >>= 1
        assert quad == 0, "Invalid quadratic part"
        return True

    def convert_sol(self, s: int) -> Optional[tuple[int, ...]]:
        lin = s & ((1 << self._lin_size) - 1)
        s >>= self._lin_size
        quad = s & ((1 << self._quad_size) - 1)
        assert s >> self._quad_size == 0, "Invalid solution"
        if not self._check_lin_match_quad(lin, quad):
            return None
        return super()._convert_sol(lin)[:-1]

    def solve_one(self, zeros: Zeros):
        # the first raw solution may fail the monomial consistency filter
        for sol in self.solve_all(zeros):
            return sol
        return None

    def evaluate(self, bv: BitVec, sol: tuple[int, ...]) -> int:
        return bv.evaluate(self._join(sol, self._quad_sizes))
