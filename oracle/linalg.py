"""Oracle: the generic vector / operator protocol and the numpy / scipy leaf operators.

Restates
  * bases            reference src/spuq/linalg/basis.py:8-65
  * Vector/FlatVector/inner   reference src/spuq/linalg/vector.py:44-235
  * Operator protocol + leaves reference src/spuq/linalg/operator.py:35-125, :339-480
  * CSR leaves       reference src/spuq/linalg/scipy_operator.py:16-61
Test infrastructure only (see oracle/__init__.py).
"""
from numbers import Number as Scalar

import numpy as np
import scipy.sparse.linalg as spsla


class BasisMismatchError(ValueError):
    pass


def check_basis(b1, b2, d1="basis1", d2="basis2"):
    if b1 != b2:
        raise BasisMismatchError("Basis don't match: %s=%s, %s=%s" % (d1, b1, d2, b2))


class Basis:
    """Equality = same class and same attributes (utils/__init__.py:16-29)."""

    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __ne__(self, other):
        return not self == other

    __hash__ = None

    def __repr__(self):
        return "<%s dim=%s>" % (type(self).__name__, self.dim)


class CanonicalBasis(Basis):
    def __init__(self, dim):
        self._dim = dim

    @property
    def dim(self):
        return self._dim

    @property
    def gramian(self):
        return 1


# ------------------------------------------------------------------------------------------
class Vector:
    """vector.py:44-137: everything is derived from copy / += / -= / *= by copy-then-in-place."""

    def copy(self):
        raise NotImplementedError

    def __add__(self, other):
        return self.copy().__iadd__(other)

    def __radd__(self, other):
        if isinstance(other, Scalar) and other == 0:
            return self
        return NotImplemented

    def __sub__(self, other):
        if hasattr(self, "__isub__"):
            return self.copy().__isub__(other)
        return self + (-other)

    def __mul__(self, other):
        if isinstance(other, Scalar):
            return self.copy().__imul__(other)
        return NotImplemented

    __rmul__ = __mul__

    def __inner__(self, other):
        return NotImplemented

    def __rinner__(self, other):
        return NotImplemented


class FlatVector(Vector):
    """vector.py:140-221."""

    def __init__(self, coeffs, basis=None):
        if not isinstance(coeffs, np.ndarray):
            coeffs = np.array(coeffs, dtype=float)
        if basis is None:
            basis = CanonicalBasis(coeffs.shape[0])
        assert basis.dim == coeffs.shape[0]
        assert coeffs.ndim == 1
        self._coeffs = coeffs
        self._basis = basis

    @property
    def basis(self):
        return self._basis

    @property
    def coeffs(self):
        return self._coeffs

    @coeffs.setter
    def coeffs(self, val):
        self._coeffs = val

    @property
    def dim(self):
        return len(self._coeffs)

    def as_array(self):
        return self._coeffs

    def flatten(self):
        return self

    def copy(self):
        return type(self)(self._coeffs.copy(), self._basis)

    def set_zero(self):
        self._coeffs[:] = 0

    def __eq__(self, other):
        return (type(self) is type(other) and self._basis == other._basis
                and bool((self._coeffs == other._coeffs).all()))

    __hash__ = None

    def __neg__(self):
        return type(self)(-self._coeffs, self._basis)

    def __iadd__(self, other):
        check_basis(self.basis, other.basis)
        self._coeffs += other._coeffs
        return self

    def __isub__(self, other):
        check_basis(self.basis, other.basis)
        self._coeffs -= other._coeffs
        return self

    def __imul__(self, other):
        self._coeffs *= other
        return self

    def __inner__(self, other):
        check_basis(self.basis, other.basis)
        return np.dot(self._coeffs, other._coeffs)

    __rinner__ = __inner__

    def __repr__(self):
        return "<%s basis=%s, coeffs=%s>" % (type(self).__name__, self.basis, self.coeffs)


def inner(v1, v2):
    """vector.py:223-235."""
    res = v1.__inner__(v2)
    if res is not NotImplemented:
        return res
    res = v2.__rinner__(v1)
    if res is not NotImplemented:
        return res
    raise NotImplementedError("%s/%s" % (v1, v2))


# ------------------------------------------------------------------------------------------
class Operator:
    """operator.py:35-125: ``A * x`` and ``A(x)`` both mean apply for a non-scalar, non-operator x."""

    def apply(self, vec):
        raise NotImplementedError

    def __mul__(self, other):
        if isinstance(other, Operator):
            return ComposedOperator(other, self)
        if isinstance(other, Scalar):
            return ScaledOperator(self, other)
        return self.apply(other)

    def __rmul__(self, other):
        return self.__mul__(other)

    def __call__(self, arg):
        return self.apply(arg)

    def _check_basis(self, vec):
        if self.domain != vec.basis:
            raise BasisMismatchError("Basis don't match: domain %s vector %s"
                                     % (self.domain, vec.basis))


class BaseOperator(Operator):
    def __init__(self, domain, codomain):
        self._domain, self._codomain = domain, codomain

    @property
    def domain(self):
        return self._domain

    @property
    def codomain(self):
        return self._codomain

    @property
    def dim(self):
        return self._domain.dim


class ComposedOperator(Operator):
    def __init__(self, first, second):
        self.op1, self.op2 = first, second

    domain = property(lambda self: self.op1.domain)
    codomain = property(lambda self: self.op2.codomain)

    def apply(self, vec):
        return self.op2.apply(self.op1.apply(vec))


class ScaledOperator(Operator):
    """``factor * A``: the one-term case of SummedOperator (operator.py:243-268)."""

    def __init__(self, op, factor):
        self.op, self.factor = op, factor

    domain = property(lambda self: self.op.domain)
    codomain = property(lambda self: self.op.codomain)

    def apply(self, vec):
        r = self.op.apply(vec)
        return r if self.factor == 1.0 else self.factor * r


class MatrixOperator(BaseOperator):
    """Dense block: apply = np.dot(arr, x); operator.py:339-385."""

    def __init__(self, arr, domain=None, codomain=None):
        arr = np.asarray(arr)
        assert arr.ndim == 2
        super().__init__(domain or CanonicalBasis(arr.shape[1]),
                         codomain or CanonicalBasis(arr.shape[0]))
        self._arr = arr

    def apply(self, vec):
        self._check_basis(vec)
        return type(vec)(np.dot(self._arr, vec.coeffs), self.codomain)

    def as_matrix(self):
        return np.asarray(self._arr)


class DiagonalMatrixOperator(BaseOperator):
    """operator.py:388-418."""

    def __init__(self, diag, domain=None, codomain=None):
        diag = np.asarray(diag)
        assert diag.ndim == 1
        domain = domain or CanonicalBasis(diag.shape[0])
        super().__init__(domain, codomain or domain)
        self._diag = diag

    def apply(self, vec):
        self._check_basis(vec)
        return type(vec)(np.multiply(self._diag, vec.coeffs), self.domain)

    def inverse(self):
        return DiagonalMatrixOperator(1.0 / self._diag, self.codomain, self.domain)

    def as_matrix(self):
        return np.diag(self._diag)


class MatrixSolveOperator(BaseOperator):
    """apply = np.linalg.solve(arr, x); operator.py:421-462."""

    def __init__(self, arr, domain=None, codomain=None):
        arr = np.asarray(arr, dtype=float)
        assert arr.ndim == 2
        super().__init__(domain or CanonicalBasis(arr.shape[0]),
                         codomain or CanonicalBasis(arr.shape[1]))
        self._arr = arr

    def apply(self, vec):
        self._check_basis(vec)
        return type(vec)(np.linalg.solve(self._arr, vec.coeffs), self.codomain)

    def as_matrix(self):
        return np.linalg.inv(self._arr)


class MultiplicationOperator(BaseOperator):
    """operator.py:465-480."""

    def __init__(self, a, domain):
        super().__init__(domain, domain)
        self._a = a

    def apply(self, vec):
        return self._a * vec


class ScipyOperator(BaseOperator):
    """CSR leaf: apply = matrix * coeffs (scipy csr_matvec); scipy_operator.py:27-40."""

    def __init__(self, matrix, domain=None, codomain=None):
        domain = domain or CanonicalBasis(matrix.shape[1])
        super().__init__(domain, codomain or domain)
        self._matrix = matrix

    @property
    def matrix(self):
        return self._matrix

    def apply(self, vec):
        out = vec.copy()
        out.coeffs = self._matrix @ out.coeffs
        return out

    def apply_to_matrix(self, X):
        return self._matrix @ X

    def as_matrix(self):
        return self._matrix


class ScipySolveOperator(BaseOperator):
    """Sparse direct solve per application; scipy_operator.py:43-61.

    The reference calls ``spsolve`` (factorise + solve) on every application.  ``factor=True``
    keeps one ``splu`` factorisation instead; it is the same SuperLU solve, only cached, and is
    what the timed CPU baseline uses (stated in bench.py's ``cpu_baseline.sample``).
    """

    def __init__(self, matrix, domain=None, codomain=None, factor=False):
        domain = domain or CanonicalBasis(matrix.shape[1])
        super().__init__(domain, codomain or domain)
        self._matrix = matrix.tocsc()
        self._lu = spsla.splu(self._matrix) if factor else None

    def apply(self, vec):
        out = vec.copy()
        if self._lu is not None:
            out.coeffs = self._lu.solve(vec.coeffs)
        else:
            out.coeffs = spsla.spsolve(self._matrix, vec.coeffs)
        return out
