"""The generic vector / operator protocol of spuq, as the drop-in boundary of the GPU path.

Mirrors (same names, argument meaning and error behaviour)
  * ``Basis``, ``CanonicalBasis``, ``BasisMismatchError``, ``check_basis``  spuq/linalg/basis.py:8-65
  * ``Vector``, ``FlatVector``, ``inner``                               spuq/linalg/vector.py:44-235
  * ``Operator``, ``BaseOperator``, ``ComposedOperator``, ``SummedOperator``, ``MatrixOperator``,
    ``DiagonalMatrixOperator``, ``MatrixSolveOperator``, ``MultiplicationOperator``
                                                                      spuq/linalg/operator.py:35-480
  * ``ScipyOperator``, ``ScipySolveOperator``                    spuq/linalg/scipy_operator.py:16-61

Division of labour.  ``FlatVector`` is the reference's host vector (a numpy array plus a basis); here
it is the *carrier* that moves per-mu coefficient arrays in and out of the device slab, and its
element-wise algebra stays numpy, exactly the reference's code path for host vectors.  The leaf
operators are what the ``assemble`` callbacks of ``MultiOperator`` return: they describe a spatial
block (dense array, diagonal, CSR matrix).  ``MultiOperator`` never calls their ``apply``; it reads
the block through ``as_csr()`` and runs everything on the device.  Their own ``apply`` (needed by
reference-style code such as test_pcg.py) goes through the same CUDA kernels as a one-column slab;
there is no numpy mat-vec in this module.
"""
from numbers import Number as Scalar

import numpy as np
import scipy.sparse as sps

__all__ = ["Scalar", "Basis", "CanonicalBasis", "BasisMismatchError", "check_basis", "Vector",
           "FlatVector", "inner", "Operator", "BaseOperator", "ComposedOperator", "SummedOperator",
           "MatrixOperator", "DiagonalMatrixOperator", "MatrixSolveOperator",
           "MultiplicationOperator", "ScipyOperator", "ScipySolveOperator"]


# ---- bases -----------------------------------------------------------------------------------
class BasisMismatchError(ValueError):
    pass


def check_basis(basis1, basis2, descr1="basis1", descr2="basis2"):
    if basis1 != basis2:
        raise BasisMismatchError("Basis don't match: %s=%s, %s=%s"
                                 % (descr1, str(basis1), descr2, str(basis2)))


class Basis:
    """Two bases are equal when they have the same class and the same attributes
    (``with_equality``, spuq/utils/__init__.py:16-29)."""

    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __ne__(self, other):
        return not self == other

    __hash__ = None

    @property
    def dim(self):  # pragma: no cover - abstract
        raise NotImplementedError

    def __repr__(self):
        return "<%s dim=%s>" % (type(self).__name__, self.dim)


class CanonicalBasis(Basis):
    def __init__(self, dim):
        self._dim = int(dim)

    @property
    def dim(self):
        return self._dim

    @property
    def gramian(self):
        return 1

    def copy(self):
        return type(self)(self._dim)


# ---- vectors ---------------------------------------------------------------------------------
class Vector:
    """Vector-space protocol: subclasses provide ``copy``, ``__iadd__``, ``__imul__`` and
    ``__isub__`` or ``__neg__``; the binary operators copy then update in place
    (vector.py:103-133)."""

    @property
    def dim(self):
        return self.basis.dim

    def copy(self):  # pragma: no cover - abstract
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

    def __rsub__(self, other):
        return NotImplemented

    def __mul__(self, other):
        if isinstance(other, Scalar):
            return self.copy().__imul__(other)
        return NotImplemented

    def __rmul__(self, other):
        if isinstance(other, Scalar):
            return self.copy().__imul__(other)
        return NotImplemented

    def __inner__(self, other):
        return NotImplemented

    def __rinner__(self, other):
        return NotImplemented

    def __repr__(self):
        return "<%s basis=%s>" % (type(self).__name__, self.basis)


class FlatVector(Vector):
    """numpy coefficient array + basis; vector.py:140-221."""

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
    def coeffs(self, value):
        self._coeffs = value

    @property
    def dim(self):
        return len(self._coeffs)

    def as_array(self):
        return self._coeffs

    def flatten(self):
        return self

    def copy(self):
        return self._create_copy(self._coeffs.copy())

    def _create_copy(self, coeffs):
        return type(self)(coeffs, self._basis)

    def set_zero(self):
        self._coeffs[:] = 0.0

    def __eq__(self, other):
        return (type(self) == type(other) and self._basis == other._basis
                and bool((self._coeffs == other._coeffs).all()))

    __hash__ = None

    def __neg__(self):
        return self._create_copy(-self._coeffs)

    def __iadd__(self, other):
        check_basis(self.basis, other.basis)
        self._coeffs += other.coeffs
        return self

    def __isub__(self, other):
        check_basis(self.basis, other.basis)
        self._coeffs -= other.coeffs
        return self

    def __imul__(self, other):
        if not isinstance(other, Scalar):
            raise TypeError("FlatVector can only be scaled by a scalar")
        self._coeffs *= other
        return self

    def __inner__(self, other):
        check_basis(self.basis, other.basis)
        return float(np.dot(self._coeffs, other.coeffs))

    def __rinner__(self, other):
        return self.__inner__(other)

    def __repr__(self):
        return "<%s basis=%s, coeffs=%s>" % (type(self).__name__, self.basis, self.coeffs)


def inner(vec1, vec2):
    """Scalar product through the ``__inner__`` / ``__rinner__`` protocol (vector.py:223-235)."""
    if not isinstance(vec1, Vector) or not isinstance(vec2, Vector):
        raise TypeError("inner() needs two Vectors")
    res = vec1.__inner__(vec2)
    if res is not NotImplemented:
        return res
    res = vec2.__rinner__(vec1)
    if res is not NotImplemented:
        return res
    raise NotImplementedError(str(vec1) + "/" + str(vec2))


# ---- operators -------------------------------------------------------------------------------
class Operator:
    """``A * x`` / ``A(x)`` apply the operator; ``A * B`` composes; ``c * A`` scales
    (operator.py:88-117)."""

    def apply(self, vec):  # pragma: no cover - abstract
        raise NotImplementedError

    @property
    def is_linear(self):
        return True

    def __mul__(self, other):
        if isinstance(other, Operator):
            return ComposedOperator(other, self)
        if isinstance(other, Scalar):
            return SummedOperator(operators=(self,), factors=(other,))
        if isinstance(other, Vector):
            return self.apply(other)
        raise TypeError("cannot multiply an Operator with %s" % type(other).__name__)

    def __rmul__(self, other):
        if not isinstance(other, Scalar):
            raise TypeError("only scalars multiply an Operator from the left")
        return self.__mul__(other)

    def __add__(self, other):
        return SummedOperator(operators=(self, other))

    def __sub__(self, other):
        return SummedOperator(operators=(self, other), factors=(1, -1))

    def __call__(self, arg):
        return self.apply(arg)

    def _check_basis(self, vec):
        if self.domain != vec.basis:
            raise BasisMismatchError("Basis don't match: domain %s vector %s"
                                     % (str(self.domain), str(vec.basis)))


class BaseOperator(Operator):
    def __init__(self, domain, codomain):
        self._domain = domain
        self._codomain = codomain

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
    """``op2 after op1`` (operator.py:160-240)."""

    def __init__(self, op1, op2):
        try:
            assert op1.codomain == op2.domain
        except AttributeError:
            pass
        self.op1, self.op2 = op1, op2

    domain = property(lambda self: self.op1.domain)
    codomain = property(lambda self: self.op2.codomain)

    def apply(self, vec):
        return self.op2.apply(self.op1.apply(vec))

    def as_matrix(self):
        return np.asarray(self.op2.as_matrix()) @ np.asarray(self.op1.as_matrix())


class SummedOperator(Operator):
    """sum_i factor_i * op_i (operator.py:243-330)."""

    def __init__(self, operators, factors=None):
        first = operators[0]
        for op in operators:
            assert first.domain == op.domain
            assert first.codomain == op.codomain
        self.operators, self.factors = operators, factors

    domain = property(lambda self: self.operators[0].domain)
    codomain = property(lambda self: self.operators[0].codomain)

    def apply(self, vec):
        r = None
        for i, op in enumerate(self.operators):
            r1 = op.apply(vec)
            if self.factors and self.factors[i] != 1.0:
                r1 = self.factors[i] * r1
            r = r1 if r is None else r.__iadd__(r1)
        return r

    def as_matrix(self):
        mats = [np.asarray(op.as_matrix()) for op in self.operators]
        facs = self.factors or [1.0] * len(mats)
        return sum(f * m for f, m in zip(facs, mats))


class _CsrLeaf(BaseOperator):
    """A leaf operator that is a fixed sparse matrix: apply = one CUDA SpMV on a one-column slab."""

    _plan = None

    def as_csr(self):  # pragma: no cover - abstract
        raise NotImplementedError

    def _device_plan(self):
        if self._plan is None:
            from .multi_operator import CsrSlabOperator   # late import: avoids a cycle
            self._plan = CsrSlabOperator(self.as_csr(), n_columns=1)
        return self._plan

    def _apply_on_device(self, vec, out_basis):
        y = self._device_plan().apply_array(np.ascontiguousarray(vec.coeffs, dtype=np.float64))
        return type(vec)(y, out_basis)


class MatrixOperator(_CsrLeaf):
    """Dense block ``arr`` (operator.py:339-385); ``apply`` = arr . x."""

    def __init__(self, arr, domain=None, codomain=None):
        if not isinstance(arr, np.ndarray):
            raise TypeError("MatrixOperator needs a numpy array")
        assert arr.ndim == 2
        if domain is None:
            domain = CanonicalBasis(arr.shape[1])
        elif domain.dim != arr.shape[1]:
            raise TypeError("size of domain basis does not match matrix dimensions")
        if codomain is None:
            codomain = CanonicalBasis(arr.shape[0])
        elif codomain.dim != arr.shape[0]:
            raise TypeError("size of domain basis does not match matrix dimensions")
        self._arr = np.asarray(arr)
        super().__init__(domain, codomain)

    @classmethod
    def from_sequence(cls, arr, domain=None, codomain=None):
        return cls(np.array(arr, dtype=float), domain, codomain)

    def apply(self, vec):
        self._check_basis(vec)
        return self._apply_on_device(vec, self.codomain)

    def as_matrix(self):
        return np.asarray(self._arr)

    def as_csr(self):
        return sps.csr_matrix(np.asarray(self._arr, dtype=np.float64))

    def transpose(self):
        return MatrixOperator(self._arr.T, self.codomain, self.domain)

    def __eq__(self, other):
        return (type(self) is type(other) and self.domain == other.domain
                and self.codomain == other.codomain and bool((self._arr == other._arr).all()))

    __hash__ = None


class DiagonalMatrixOperator(_CsrLeaf):
    """operator.py:388-418."""

    def __init__(self, diag, domain=None, codomain=None):
        assert isinstance(diag, np.ndarray) and diag.ndim == 1
        if domain is None:
            domain = CanonicalBasis(diag.shape[0])
        if codomain is None:
            codomain = domain
        assert domain.dim == codomain.dim
        self._diag = np.asarray(diag)
        super().__init__(domain, codomain)

    def apply(self, vec):
        self._check_basis(vec)
        return self._apply_on_device(vec, self.domain)

    def as_matrix(self):
        return np.diag(self._diag)

    def as_csr(self):
        return sps.diags(np.asarray(self._diag, dtype=np.float64), format="csr")

    def transpose(self):
        return DiagonalMatrixOperator(self._diag, self.codomain, self.domain)

    def inverse(self):
        return DiagonalMatrixOperator(1.0 / self._diag, self.codomain, self.domain)


class MatrixSolveOperator(_CsrLeaf):
    """``apply`` = arr^{-1} x (operator.py:421-462).  The reference solves with LAPACK on every
    application; here the inverse is formed once on the host at construction (set-up, like an
    assembly) and applied on the device as an explicit matrix."""

    def __init__(self, arr, domain=None, codomain=None):
        if not isinstance(arr, np.ndarray):
            arr = np.array(arr, dtype=float)
        assert arr.ndim == 2
        if domain is None:
            domain = CanonicalBasis(arr.shape[0])
        elif domain.dim != arr.shape[0]:
            raise TypeError("size of domain basis does not match matrix dimensions")
        if codomain is None:
            codomain = CanonicalBasis(arr.shape[1])
        elif codomain.dim != arr.shape[1]:
            raise TypeError("size of domain basis does not match matrix dimensions")
        self._arr = np.asarray(arr)
        self._inv = None
        super().__init__(domain, codomain)

    def apply(self, vec):
        self._check_basis(vec)
        return self._apply_on_device(vec, self.codomain)

    def as_matrix(self):
        if self._inv is None:
            self._inv = np.linalg.inv(self._arr)
        return self._inv

    def as_csr(self):
        return sps.csr_matrix(self.as_matrix())

    def transpose(self):
        return MatrixSolveOperator(self._arr.T, self.codomain, self.domain)


class MultiplicationOperator(BaseOperator):
    """Multiplication with a scalar (operator.py:465-480); works on any Vector."""

    def __init__(self, a, domain):
        self._a = a
        super().__init__(domain, domain)

    def apply(self, vec):
        return self._a * vec

    def transpose(self):
        return self

    def inverse(self):
        return MultiplicationOperator(1.0 / self._a, self.domain)


class ScipyOperator(_CsrLeaf):
    """CSR block (scipy_operator.py:16-40)."""

    def __init__(self, matrix, domain=None, codomain=None):
        if not sps.issparse(matrix):
            raise TypeError("ScipyOperator needs a scipy sparse matrix")
        if domain is None:
            domain = CanonicalBasis(matrix.shape[1])
        if codomain is None:
            codomain = CanonicalBasis(matrix.shape[0]) if matrix.shape[0] != matrix.shape[1] else domain
        self._matrix = matrix
        super().__init__(domain, codomain)

    @property
    def matrix(self):
        return self._matrix

    def apply(self, vec):
        return self._apply_on_device(vec, self.codomain)

    def as_matrix(self):
        return self._matrix

    def as_csr(self):
        return sps.csr_matrix(self._matrix, dtype=np.float64)


class ScipySolveOperator(BaseOperator):
    """Sparse direct solve block (scipy_operator.py:43-61): the callback result handed to
    ``PreconditioningOperator``.  It only *describes* the mean block; the GPU preconditioner solves it
    exactly on the device -- with the strip solver when it recognises a constant-coefficient 5-point
    block, through a sparse LU factorisation otherwise (multi_operator.mean_solver_for)."""

    def __init__(self, matrix, domain=None, codomain=None):
        if not sps.issparse(matrix):
            raise TypeError("ScipySolveOperator needs a scipy sparse matrix")
        if domain is None:
            domain = CanonicalBasis(matrix.shape[1])
        self._matrix = matrix
        super().__init__(domain, codomain or domain)

    @property
    def matrix(self):
        return self._matrix

    def as_csr(self):
        return sps.csr_matrix(self._matrix, dtype=np.float64)

    def apply(self, vec):
        from .multi_operator import mean_solver_for
        if getattr(self, "_solver", None) is None:
            self._solver = mean_solver_for(self.as_csr(), n_columns=1)
        y = self._solver.apply_array(np.ascontiguousarray(vec.coeffs, dtype=np.float64))
        return type(vec)(y, self.codomain)
