"""The stochastic-Galerkin operator and the mean-based preconditioner on the GPU.

Mirrors the reference classes
  * ``MultiOperator(coeff_field, assemble_0, assemble_m=None, domain=None, codomain=None)`` and
    ``.apply(w)``          spuq/application/egsz/multi_operator2.py:26-84
  * ``PreconditioningOperator(mean_func, assemble_solver, domain=None, codomain=None)`` and
    ``.apply(w)``          spuq/application/egsz/multi_operator2.py:139-165
behind the same constructor arguments, the same ``Operator`` protocol (``A * w``, ``A(w)``) and the
same error behaviour (TypeError for a non-MultiVector argument), so they are drop-ins for
``apply()`` and for the ``pcg`` solver that wraps it.

What changes underneath: the assemble callbacks are invoked ONCE per expansion term (the reference
calls them for every (mu, m)), their blocks are put on one shared CSR pattern in HBM, the index set
is turned into neighbour / beta tables (``LambdaTables``), and ``apply`` is a single launch of the
sm_100a kernel of libspuq_sg over the column-major slab.  The result is a new vector; the argument
is never modified (multi_operator2.py:42).
"""
import ctypes as C
import logging

import numpy as np
import scipy.sparse as sps
import torch

from . import _abi
from .coefficient_field import CoefficientField
from .lambda_tables import LambdaTables
from .linalg import Basis, Operator
from .multi_vector import MultiVector, SlabMultiVector

logger = logging.getLogger(__name__)

__all__ = ["MultiOperator", "PreconditioningOperator", "DeviceBlocks", "CsrSlabOperator",
           "MeanSolveSlabOperator", "SparseLUSlabOperator", "mean_solver_for", "SGPlan"]


# ---- spatial blocks on the device ----------------------------------------------------------------
def _block_to_csr(op):
    """CSR matrix of a leaf operator returned by an assemble callback."""
    if hasattr(op, "as_csr"):
        return op.as_csr()
    for attr in ("matrix", "_matrix"):
        m = getattr(op, attr, None)
        if m is not None and sps.issparse(m):
            return sps.csr_matrix(m, dtype=np.float64)
    if hasattr(op, "_diag"):
        return sps.diags(np.asarray(op._diag, dtype=np.float64), format="csr")
    if hasattr(op, "_arr"):
        return sps.csr_matrix(np.asarray(op._arr, dtype=np.float64))
    if hasattr(op, "as_matrix"):
        m = op.as_matrix()
        return sps.csr_matrix(m if sps.issparse(m) else np.asarray(m, dtype=np.float64))
    raise TypeError("cannot read a matrix out of %r" % (op,))


class DeviceBlocks:
    """A_0 and A_1..A_M as value arrays on ONE shared CSR pattern in device memory."""

    def __init__(self, rowptr, colidx, vals, shape):
        self.rowptr, self.colidx, self.vals, self.shape = rowptr, colidx, vals, shape

    @property
    def M(self):
        return self.vals.shape[0] - 1

    @property
    def nnz(self):
        return int(self.colidx.numel())

    @classmethod
    def from_scipy(cls, mats, device=None):
        """Union pattern of the given CSR matrices (explicit zeros where a block has no entry)."""
        device = device or _abi.device()
        mats = [sps.csr_matrix(m, dtype=np.float64) for m in mats]
        shape = mats[0].shape
        for m in mats:
            assert m.shape == shape, "all spatial blocks must have the same shape"
            m.sum_duplicates()
            m.sort_indices()
        same = all(m.nnz == mats[0].nnz and np.array_equal(m.indptr, mats[0].indptr)
                   and np.array_equal(m.indices, mats[0].indices) for m in mats[1:])
        if same:
            indptr, indices = mats[0].indptr, mats[0].indices
            vals = np.stack([m.data for m in mats])
        else:
            pat = sps.csr_matrix(shape, dtype=np.float64)
            for m in mats:
                one = m.copy()
                one.data = np.ones_like(one.data)
                pat = pat + one
            pat.sort_indices()
            indptr, indices = pat.indptr, pat.indices
            rows = np.repeat(np.arange(shape[0]), np.diff(indptr))
            key = rows.astype(np.int64) * shape[1] + indices
            vals = np.zeros((len(mats), len(indices)))
            for k, m in enumerate(mats):
                mrows = np.repeat(np.arange(shape[0]), np.diff(m.indptr))
                mkey = mrows.astype(np.int64) * shape[1] + m.indices
                vals[k, np.searchsorted(key, mkey)] = m.data
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device)  # noqa: E731
        return cls(t(indptr, np.int32), t(indices, np.int32), t(vals, np.float64), shape)


class SGPlan:
    """Owner of one ``spuq_sg_plan`` handle plus the device arrays it borrows."""

    def __init__(self, blocks, L, tables=None, m_range=None, include_mean=True, path=_abi.PATH_AUTO):
        lib = _abi.lib()
        self.blocks = blocks
        self.L = int(L)
        M = 0 if tables is None else tables.maxm
        assert blocks.M >= M, "not enough spatial blocks for the expansion length"
        self.M = M
        m_begin, m_end = (0, M) if m_range is None else m_range
        ptrs = (C.c_void_p * (M + 1))()
        for k in range(M + 1):
            ptrs[k] = blocks.vals[k].data_ptr()
        handle = C.c_void_p()
        if M:
            keep = [np.ascontiguousarray(a) for a in (tables.nbr_up, tables.nbr_dn, tables.beta_up,
                                                      tables.beta_dn, tables.beta_0)]
            args = [_abi.ptr(a) for a in keep]
        else:
            args = [None] * 5
        dev_index = _abi.device_index()
        _abi.check(lib.spuq_sg_plan_create(C.byref(handle), dev_index, blocks.shape[0], blocks.shape[1],
                                           self.L, M, _abi.ptr(blocks.rowptr), _abi.ptr(blocks.colidx),
                                           C.cast(ptrs, C.c_void_p), *args, m_begin, m_end,
                                           1 if include_mean else 0, path), "plan_create")
        self.handle = handle
        self.n_rows, self.n_cols = blocks.shape

    def info(self):
        path, nterms, nnz = C.c_int32(), C.c_int64(), C.c_int64()
        ab, af = C.c_int64(), C.c_int64()
        _abi.check(_abi.lib().spuq_sg_plan_info(self.handle, C.byref(path), C.byref(nterms), C.byref(nnz),
                                                C.byref(ab), C.byref(af)), "plan_info")
        return {"path": _abi.PATH_NAMES.get(path.value, str(path.value)), "n_terms": nterms.value,
                "nnz": nnz.value, "algorithmic_bytes": ab.value, "algorithmic_flops": af.value}

    def tune(self, variant=0, cols_per_block=0, rows_per_thread=0, ctas_per_sm=0):
        _abi.check(_abi.lib().spuq_sg_plan_tune(self.handle, variant, cols_per_block, rows_per_thread,
                                                ctas_per_sm), "plan_tune")

    def apply(self, W, Y):
        """Y = A W on (L, n) contiguous fp64 device tensors (column-major n x L slabs)."""
        assert W.dtype == torch.float64 and Y.dtype == torch.float64
        assert W.is_contiguous() and Y.is_contiguous()
        assert tuple(W.shape) == (self.L, self.n_cols) and tuple(Y.shape) == (self.L, self.n_rows)
        _abi.check(_abi.lib().spuq_sg_apply(self.handle, _abi.ptr(W), self.n_cols, _abi.ptr(Y), self.n_rows,
                                            _abi.stream_ptr()), "apply")

    def apply_part(self, W, Y, part):
        """Part of the output grid rows (1: tile rows that read no halo row, 2: the rest, 0: all); see
        ``spuq_sg_apply_part``."""
        assert W.is_contiguous() and Y.is_contiguous()
        assert tuple(W.shape) == (self.L, self.n_cols) and tuple(Y.shape) == (self.L, self.n_rows)
        _abi.check(_abi.lib().spuq_sg_apply_part(self.handle, _abi.ptr(W), self.n_cols, _abi.ptr(Y), self.n_rows,
                                                 int(part), _abi.stream_ptr()), "apply_part")

    @property
    def kernel_name(self):
        return _abi.lib().spuq_sg_plan_kernel_name(self.handle).decode()

    @property
    def last_launches(self):
        return int(_abi.lib().spuq_sg_plan_last_launches(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            _abi.lib().spuq_sg_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CsrSlabOperator:
    """One sparse matrix applied to every column of a slab: an SG plan with M = 0.  Backs the leaf
    operators of ``linalg`` and explicit-matrix preconditioners."""

    def __init__(self, csr, n_columns=1):
        self.blocks = DeviceBlocks.from_scipy([csr])
        self.plan = SGPlan(self.blocks, n_columns)

    def apply_slab(self, W, Y):
        self.plan.apply(W, Y)

    def apply_array(self, x):
        dev = _abi.device()
        W = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).reshape(1, -1)).to(dev)
        Y = torch.empty((1, self.plan.n_rows), dtype=torch.float64, device=dev)
        self.plan.apply(W, Y)
        return Y.cpu().numpy().reshape(-1)


class MeanSolveSlabOperator:
    """Exact solve with the mean block on every column (K4): recognises a constant-coefficient
    5-point block on an nx x ny grid and uses the fast-diagonalisation solver of libspuq_sg."""

    def __init__(self, csr, n_columns=1, stencil=None):
        """``stencil`` = (nx, ny, d, ox, oy) skips the recognition (and the need for the matrix)."""
        if stencil is None:
            csr = sps.csr_matrix(csr, dtype=np.float64)
            self.N = csr.shape[0]
            nx, ny, d, ox, oy = self._recognise(csr)
        else:
            nx, ny, d, ox, oy = stencil
            self.N = int(nx) * int(ny)
        self.stencil = (int(nx), int(ny), float(d), float(ox), float(oy))
        self.L = int(n_columns)
        handle = C.c_void_p()
        _abi.check(_abi.lib().spuq_sg_precond_mean_fd(C.byref(handle), _abi.device_index(), nx, ny, d, ox, oy),
                   "precond_mean_fd")
        self.handle = handle
        self._work = None

    @staticmethod
    def _recognise(csr):
        N = csr.shape[0]
        diag = csr.diagonal()
        d = float(diag[0])
        if not np.all(diag == d):
            raise _abi.SpuqError("mean block is not constant-coefficient (diagonal varies)")
        coo = csr.tocoo()
        coo.eliminate_zeros()                 # explicitly stored zeros are not couplings
        off = coo.col - coo.row
        far = np.abs(off[np.abs(off) > 1])
        nx = int(far.min()) if far.size else N
        ny = N // nx
        if nx * ny != N:
            raise _abi.SpuqError("mean block is not a tensor grid")
        ox_vals = coo.data[np.abs(off) == 1]
        oy_vals = coo.data[np.abs(off) == nx] if ny > 1 else np.zeros(0)
        ox = float(ox_vals[0]) if ox_vals.size else 0.0
        oy = float(oy_vals[0]) if oy_vals.size else 0.0
        ref = sps.kron(sps.identity(ny), sps.diags([ox, d / 2 if ny > 1 else d, ox], [-1, 0, 1], shape=(nx, nx)))
        if ny > 1:
            ref = ref + sps.kron(sps.diags([oy, d / 2, oy], [-1, 0, 1], shape=(ny, ny)), sps.identity(nx))
        if abs(ref - csr).max() != 0.0:
            raise _abi.SpuqError("mean block is not a constant-coefficient 5-point stencil; "
                                 "the exact GPU preconditioner does not apply")
        return nx, ny, d, ox, oy

    def set_columns(self, L):
        self.L = int(L)

    def apply_slab(self, R, S):
        lib = _abi.lib()
        L, N = R.shape
        if not (R.is_contiguous() and S.is_contiguous() and S.shape == R.shape):
            raise ValueError("MeanSolveSlabOperator.apply_slab needs contiguous L x N slabs of equal shape")
        need = int(lib.spuq_sg_precond_work_bytes(self.handle, N, L))
        if self._work is None or self._work.numel() * 8 < need:
            self._work = torch.empty(need // 8 + 1, dtype=torch.float64, device=R.device)
        _abi.check(lib.spuq_sg_precond_apply(self.handle, N, L, _abi.ptr(R), _abi.ptr(S), _abi.ptr(self._work),
                                             _abi.stream_ptr()), "precond_apply")

    def apply_array(self, x):
        dev = _abi.device()
        R = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).reshape(1, -1)).to(dev)
        S = torch.empty_like(R)
        self.apply_slab(R, S)
        return S.cpu().numpy().reshape(-1)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _abi.lib().spuq_sg_precond_destroy(self.handle)
        except Exception:
            pass


class SparseLUSlabOperator:
    """Exact solve with ANY sparse mean block on every column (K4, general form): the matrix is
    factorised once on the host with SuperLU (``scipy.sparse.linalg.splu``, the solver behind the
    reference's ``ScipySolveOperator.apply``, linalg/scipy_operator.py:45-49) and the triangular solves
    run on the device for the whole slab (csrc/precond_lu.cu).  Same interface as
    ``MeanSolveSlabOperator``."""

    def __init__(self, csr, n_columns=1):
        import scipy.sparse.linalg as spla
        csr = sps.csr_matrix(csr, dtype=np.float64)
        if csr.shape[0] != csr.shape[1]:
            raise ValueError("the mean block must be square")
        self.N = n = csr.shape[0]
        self.L = int(n_columns)
        lu = spla.splu(csr.tocsc())
        Lr = sps.csr_matrix(lu.L)                      # unit lower triangular
        Ur = sps.csr_matrix(lu.U)
        Lr.sort_indices(); Ur.sort_indices()
        Ls = sps.tril(Lr, k=-1, format="csr")
        Us = sps.triu(Ur, k=1, format="csr")
        diag = np.ascontiguousarray(Ur.diagonal(), dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)        # noqa: E731
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)      # noqa: E731
        keep = [i32(Ls.indptr), i32(Ls.indices), f64(Ls.data), i32(Us.indptr), i32(Us.indices), f64(Us.data),
                diag, i32(lu.perm_r), i32(lu.perm_c)]
        self.nnz_factors = int(Ls.nnz + Us.nnz + n)
        handle = C.c_void_p()
        _abi.check(_abi.lib().spuq_sg_precond_sparse_lu(C.byref(handle), _abi.device_index(), n,
                                                        *[_abi.ptr(a) for a in keep]), "precond_sparse_lu")
        self.handle = handle
        self._work = None

    set_columns = MeanSolveSlabOperator.set_columns
    apply_slab = MeanSolveSlabOperator.apply_slab
    apply_array = MeanSolveSlabOperator.apply_array
    __del__ = MeanSolveSlabOperator.__del__


def mean_solver_for(csr, n_columns=1):
    """The device solver for a mean block: the fast strip solver when the block is a constant-coefficient
    5-point stencil on a tensor grid, the sparse-LU solver for everything else (variable coefficients,
    permuted or unstructured numbering) -- the reference accepts any matrix here."""
    try:
        return MeanSolveSlabOperator(csr, n_columns)
    except _abi.SpuqError as e:
        if "mean block is not" not in str(e):
            raise
        return SparseLUSlabOperator(csr, n_columns)


# ---- the SG operator -------------------------------------------------------------------------------
class MultiOperator(Operator):
    """Discrete SG operator on one shared spatial basis (EGSZ (2.6) generalised to spuq's
    orthonormal polynomials), applied by the B200 kernel."""

    PLAN_CACHE = 2          # kernel plans kept alive (least recently used index sets are released)

    def __init__(self, coeff_field, assemble_0, assemble_m=None, domain=None, codomain=None,
                 m_range=None, include_mean=True, path=_abi.PATH_AUTO):
        if not isinstance(coeff_field, CoefficientField):
            raise TypeError("coeff_field must be a CoefficientField")
        if not callable(assemble_0) or not (assemble_m is None or callable(assemble_m)):
            raise TypeError("assemble_0 / assemble_m must be callable")
        for b in (domain, codomain):
            if b is not None and not isinstance(b, Basis):
                raise TypeError("domain / codomain must be a Basis")
        self._assemble_0 = assemble_0
        self._assemble_m = assemble_m or assemble_0
        self._coeff_field = coeff_field
        self._domain, self._codomain = domain, codomain
        self._m_range, self._include_mean, self._path = m_range, include_mean, path
        self._blocks = None          # (basis, maxm, DeviceBlocks)
        self._plans = {}             # Lambda key -> SGPlan
        self._preset_blocks = None
        self._rvs = None

    @classmethod
    def from_blocks(cls, blocks, rvs, domain=None, codomain=None, m_range=None, include_mean=True,
                    path=_abi.PATH_AUTO):
        """Operator over blocks that already live on the device (``DeviceBlocks``); ``rvs[m]`` is
        the random variable of expansion term m.  Used for large synthetic problems where building
        scipy matrices on the host first would only add a copy."""
        self = cls.__new__(cls)
        self._assemble_0 = self._assemble_m = None
        self._coeff_field = None
        self._domain, self._codomain = domain, codomain
        self._m_range, self._include_mean, self._path = m_range, include_mean, path
        self._blocks = None
        self._plans = {}
        self._preset_blocks = blocks
        self._rvs = list(rvs)
        return self

    @property
    def domain(self):
        return self._domain

    @property
    def codomain(self):
        return self._codomain

    # -- set-up --------------------------------------------------------------------------------
    def _n_terms_available(self):
        return len(self._rvs) if self._coeff_field is None else len(self._coeff_field)

    def _rv(self, m):
        return self._rvs[m] if self._coeff_field is None else self._coeff_field[m][1]

    def _device_blocks(self, basis, maxm):
        if self._preset_blocks is not None:
            return self._preset_blocks
        if self._blocks is not None and self._blocks[0] == basis and self._blocks[1] >= maxm:
            return self._blocks[2]
        ops = [self._assemble_0(basis, self._coeff_field.mean_func)]
        for m in range(maxm):
            ops.append(self._assemble_m(basis, self._coeff_field[m][0]))
        blocks = DeviceBlocks.from_scipy([_block_to_csr(op) for op in ops])
        self._blocks = (basis, maxm, blocks)
        self._plans.clear()
        return blocks

    def plan_for(self, w):
        """The (cached) kernel plan for the index set and basis of the slab vector ``w``."""
        maxm = w.max_order                                           # multi_operator2.py:44
        avail = self._n_terms_available()
        if avail < maxm:                                             # :45-48
            logger.warning("insufficient length of coefficient field for MultiVector (%i instead of %i",
                           avail, maxm)
            maxm = avail
        key = (w.index_array.shape, w.index_array.tobytes(), maxm, w.n)
        plan = self._plans.get(key)
        if plan is not None:
            self._plans[key] = self._plans.pop(key)                 # most recently used last
        if plan is None:
            blocks = self._device_blocks(w.spatial_basis, maxm)
            assert blocks.shape == (w.n, w.n), "spatial blocks do not match the vector dimension"
            rvs = [self._rv(m) for m in range(maxm)]
            tables = None
            if self._plans:                                      # Lambda usually GROWS between solves
                last = next(reversed(self._plans.values()))
                if getattr(last, "tables", None) is not None and last.tables.L < len(w):
                    tables = last.tables.grown(w.index_array, rvs, maxm)
            if tables is None:
                tables = LambdaTables(w.index_array, rvs, maxm)
            m_range = self._m_range
            if m_range is not None:
                m_range = (min(m_range[0], tables.maxm), min(m_range[1], tables.maxm))
            plan = SGPlan(blocks, len(w), tables, m_range, self._include_mean, self._path)
            plan.tables = tables
            self._plans[key] = plan
            # a plan owns a private re-layout of all blocks (about 2 GB on config 3) and its schedule: an
            # adaptive loop that grows Lambda every iteration must not keep one per index set.  Keep the
            # last PLAN_CACHE sets (current + previous), release the rest now rather than at collection time.
            while len(self._plans) > self.PLAN_CACHE:
                old_key = next(iter(self._plans))
                self._plans.pop(old_key).close()
        return plan

    # -- application ---------------------------------------------------------------------------
    def apply(self, w):
        if not isinstance(w, MultiVector):
            raise TypeError("MultiOperator.apply needs a MultiVector, got %s" % type(w).__name__)
        slab_in = w if isinstance(w, SlabMultiVector) else SlabMultiVector.from_multivector(w)
        plan = self.plan_for(slab_in)
        out = slab_in.like(torch.empty_like(slab_in.slab))
        plan.apply(slab_in.slab, out.slab)
        if isinstance(w, SlabMultiVector):
            return out
        return out.to_multivector(type(w))


class PreconditioningOperator(Operator):
    """Mean-based block-diagonal preconditioner  s[mu] = A_0^{-1} rho[mu]  (EGSZ section 7.1)."""

    def __init__(self, mean_func, assemble_solver, domain=None, codomain=None):
        if not callable(assemble_solver):
            raise TypeError("assemble_solver must be callable")
        self._assemble_solver = assemble_solver
        self._mean_func = mean_func
        self._domain, self._codomain = domain, codomain
        self._solver = None          # (basis, slab operator)

    @property
    def domain(self):
        return self._domain

    @property
    def codomain(self):
        return self._codomain

    def slab_operator(self, w):
        """Device solver for the basis of ``w`` (built once; the reference re-assembles and
        re-factorises per column and per iteration, multi_operator2.py:159)."""
        basis = w.spatial_basis
        if self._solver is None or self._solver[0] != basis:
            # the reference computes `A0 * w[mu]` with whatever operator the callback returns
            # (multi_operator2.py:159): dispatch on the operator TYPE -- a solve-type leaf becomes a device
            # solver, a multiply-type leaf (ScipyOperator, MatrixOperator, DiagonalMatrixOperator, and
            # MatrixSolveOperator through its explicit inverse) is applied as the matrix it is
            from .linalg import ScipySolveOperator, _CsrLeaf
            op = self._assemble_solver(basis, self._mean_func)
            if hasattr(op, "slab_solver"):
                solver = op.slab_solver()
            elif isinstance(op, ScipySolveOperator):
                solver = mean_solver_for(op.as_csr(), len(w))
            elif isinstance(op, _CsrLeaf):
                solver = CsrSlabOperator(op.as_csr(), len(w))
            elif hasattr(op, "as_matrix"):
                # any other operator that can state the matrix of its `apply` (operator.py:110-125)
                m = op.as_matrix()
                solver = CsrSlabOperator(sps.csr_matrix(m if sps.issparse(m) else np.asarray(m, dtype=np.float64)), len(w))
            else:
                raise TypeError("PreconditioningOperator: the assemble_solver callback returned %s, which has "
                                "no device form (supported: ScipySolveOperator, MatrixSolveOperator, "
                                "ScipyOperator, MatrixOperator, DiagonalMatrixOperator)" % type(op).__name__)
            self._solver = (basis, solver)
        return self._solver[1]

    def apply(self, w):
        if not isinstance(w, MultiVector):
            raise TypeError("PreconditioningOperator.apply needs a MultiVector")
        slab_in = w if isinstance(w, SlabMultiVector) else SlabMultiVector.from_multivector(w)
        solver = self.slab_operator(slab_in)
        if isinstance(solver, CsrSlabOperator) and solver.plan.L != len(slab_in):
            solver.plan.close()
            solver.plan = SGPlan(solver.blocks, len(slab_in))
        out = slab_in.like(torch.empty_like(slab_in.slab))
        solver.apply_slab(slab_in.slab, out.slab)
        if isinstance(w, SlabMultiVector):
            return out
        return out.to_multivector(type(w))
