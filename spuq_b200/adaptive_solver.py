"""The direct caller of ``pcg``: right-hand side and solve of the SG system, slab-native.

Drop-in for ``prepare_rhs(A, w, coeff_field, pde)`` and ``pcg_solve(A, w, coeff_field, pde, stats,
pcg_eps, pcg_maxiter)`` of spuq/application/egsz/adaptive_solver2.py:39-85:
same arguments, same results, same ``stats`` keys.  ``pde`` is the reference's discretisation object
(``FEMDiscretisation``, fem_discretisation.py:228-352) as far as this path uses it:
``assemble_rhs``, ``set_dirichlet_bc_entries``, ``assemble_solve_operator``, ``energy_norm`` and a
``norm(array, "L2")`` standing for dolfin's ``norm`` (fenics_utils.py:140-159).

What is different from the reference is only where the data live: ``w`` is a ``SlabMultiVector``,
the right-hand side is written straight into the columns of a zero slab (at most M + 1 of the L
columns are non-zero), the solve runs on the device (``pcg``), and the residual ``b - A w`` and its
norms are evaluated on the slab without converting back to a dictionary of vectors.
"""
import math

import numpy as np
import scipy.sparse as sps
import torch

from . import _abi
from .multi_operator import CsrSlabOperator, PreconditioningOperator, _block_to_csr
from .multi_vector import MultiVector, SlabMultiVector
from .multiindex import Multiindex
from .pcg import pcg

__all__ = ["prepare_rhs", "pcg_solve", "error_norm"]


def _as_slab(w):
    if isinstance(w, SlabMultiVector):
        return w
    if isinstance(w, MultiVector):
        return SlabMultiVector.from_multivector(w)
    raise TypeError("expected a MultiVector, got %s" % type(w).__name__)


def prepare_rhs(A, w, coeff_field, pde):
    """b[0] = rhs(mean) + sum_m beta_m(0)[0] g_m,  b[e_m] = beta_m(0)[1] g_m'  for m < w.max_order,
    every other column zero; adaptive_solver2.py:39-68 (same order of operations per column)."""
    w = _as_slab(w)
    b = 0 * w                                                               # :40
    zero = Multiindex()
    active = b.active_indices()
    basis = w.spatial_basis
    col0 = np.array(pde.assemble_rhs(basis=basis, coeff=coeff_field.mean_func, withNeumannBC=True),
                    dtype=np.float64)                                       # :43-44
    zero_func = 0.0                                                         # the zero load of :45-50

    class _Col:                       # what set_dirichlet_bc_entries sees: a vector with .coeffs / .basis
        def __init__(self, coeffs):
            self.coeffs, self.basis = coeffs, basis

    for m in range(w.max_order):                                            # :52
        eps_m = zero.inc(m)
        am_f, am_rv = coeff_field[m]
        beta = am_rv.orth_polys.get_beta(0)
        if eps_m in active:                                                 # :57
            g0 = _Col(np.array(pde.assemble_rhs(basis=basis, coeff=am_f, withNeumannBC=False, f=zero_func),
                               dtype=np.float64))
            pde.set_dirichlet_bc_entries(g0, homogeneous=True)
            b[eps_m].coeffs = 0.0 + beta[1] * g0.coeffs                     # :61 (the column was zero)
        g0 = _Col(np.array(pde.assemble_rhs(basis=basis, coeff=am_f, f=zero_func), dtype=np.float64))
        pde.set_dirichlet_bc_entries(g0, homogeneous=True)
        col0 += beta[0] * g0.coeffs                                         # :66
    b[zero].coeffs = col0
    return b


def error_norm(vec1, vec2, normstr, pde):
    """sqrt(sum_mu norm(vec1[mu] - vec2[mu])^2), fenics_utils.py:140-152, on slabs.

    ``"L2"`` uses ``pde.l2_weight`` (lumped mass, scalar or per-node array) when the discretisation
    offers it, and the energy norm the mean stiffness block ``pde.assemble_lhs(basis, mean)``: both
    then are one pass over the slab on the device.  Otherwise the per-column host functions of the
    reference are called."""
    v1, v2 = _as_slab(vec1), _as_slab(vec2)
    if v1.active_indices() != v2.active_indices():
        raise AssertionError("error_norm: different index sets")
    D = v1.slab - v2.slab                                                   # L x N
    if isinstance(normstr, str):
        wgt = getattr(pde, "l2_weight", None) if normstr == "L2" else None
        if wgt is not None:
            if np.ndim(wgt) == 0:
                return math.sqrt(float(wgt) * float(torch.sum(D * D)))
            wt = torch.as_tensor(np.asarray(wgt, dtype=np.float64), device=D.device)
            return math.sqrt(float(torch.sum(D * D * wt[None, :])))
        cols = D.cpu().numpy()
        return math.sqrt(sum(pde.norm(cols[k], normstr) ** 2 for k in range(cols.shape[0])))
    stiff = getattr(normstr, "stiffness", None)
    if stiff is not None:
        op = CsrSlabOperator(stiff, D.shape[0])
        T = torch.empty_like(D)
        op.apply_slab(D, T)
        return math.sqrt(max(float(torch.sum(D * T)), 0.0))
    cols = D.cpu().numpy()
    return math.sqrt(sum(normstr(cols[k]) ** 2 for k in range(cols.shape[0])))


class _EnergyNorm:
    """pde.energy_norm with the matrix it is the norm of attached, so that ``error_norm`` can apply
    it to a whole slab on the device."""

    def __init__(self, fn, stiffness):
        self._fn, self.stiffness = fn, stiffness

    def __call__(self, v):
        return self._fn(v)


def pcg_solve(A, w, coeff_field, pde, stats, pcg_eps, pcg_maxiter):
    """adaptive_solver2.py:71-85.  Returns ``(w, zeta)`` and fills ``stats`` with RESIDUAL-L2,
    RESIDUAL-H1A, DOFS and CELLS; also TIME-PCG, the wall time of the call with the device drained (the
    reference's adaptive loop stores it around this call, adaptive_solver2.py:150)."""
    import time
    t0 = time.perf_counter()
    w = _as_slab(w)
    b = prepare_rhs(A, w, coeff_field, pde)
    P = PreconditioningOperator(coeff_field.mean_func, pde.assemble_solve_operator)
    w, zeta, numit = pcg(A, b, P, w0=w, eps=pcg_eps, maxiter=pcg_maxiter)
    stats["PCG-ITERATIONS"] = numit                  # (logged by the reference, :77)
    b2 = A * w                                                              # :79
    energy = pde.energy_norm
    if hasattr(pde, "assemble_lhs"):
        K0 = pde.assemble_lhs(w.spatial_basis, coeff_field.mean_func)
        energy = _EnergyNorm(energy, sps.csr_matrix(K0, dtype=np.float64) if sps.issparse(K0) else _block_to_csr(K0))
    stats["RESIDUAL-L2"] = error_norm(b, b2, "L2", pde)
    stats["RESIDUAL-H1A"] = error_norm(b, b2, energy, pde)
    stats["DOFS"] = len(b) * b.spatial_basis.dim
    stats["CELLS"] = len(b) * int(pde.num_cells(b.spatial_basis))
    if torch.cuda.is_available() and not _abi.is_emulated():
        torch.cuda.synchronize()
    stats["TIME-PCG"] = time.perf_counter() - t0
    return w, zeta
