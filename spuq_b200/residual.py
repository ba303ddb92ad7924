"""Slab-native pieces of the residual estimator.

The estimator of spuq/application/egsz/residual_estimator2.py:65-137 evaluates FEM forms (volume and
edge residuals, FEniCS) of, for every active mu and term m,

    res = -beta[0] * w[mu] + beta[1] * w[mu + e_m] + beta[-1] * w[mu - e_m]          (:117-131)

with ``beta = rv_m.orth_polys.get_beta(mu[m])``.  The forms need the discretisation and are out of
scope; the combination itself is the same neighbour gather as in ``MultiOperator.apply`` and is
provided here for all columns of one term at once (``spuq_sg_neighbour_combine``, one streaming
kernel, same operation order as the reference so the result is bit-identical), together with the
per-column energy norms the tail bounds use.
"""
import numpy as np
import torch

from . import _abi
from .lambda_tables import LambdaTables
from .multi_operator import CsrSlabOperator
from .multi_vector import MultiVector, SlabMultiVector

__all__ = ["neighbour_combination", "column_energy_norms"]


def _tables_for(w, coeff_field, maxm):
    return LambdaTables(w.index_array, [coeff_field[m][1] for m in range(maxm)], maxm)


def neighbour_combination(w, coeff_field, m, tables=None):
    """SlabMultiVector of ``res`` (residual_estimator2.py:117-131) for term ``m`` and every active mu."""
    if not isinstance(w, MultiVector):
        raise TypeError("expected a MultiVector, got %s" % type(w).__name__)
    w = w if isinstance(w, SlabMultiVector) else SlabMultiVector.from_multivector(w)
    maxm = min(w.max_order, len(coeff_field))                # :101-105
    if not 0 <= m < maxm:
        raise IndexError("term %d outside 0..%d" % (m, maxm - 1))
    tables = tables or _tables_for(w, coeff_field, maxm)
    dev = w.slab.device
    up = torch.from_numpy(np.ascontiguousarray(tables.nbr_up[m], dtype=np.int32)).to(dev)
    dn = torch.from_numpy(np.ascontiguousarray(tables.nbr_dn[m], dtype=np.int32)).to(dev)
    bu = torch.from_numpy(np.ascontiguousarray(tables.beta_up[m], dtype=np.float64)).to(dev)
    bd = torch.from_numpy(np.ascontiguousarray(tables.beta_dn[m], dtype=np.float64)).to(dev)
    b0 = torch.from_numpy(np.ascontiguousarray(tables.beta_0[m], dtype=np.float64)).to(dev)
    L, n = w.slab.shape
    out = torch.empty_like(w.slab)
    _abi.check(_abi.lib().spuq_sg_neighbour_combine(n, L, _abi.ptr(w.slab), n, _abi.ptr(up), _abi.ptr(dn), _abi.ptr(bu),
                                                    _abi.ptr(bd), _abi.ptr(b0), _abi.ptr(out), n, _abi.stream_ptr()),
               "neighbour_combine")
    return w.like(out)


def column_energy_norms(w, stiffness):
    """sqrt(w[mu]^T K w[mu]) for every active mu (K: a scipy sparse matrix, e.g. the mean stiffness)."""
    w = w if isinstance(w, SlabMultiVector) else SlabMultiVector.from_multivector(w)
    op = CsrSlabOperator(stiffness, len(w))
    T = torch.empty_like(w.slab)
    op.apply_slab(w.slab, T)
    return torch.sqrt(torch.clamp((w.slab * T).sum(dim=1), min=0.0))
