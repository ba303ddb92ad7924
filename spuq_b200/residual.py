"""Slab-native pieces of the residual estimator.

The estimator of spuq/application/egsz/residual_estimator2.py:65-137 evaluates FEM forms (volume and
edge residuals, FEniCS) of, for every active mu and term m,

    res = -beta[0] * w[mu] + beta[1] * w[mu + e_m] + beta[-1] * w[mu - e_m]          (:117-131)

with ``beta = rv_m.orth_polys.get_beta(mu[m])``.  The forms need the discretisation and are out of
scope; the combination itself is the same neighbour gather as in ``MultiOperator.apply`` and is
provided here for all columns of one term at once (``spuq_sg_neighbour_combine``, one streaming
kernel, same operation order as the reference so the result is bit-identical), together with the
per-column energy norms the tail bounds use.

``evaluate_upper_tail_bound`` is ``ResidualEstimator.evaluateUpperTailBound`` (:176-290): the energy norms
of all columns come from ONE slab product on the device, the boundary of Lambda and the sums (3.11) /
(3.15) / (3.16) are integer and scalar work over |boundary| x M on the host.
"""
import numpy as np
import torch

from . import _abi
from .lambda_tables import LambdaTables
from .multi_operator import CsrSlabOperator
from .multi_vector import MultiVector, SlabMultiVector

__all__ = ["neighbour_combination", "column_energy_norms", "lambda_boundary", "evaluate_upper_tail_bound"]


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


def lambda_boundary(index_array):
    """The indices mu +- e_m (m in the joint support of Lambda) that are not in Lambda, as the reference's
    generator yields them (residual_estimator2.py:213-223): a dict {index tuple: number of times it is
    yielded} -- an index adjacent to k members of Lambda is yielded k times, and :279 adds its zeta k times.
    Tuples are zero padded to the width of ``index_array``."""
    arr = np.asarray(index_array)
    members = {tuple(int(v) for v in r) for r in arr}
    support = [int(m) for m in np.flatnonzero(arr.any(axis=0))]
    count = {}
    for mu in members:
        t = list(mu)
        for m in support:
            t[m] += 1
            up = tuple(t)
            if up not in members:
                count[up] = count.get(up, 0) + 1
            t[m] -= 2
            if t[m] >= 0:
                dn = tuple(t)
                if dn not in members:
                    count[dn] = count.get(dn, 0) + 1
            t[m] += 1
    return count


def evaluate_upper_tail_bound(w, coeff_field, stiffness, ainfty, add_maxm=10):
    """(global_zeta, zeta, zeta_bar, eval_zeta_m) of residual_estimator2.py:176-290.

    ``stiffness``: the matrix of the energy norm (pde.energy_norm, :266; the mean stiffness);
    ``ainfty(m)``: ||a_m / a_mean||_inf (:181-205, a property of the coefficient functions).
    zeta / zeta_bar are dicts keyed by Multiindex like the reference's."""
    from math import sqrt
    from .multiindex import Multiindex
    ws = w if isinstance(w, SlabMultiVector) else SlabMultiVector.from_multivector(w)
    arr = np.asarray(ws.index_array)
    L, width = arr.shape
    M = ws.max_order + add_maxm                                     # :271
    try:
        M = min(M, len(coeff_field))                                # the reference's own variant (:270)
    except TypeError:
        pass
    normw = column_energy_norms(ws, stiffness).cpu().numpy()        # :207-211, one slab product
    if width < M:
        arr = np.hstack([arr, np.zeros((L, M - width), arr.dtype)])
    pos = {tuple(int(v) for v in r): k for k, r in enumerate(arr)}
    in_supp = np.zeros(max(M, arr.shape[1]), dtype=bool)
    in_supp[np.flatnonzero(arr.any(axis=0))] = True
    a_inf = [float(ainfty(m)) for m in range(M)]
    beta = [[coeff_field[m][1].orth_polys.get_beta(k) for k in range(int(arr[:, m].max()) + 3)] for m in range(M)]

    def eval_zeta(nu):                                              # (3.11), :241-255
        z = 0
        t = list(nu)
        for m in range(M):
            b = beta[m][t[m]] if t[m] < len(beta[m]) else coeff_field[m][1].orth_polys.get_beta(t[m])
            t[m] += 1
            k = pos.get(tuple(t))
            if k is not None:
                z += a_inf[m] * b[1] * normw[k]
            t[m] -= 2
            if t[m] >= 0:
                k = pos.get(tuple(t))
                if k is not None:
                    z += a_inf[m] * b[-1] * normw[k]
            t[m] += 1
        return z

    key = lambda t: Multiindex(np.array(t, dtype=np.int64))           # noqa: E731
    zeta = {}
    for nu, times in lambda_boundary(arr).items():
        val = eval_zeta(nu)
        acc = 0
        for _ in range(times):                                      # `zeta[nu] += ...` once per yield (:279)
            acc += val
        zeta[key(nu)] = acc
    zeta_bar = {}
    for k, r in enumerate(arr):                                      # (3.15), :226-238
        zz = 0
        for m in range(M):
            if in_supp[m]:
                continue
            zz += (beta[m][int(r[m])][1] * a_inf[m]) ** 2
        zeta_bar[key(tuple(int(v) for v in r))] = normw[k] * sqrt(zz)
    global_zeta = sqrt(sum(v ** 2 for v in zeta.values()) + sum(v ** 2 for v in zeta_bar.values()))   # (3.16), :287

    def eval_zeta_m(mu, m):                                         # :256-262
        t = tuple(int(v) for v in mu.padded(arr.shape[1])) if hasattr(mu, "padded") else tuple(mu)
        b = coeff_field[m][1].orth_polys.get_beta(t[m] if m < len(t) else 0)
        return a_inf[m] * b[1] * normw[pos[t]]
    return global_zeta, zeta, zeta_bar, eval_zeta_m
