"""Evaluation of the computed expansion at parameter samples, on the slab.

Drop-in for ``compute_parametric_sample_solution(RV_samples, coeff_field, w, proj_basis, cache)`` of
spuq/application/egsz/sampling.py:59-75: ``sum_mu w[mu] * prod_m
P_{mu_m}(y_m)``.  The reference projects every w[mu] onto a common basis first (:46-52); with a shared
basis that projection is the identity, and the sum over Lambda is one vector-matrix product with the
slab (``spuq_sg_combine``: a streaming pass, 8 n |Lambda| bytes per 8 samples).
``sample_solutions`` evaluates many samples in one call (what a Monte-Carlo loop over
``compute_parametric_sample_solution`` does one by one).
"""
import numpy as np
import torch

from . import _abi
from .linalg import FlatVector
from .multi_vector import MultiVector, SlabMultiVector

__all__ = ["compute_parametric_sample_solution", "sample_solutions", "sample_coefficients"]


def sample_coefficients(coeff_field, Lambda, samples):
    """S x |Lambda| array of prod_m P_{mu_m}(y^s_m): one ``sample_realization`` per sample."""
    samples = [list(y) for y in samples]
    out = np.empty((len(samples), len(Lambda)), dtype=np.float64)
    for s, y in enumerate(samples):
        sample_map, _ = coeff_field.sample_realization(Lambda, y)
        out[s] = [sample_map[mu] for mu in Lambda]
    return out


def sample_solutions(samples, coeff_field, w):
    """S x n device tensor: the expansion ``w`` evaluated at every parameter sample of ``samples``."""
    if not isinstance(w, MultiVector):
        raise TypeError("expected a MultiVector, got %s" % type(w).__name__)
    w = w if isinstance(w, SlabMultiVector) else SlabMultiVector.from_multivector(w)
    Lambda = w.active_indices()
    coef = sample_coefficients(coeff_field, Lambda, samples)
    S, L = coef.shape
    slab = w.slab                                         # L x n, i.e. the column-major n x L slab
    n = slab.shape[1]
    dev = slab.device
    coef_dev = torch.from_numpy(np.ascontiguousarray(coef)).to(dev)
    out = torch.empty((S, n), dtype=torch.float64, device=dev)
    _abi.check(_abi.lib().spuq_sg_combine(n, L, S, _abi.ptr(slab), n, _abi.ptr(coef_dev), _abi.ptr(out), n,
                                          _abi.stream_ptr()), "combine")
    return out


def compute_parametric_sample_solution(RV_samples, coeff_field, w, proj_basis=None, cache=None):
    """sampling.py:59-75 for vectors on one shared basis; returns a ``FlatVector`` on that basis.
    ``proj_basis`` must be that basis (or None); ``cache`` is accepted and ignored (it only held the
    per-mu projections)."""
    wb = w.spatial_basis if isinstance(w, SlabMultiVector) else None
    if proj_basis is not None and wb is not None and proj_basis != wb:
        raise ValueError("compute_parametric_sample_solution: the projection basis must be the shared basis of w")
    vals = sample_solutions([RV_samples], coeff_field, w)
    basis = wb if wb is not None else w[w.active_indices()[0]].basis
    return FlatVector(vals[0].cpu().numpy(), basis)
