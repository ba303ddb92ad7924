"""Shared problem builders for the parity tests (CPU-emulated and GPU runs use the same cases).

Each case returns the synthetic inputs (host tensors), the oracle objects built from the SAME
numbers, and the oracle's result, so a test only has to push the inputs through the product path
and compare.
"""
import numpy as np
import scipy.sparse as sps

from oracle import linalg as olin
from oracle import multiindex as omi
from oracle import polys as opolys
from oracle import sg as osg
from spuq_b200 import synthetic


def oracle_rv(kind, mu=0.0):
    return opolys.UniformRV(-1, 1) if kind == "legendre" else opolys.NormalRV(mu, 1)


def oracle_operator(mats, rvs, dense=False):
    """Oracle MultiOperator over cached leaf operators (one per block, like the synthetic
    callbacks of SURVEY.md 8d): ScipyOperator leaves, or dense MatrixOperator leaves (config C1)."""
    M = len(mats) - 1
    cf = osg.ListCoefficientField("mean", list(range(M)), rvs)
    cache = {}

    def assemble(basis, f):
        k = 0 if f == "mean" else 1 + f
        if k not in cache:
            cache[k] = (olin.MatrixOperator(mats[k].toarray(), basis, basis) if dense
                        else olin.ScipyOperator(mats[k], basis, basis))
        return cache[k]
    return osg.MultiOperator(cf, assemble), cf


def oracle_multivector(idx, W, basis=None):
    """idx: L x M array (any order), W: N x L array with matching columns."""
    basis = basis or olin.CanonicalBasis(W.shape[0])
    w = osg.MultiVector()
    for k, row in enumerate(idx):
        w[omi.Multiindex(np.asarray(row, dtype=np.int64))] = olin.FlatVector(W[:, k].copy(), basis)
    return w


class Case:
    pass


def fd_case(n, M, idx, kind="legendre", mu=0.0, seed=1234, dense=False, with_oracle=True):
    """Synthetic diffusion problem on an n x n grid with M terms over the index set idx (sorted)."""
    c = Case()
    c.n, c.N, c.M, c.idx, c.kind, c.mu = n, n * n, M, idx, kind, mu
    c.rowptr, c.colidx, c.vals = synthetic.fd_blocks(n, M)
    c.mats = synthetic.blocks_to_scipy(c.rowptr, c.colidx, c.vals, c.N)
    rng = np.random.RandomState(seed)
    c.W = rng.random_sample((c.N, idx.shape[0]))
    c.rvs = synthetic.make_rvs(kind, M, mu)
    if with_oracle:
        c.orvs = [oracle_rv(kind, mu)] * M
        c.A, c.cf = oracle_operator(c.mats, c.orvs, dense)
        c.w = oracle_multivector(idx, c.W)
        c.Y = (c.A * c.w).to_slab()
    return c


def relerr(Y, Yref):
    """max over columns mu of ||y[mu] - yref[mu]||_2 / ||yref[mu]||_2: the per-column criterion (a column
    whose reference is exactly zero must be reproduced as zero)."""
    num = np.linalg.norm(Y - Yref, axis=0)
    den = np.linalg.norm(Yref, axis=0)
    ok = den > 0
    worst = float(np.max(num[ok] / den[ok])) if ok.any() else 0.0
    if (~ok).any() and float(np.max(num[~ok])) > 0.0:
        return float("inf")
    return worst


def random_sparse_case(N, M, idx, density=0.05, seed=7, kind="hermite", mu=0.5):
    """Unstructured blocks with different patterns per block (forces the union-pattern code and the
    CSR/ELL kernels), non-symmetric RV so beta0 != 0."""
    c = Case()
    c.N, c.M, c.idx, c.kind, c.mu = N, M, idx, kind, mu
    rng = np.random.RandomState(seed)
    c.mats = []
    for k in range(M + 1):
        m = sps.random(N, N, density=density, random_state=rng, format="csr")
        m = m + sps.identity(N) * (2.0 + k)
        c.mats.append(sps.csr_matrix(m))
    c.W = rng.random_sample((N, idx.shape[0]))
    c.rvs = synthetic.make_rvs(kind, M, mu)
    c.orvs = [oracle_rv(kind, mu)] * M
    c.A, c.cf = oracle_operator(c.mats, c.orvs)
    c.w = oracle_multivector(idx, c.W)
    c.Y = (c.A * c.w).to_slab()
    return c
