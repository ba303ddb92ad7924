"""TEST INFRASTRUCTURE (oracle side): the reference's apply loop on a SAMPLE of output columns.

Output column mu of MultiOperator.apply (multi_operator2.py:52-83) depends only on w[mu] and on the
neighbours w[mu +- e_m]; everything else about the loop body is independent of the other columns.
For problems whose full oracle apply would take hours (config 3: 988 x 51 SpMVs of 10^6 rows) the
checker therefore runs `oracle.sg.MultiOperator.apply(w, only=...)` -- the unchanged loop body -- on a
few output columns, with an oracle MultiVector that holds real data in the columns that body reads
and one shared zero vector everywhere else.  Index set, order, neighbour look-ups (`mu.inc(m) in
Lambda`) and beta values are all the oracle's own: nothing here comes from the product's tables.
"""
import time

import numpy as np

from . import linalg as olin
from . import multiindex as omi
from . import polys as opolys
from . import sg as osg
from . import synth

__all__ = ["oracle_rv", "SampledApply"]


def oracle_rv(kind, mu=0.0):
    return opolys.UniformRV(-1, 1) if kind == "legendre" else opolys.NormalRV(mu, 1)


class SampledApply:
    def __init__(self, mats, idx, kind, mu, pick, get_column):
        """mats: scipy CSR blocks [A_0, A_1..A_M]; idx: sorted L x M index array; pick: positions of the
        output columns to evaluate; get_column(k) -> 1-D array, the k-th column of the input slab."""
        self.N = N = mats[0].shape[0]
        self.L = L = idx.shape[0]
        M = len(mats) - 1
        rv = oracle_rv(kind, mu)
        cf = osg.ListCoefficientField("mean", list(range(M)), [rv] * M)
        basis = olin.CanonicalBasis(N)
        leaves = {}

        def assemble(b, f):
            k = 0 if f == "mean" else 1 + f
            if k not in leaves:
                leaves[k] = olin.ScipyOperator(mats[k], b, b)
            return leaves[k]

        self.A = osg.MultiOperator(cf, assemble)
        self.pick = [int(k) for k in pick]
        touched = synth.touched_columns(idx, self.pick)
        shared = olin.FlatVector(np.zeros(N), basis)     # untouched columns share one zero vector
        keys = [omi.Multiindex(r.astype(np.int64)) for r in idx]
        self.w = osg.MultiVector()
        for k, key in enumerate(keys):
            self.w[key] = olin.FlatVector(np.asarray(get_column(k), dtype=np.float64), basis) if k in touched else shared
        Lambda = self.w.active_indices()
        assert all(Lambda[k] == keys[k] for k in self.pick), "index array is not in active_indices() order"
        self.only = [Lambda[k] for k in self.pick]

    def run(self, only=None):
        """{position: y[mu]} for the sampled columns and the seconds it took."""
        t0 = time.perf_counter()
        v = self.A.apply(self.w, only=self.only if only is None else only)
        dt = time.perf_counter() - t0
        return {k: v[mu].coeffs for k, mu in zip(self.pick, self.only)} if only is None else v, dt
