"""K5: integer set-up of the active index set Lambda and the beta coefficient tables.

Replaces, for a whole set at once, the per-index Python work of the reference's apply loop
(spuq/application/egsz/multi_operator2.py:43,66,72-79): ``sorted(keys)``, ``mu.inc(m) in Lambda``,
``mu.dec(m) in Lambda`` and ``rv.orth_polys.get_beta(mu[m])``.  Sorting and neighbour lookup run in
the C++ host code of libspuq_sg (``spuq_sg_lambda_sort`` / ``spuq_sg_lambda_neighbours``); the beta
values come from ``polyquad`` (one small table per expansion term instead of a new polynomial-family
object per (mu, m) as in spuq/stochastics/random_variable.py:151-155,180-182).
"""
import numpy as np

from . import _abi
from .multiindex import Multiindex
from .polyquad import beta_table

__all__ = ["as_index_array", "sort_permutation", "neighbour_tables", "beta_tables", "LambdaTables"]


def as_index_array(indices, width=None):
    """L x M int32 array (zero padded) from Multiindex objects or an integer array."""
    if isinstance(indices, np.ndarray):
        arr = np.ascontiguousarray(indices, dtype=np.int32)
        if arr.ndim != 2:
            raise ValueError("index array must be L x M")
        if width is not None and width > arr.shape[1]:
            arr = np.hstack([arr, np.zeros((arr.shape[0], width - arr.shape[1]), np.int32)])
        return np.ascontiguousarray(arr)
    indices = list(indices)
    w = max([len(mu) for mu in indices] + [width or 0, 0])
    arr = np.zeros((len(indices), w), dtype=np.int32)
    for k, mu in enumerate(indices):
        if not isinstance(mu, Multiindex):
            mu = Multiindex(mu)
        arr[k, :len(mu)] = mu.as_array
    return arr


def sort_permutation(arr):
    """Permutation that puts the rows of ``arr`` into ``MultiVector.active_indices()`` order."""
    arr = np.ascontiguousarray(arr, dtype=np.int32)
    L, M = arr.shape
    perm = np.empty(L, dtype=np.int32)
    _abi.check(_abi.host_lib().spuq_sg_lambda_sort(_abi.ptr(arr), L, M, _abi.ptr(perm)), "lambda_sort")
    return perm


def neighbour_tables(sorted_arr):
    """(nbr_up, nbr_dn): M x L int32 positions of mu+e_m / mu-e_m in the sorted set, -1 = absent."""
    arr = np.ascontiguousarray(sorted_arr, dtype=np.int32)
    L, M = arr.shape
    up = np.full((M, L), -1, dtype=np.int32)
    dn = np.full((M, L), -1, dtype=np.int32)
    if L and M:
        _abi.check(_abi.host_lib().spuq_sg_lambda_neighbours(_abi.ptr(arr), L, M, _abi.ptr(up), _abi.ptr(dn)),
                   "lambda_neighbours")
    return up, dn


def beta_tables(sorted_arr, rvs, maxm):
    """(beta_up, beta_dn, beta_0): maxm x L fp64 with the entries of
    ``rvs[m].orth_polys.get_beta(mu[m])`` = (beta0, beta_up, beta_dn) for every (m, mu)."""
    arr = np.asarray(sorted_arr)
    L = arr.shape[0]
    b_up = np.zeros((maxm, L)); b_dn = np.zeros((maxm, L)); b_0 = np.zeros((maxm, L))
    cache = {}
    for m in range(maxm):
        deg = arr[:, m].astype(np.int64) if m < arr.shape[1] else np.zeros(L, np.int64)
        rv = rvs[m]
        key = id(rv)
        need = int(deg.max()) if L else 0
        tab = cache.get(key)
        if tab is None or tab.shape[0] <= need:
            tab = cache[key] = beta_table(rv.orth_polys, need)
        b_0[m], b_up[m], b_dn[m] = tab[deg, 0], tab[deg, 1], tab[deg, 2]
    return b_up, b_dn, b_0


class LambdaTables:
    """Everything the SG plan needs to know about Lambda, in sorted order."""

    def __init__(self, indices, rvs, maxm=None):
        arr = as_index_array(indices)
        perm = sort_permutation(arr)
        self.perm = perm
        self.index_array = np.ascontiguousarray(arr[perm])
        L, width = self.index_array.shape
        # maxm = min(longest index, len(coeff_field)); multi_operator2.py:44-48
        used = self.index_array.any(axis=0)
        max_order = int(np.flatnonzero(used).max()) + 1 if used.any() else 0
        self.max_order = max_order
        if maxm is None:
            maxm = max_order
        self.maxm = int(min(maxm, max_order))
        if self.index_array.shape[1] < self.maxm:
            self.index_array = as_index_array(self.index_array, self.maxm)
        core = np.ascontiguousarray(self.index_array[:, :max(self.maxm, 0)])
        # neighbours must be looked up in the FULL set (all components), not in the truncated one
        full_up, full_dn = neighbour_tables(self.index_array)
        self.nbr_up = np.ascontiguousarray(full_up[:self.maxm])
        self.nbr_dn = np.ascontiguousarray(full_dn[:self.maxm])
        self.beta_up, self.beta_dn, self.beta_0 = beta_tables(core, rvs, self.maxm)
        self.L = L

    def key(self):
        return (self.index_array.shape, self.index_array.tobytes(), self.maxm)

    def grown(self, indices, rvs, maxm=None):
        """Tables of the set ``indices`` (a superset of this one) by UPDATING the neighbour tables instead
        of looking every (mu, m) up again: the positions of the old entries are shifted to their places in
        the new sorted order, and only the links of the added indices are looked up (``mu +- e_m`` of an
        added mu, both directions).  Bit-identical to ``LambdaTables(indices, rvs, maxm)``; returns None when
        the update does not apply (not a superset, or the number of terms changes), the caller then
        rebuilds."""
        arr = as_index_array(indices)
        if arr.shape[0] < self.L:
            return None
        width = max(arr.shape[1], self.index_array.shape[1])
        arr = as_index_array(arr, width)
        old = as_index_array(self.index_array, width)
        perm = sort_permutation(arr)
        srt = np.ascontiguousarray(arr[perm])
        pos = {tuple(int(v) for v in r): k for k, r in enumerate(srt)}
        try:
            newpos = np.array([pos[tuple(int(v) for v in r)] for r in old], dtype=np.int64)
        except KeyError:
            return None                                              # an old index is missing: not a superset
        used = srt.any(axis=0)
        max_order = int(np.flatnonzero(used).max()) + 1 if used.any() else 0
        new_maxm = int(min(max_order if maxm is None else maxm, max_order))
        if new_maxm != self.maxm or max_order != self.max_order:
            return None
        L = srt.shape[0]
        out = object.__new__(LambdaTables)
        out.perm, out.index_array, out.L = perm, srt, L
        out.max_order, out.maxm = max_order, new_maxm
        up = np.full((new_maxm, L), -1, dtype=np.int32)
        dn = np.full((new_maxm, L), -1, dtype=np.int32)
        shift = np.concatenate([newpos, [-1]]).astype(np.int32)       # old position -> new position, -1 -> -1
        up[:, newpos] = shift[self.nbr_up]
        dn[:, newpos] = shift[self.nbr_dn]
        is_old = np.zeros(L, dtype=bool)
        is_old[newpos] = True
        for k in np.flatnonzero(~is_old):                             # links of the added indices only
            t = [int(v) for v in srt[k]]
            for m in range(new_maxm):
                t[m] += 1
                j = pos.get(tuple(t))
                if j is not None:
                    up[m, k] = j
                    dn[m, j] = k
                t[m] -= 2
                if t[m] >= 0:
                    j = pos.get(tuple(t))
                    if j is not None:
                        dn[m, k] = j
                        up[m, j] = k
                t[m] += 1
        out.nbr_up, out.nbr_dn = up, dn
        core = np.ascontiguousarray(srt[:, :max(new_maxm, 0)])
        out.beta_up, out.beta_dn, out.beta_0 = beta_tables(core, rvs, new_maxm)
        return out
