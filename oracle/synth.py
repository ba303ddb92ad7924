"""TEST INFRASTRUCTURE (oracle side): the synthetic inputs of SURVEY.md section 8d in plain numpy/scipy.

The product generates the same inputs with torch (spuq_b200/synthetic.py) so that the benchmark-size
cases can be built directly in HBM; this module restates them with numpy only, so that the oracle,
`bench.py --impl reference` and the CPU baseline need nothing from the product package (no shared
object is loaded on that side).  tests/test_oracle_synth.py checks that both generators produce the
same numbers bit for bit.

Model problem of the reference's sample set-ups (spuq/application/egsz/sample_problems2.py):
unit square, homogeneous Dirichlet boundary, n x n interior unknowns, rows i = iy*n + ix, 5-point
stencil with face-averaged coefficient; mean a_0 = 1 and a_m(x) = A_m cos(pi k1 x1) cos(pi k2 x2)
(:36) with (k1, k2) the (m+1)-th pair of the 2-D complete-order generator (:201-203) and
A_m = amp / (m + start)^decay from the zeta rule (:164-169, :184-187).
"""
import math
from itertools import islice

import numpy as np
import scipy.sparse as sps
from scipy.special import zeta as _hurwitz_zeta

from .multiindex import Multiindex, MultiindexSet

__all__ = ["kl_frequencies", "kl_amplitudes", "fd_matrices", "tile_permutation", "permuted", "index_set", "sorted_indices",
           "touched_columns", "neighbour_positions"]


def kl_frequencies(M):
    gen = MultiindexSet.createCompleteOrderSet(2)
    next(gen)
    return [(int(mu[0]), int(mu[1])) for mu in islice(gen, M)]


def kl_amplitudes(M, gamma=0.9, decay=2.0):
    start = 1
    while _hurwitz_zeta(decay, start) >= gamma:
        start += 1
    amp = gamma / _hurwitz_zeta(decay, start)
    return [float(amp / (float(m) + start) ** decay) for m in range(M)]


def _nodal(n, k1, k2, amp):
    x = np.linspace(0.0, 1.0, n + 2)
    return amp * np.cos(math.pi * k2 * x)[:, None] * np.cos(math.pi * k1 * x)[None, :]


def _pattern(n):
    N = n * n
    i = np.arange(N, dtype=np.int64)
    ix, iy = i % n, i // n
    cols = np.stack([i - n, i - 1, i, i + 1, i + n], axis=1)
    valid = np.stack([iy > 0, ix > 0, np.ones(N, bool), ix < n - 1, iy < n - 1], axis=1)
    rowptr = np.zeros(N + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(valid.sum(axis=1))
    return rowptr, cols[valid], valid


def _fd_values(a, valid):
    c = a[1:-1, 1:-1]
    cs = 0.5 * (c + a[:-2, 1:-1])
    cn = 0.5 * (c + a[2:, 1:-1])
    cw = 0.5 * (c + a[1:-1, :-2])
    ce = 0.5 * (c + a[1:-1, 2:])
    vals = np.stack([-cs, -cw, cs + cw + ce + cn, -ce, -cn], axis=2).reshape(-1, 5)
    return vals[valid]


def fd_matrices(n, M, gamma=0.9, decay=2.0, mean=1.0):
    """[A_0, A_1, ..., A_M] as scipy CSR on one shared 5-point pattern."""
    rowptr, colidx, valid = _pattern(n)
    N = n * n
    amps = kl_amplitudes(M, gamma, decay)
    fields = [np.full((n + 2, n + 2), float(mean))]
    fields += [_nodal(n, k1, k2, amps[m]) for m, (k1, k2) in enumerate(kl_frequencies(M))]
    return [sps.csr_matrix((_fd_values(a, valid), colidx.astype(np.int32), rowptr.astype(np.int32)), shape=(N, N))
            for a in fields]


def tile_permutation(n, tile=16, seed=4321):
    """Renumbering of the n x n grid that keeps the locality of an unstructured-mesh numbering but
    destroys the stencil structure: tiles of tile x tile nodes in row-major tile order, the nodes of a
    tile in random order.  perm[new] = old."""
    rng = np.random.RandomState(seed)
    ids = np.arange(n * n).reshape(n, n)
    out = []
    for ty in range(0, n, tile):
        for tx in range(0, n, tile):
            blk = ids[ty:ty + tile, tx:tx + tile].ravel().copy()
            rng.shuffle(blk)
            out.append(blk)
    return np.concatenate(out)


def permuted(mats, perm):
    """P A P^T for every block: row/column new = old perm[new]; sorted indices, one shared pattern."""
    out = []
    for m in mats:
        q = sps.csr_matrix(m)[perm][:, perm].tocsr()
        q.sort_indices()
        out.append(q)
    return out


def sorted_indices(arr):
    """Rows of ``arr`` in MultiVector.active_indices() order (Multiindex.cmp_by_order)."""
    keys = sorted(Multiindex(np.asarray(r, dtype=np.int64)) for r in arr)
    width = arr.shape[1]
    out = np.zeros((len(keys), width), dtype=np.int32)
    for k, mu in enumerate(keys):
        a = mu.as_array
        out[k, :len(a)] = a
    return out


def index_set(spec):
    """spec = ("td", M, p) | ("aniso", M, T) | ("prefix", M, p, L): the index sets of SURVEY.md 8d."""
    kind = spec[0]
    if kind in ("td", "prefix"):
        arr = sorted_indices(np.asarray(MultiindexSet.createCompleteOrderSet(spec[1], spec[2]).arr))
        return arr if kind == "td" else np.ascontiguousarray(arr[:spec[3]])
    if kind == "aniso":
        M, T = spec[1], float(spec[2])
        w = [2.0 * math.log(m + 2.0) for m in range(M)]
        out, cur = [], [0] * M

        def rec(m, budget):
            if m == M:
                out.append(list(cur))
                return
            k = 0
            while k * w[m] <= budget + 1e-12:
                cur[m] = k
                rec(m + 1, budget - k * w[m])
                k += 1
            cur[m] = 0
        rec(0, T)
        return sorted_indices(np.array(out, dtype=np.int32))
    raise ValueError(spec)


def neighbour_positions(idx):
    """dict position -> list of positions of mu +- e_m that are in the (sorted) set ``idx``."""
    pos = {tuple(int(v) for v in r): k for k, r in enumerate(idx)}
    out = {}
    for k, r in enumerate(idx):
        t = [int(v) for v in r]
        nb = []
        for m in range(len(t)):
            t[m] += 1
            j = pos.get(tuple(t))
            if j is not None:
                nb.append(j)
            t[m] -= 2
            if t[m] >= 0:
                j = pos.get(tuple(t))
                if j is not None:
                    nb.append(j)
            t[m] += 1
        out[k] = nb
    return out


def touched_columns(idx, pick):
    """Positions of the columns an apply restricted to the output columns ``pick`` reads."""
    nb = neighbour_positions(idx)
    touched = set(int(k) for k in pick)
    for k in pick:
        touched.update(nb[int(k)])
    return touched
