"""Deterministic synthetic inputs for the SG operator.

The reference assembles its spatial blocks with FEniCS, which is not available; what it needs from
an assembly is only a callback ``(basis, coeff) -> Operator``.  This module supplies such blocks for
the model problem of the reference's sample set-ups (spuq/application/egsz/sample_problems2.py):

  * domain: unit square, homogeneous Dirichlet boundary, uniform n x n grid of interior unknowns,
    rows ordered i = iy*n + ix;
  * discretisation: 5-point finite-difference / lumped-P1 stiffness with face-averaged coefficient,
    all blocks on ONE shared CSR pattern;
  * coefficient: mean a_0 = 1 and  a_m(x) = A_m cos(pi k1 x1) cos(pi k2 x2)
    (sample_problems2.py:36), (k1, k2) the (m+1)-th pair of the 2-D complete-order generator
    (freqskip = 0, :201-203), amplitude A_m = amp / (m + start)^decay with decay = 2, gamma = 0.9
    and start / amp from the zeta rule (:164-169, :184-187), which keeps the operator SPD.

Everything is written with torch so the same code builds small cases on the host (for the oracle
and the parity tests) and the benchmark-size cases directly in HBM.
"""
import math
from itertools import islice

import numpy as np
import scipy.sparse as sps
import torch
from scipy.special import zeta as _hurwitz_zeta

from .multiindex import MultiindexSet
from .lambda_tables import sort_permutation
from .random_variable import NormalRV, UniformRV

__all__ = ["kl_frequencies", "kl_amplitudes", "fd_pattern", "fd_values", "fd_blocks", "tile_permutation", "permute_blocks",
           "fd_blocks_rows", "fd_pattern_rect", "complete_order_indices", "anisotropic_indices", "truncated_complete_indices",
           "make_rvs", "blocks_to_scipy", "FDDiscretisation", "GridBasis", "GridProlongation"]


def kl_frequencies(M):
    """(k1, k2) of terms 0..M-1: the 2-D complete-order sequence without its first element (0,0)."""
    gen = MultiindexSet.createCompleteOrderSet(2)
    next(gen)
    return [(int(mu[0]), int(mu[1])) for mu in islice(gen, M)]


def kl_amplitudes(M, gamma=0.9, decay=2.0):
    """A_m = amp / (m + start)^decay, with start the smallest integer such that
    zeta(decay, start) < gamma and amp = gamma / zeta(decay, start)."""
    start = 1
    while _hurwitz_zeta(decay, start) >= gamma:
        start += 1
    amp = gamma / _hurwitz_zeta(decay, start)
    return [float(amp / (float(m) + start) ** decay) for m in range(M)]


def fd_pattern(n, device="cpu"):
    """Shared CSR pattern of the 5-point stencil on an n x n grid (columns ascending in a row).
    Returns (rowptr int32 [N+1], colidx int32 [nnz], valid mask [N, 5])."""
    N = n * n
    i = torch.arange(N, device=device, dtype=torch.int64)
    ix, iy = i % n, i // n
    cols = torch.stack([i - n, i - 1, i, i + 1, i + n], dim=1)
    valid = torch.stack([iy > 0, ix > 0, torch.ones_like(ix, dtype=torch.bool), ix < n - 1, iy < n - 1], dim=1)
    counts = valid.sum(dim=1)
    rowptr = torch.zeros(N + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(counts, 0)
    colidx = cols[valid]
    return rowptr.to(torch.int32), colidx.to(torch.int32), valid


def fd_values(a_nodal, valid):
    """Stencil values on the shared pattern for a coefficient given at the (n+2) x (n+2) grid nodes
    (boundary ring included): off-diagonals are minus the face average, the diagonal is the sum of
    the four face coefficients (Dirichlet neighbours eliminated)."""
    c = a_nodal[1:-1, 1:-1]
    cs = 0.5 * (c + a_nodal[:-2, 1:-1])     # face towards iy-1
    cn = 0.5 * (c + a_nodal[2:, 1:-1])      # face towards iy+1
    cw = 0.5 * (c + a_nodal[1:-1, :-2])     # face towards ix-1
    ce = 0.5 * (c + a_nodal[1:-1, 2:])      # face towards ix+1
    vals = torch.stack([-cs, -cw, cs + cw + ce + cn, -ce, -cn], dim=2).reshape(-1, 5)
    return vals[valid]


def _coefficient(n, k1, k2, amp, device, dtype=torch.float64):
    # the 1-D cosine factors are evaluated with numpy on the host (n + 2 values) so that every device and
    # the numpy restatement used by the CPU-side checker see bit-identical blocks; the outer product and the
    # face averages are single IEEE operations and round the same everywhere
    x = np.linspace(0.0, 1.0, n + 2)
    cx = torch.from_numpy(np.cos(math.pi * k1 * x)).to(device=device, dtype=dtype)
    cy = torch.from_numpy(np.cos(math.pi * k2 * x)).to(device=device, dtype=dtype)
    return amp * cy[:, None] * cx[None, :]          # [iy, ix]: x1 runs along ix


def fd_blocks(n, M, device="cpu", gamma=0.9, decay=2.0, mean=1.0):
    """rowptr, colidx, vals[(M+1), nnz]: block 0 is the mean block, block 1+m the m-th term."""
    rowptr, colidx, valid = fd_pattern(n, device)
    vals = torch.empty((M + 1, colidx.numel()), dtype=torch.float64, device=device)
    vals[0] = fd_values(torch.full((n + 2, n + 2), float(mean), dtype=torch.float64, device=device), valid)
    amps = kl_amplitudes(M, gamma, decay)
    for m, (k1, k2) in enumerate(kl_frequencies(M)):
        vals[1 + m] = fd_values(_coefficient(n, k1, k2, amps[m], device), valid)
    return rowptr, colidx, vals


def fd_blocks_rows(n, M, y0, y1, device="cpu", gamma=0.9, decay=2.0, mean=1.0):
    """Blocks of the ROW SHARD [y0, y1) of the n x n grid with one halo grid row on each side: the
    local grid has (y1 - y0 + 2) x n points, local row j <-> global row y0 - 1 + j.  Values are the
    global stencil values of those rows; couplings that leave the local grid are dropped (they only
    touch the halo rows, whose outputs are not used), and a halo row outside the global grid (first /
    last shard) is empty.  Returns rowptr, colidx, vals[(M+1), nnz] on the local 5-point pattern."""
    nyl = y1 - y0 + 2
    rowptr, colidx, valid = fd_pattern_rect(n, nyl, device)

    def local_vals(a_nodal):
        c = a_nodal[1:-1, 1:-1]
        cs = 0.5 * (c + a_nodal[:-2, 1:-1]); cn = 0.5 * (c + a_nodal[2:, 1:-1])
        cw = 0.5 * (c + a_nodal[1:-1, :-2]); ce = 0.5 * (c + a_nodal[1:-1, 2:])
        full = torch.stack([-cs, -cw, cs + cw + ce + cn, -ce, -cn], dim=2)           # [n, n, 5] global rows
        pad = torch.zeros((1, n, 5), dtype=full.dtype, device=full.device)
        full = torch.cat([pad, full, pad], dim=0)                                   # global rows -1 .. n
        loc = full[y0:y1 + 2].reshape(-1, 5)                                        # global rows y0-1 .. y1
        return loc[valid]

    vals = torch.empty((M + 1, colidx.numel()), dtype=torch.float64, device=device)
    vals[0] = local_vals(torch.full((n + 2, n + 2), float(mean), dtype=torch.float64, device=device))
    amps = kl_amplitudes(M, gamma, decay)
    for m, (k1, k2) in enumerate(kl_frequencies(M)):
        vals[1 + m] = local_vals(_coefficient(n, k1, k2, amps[m], device))
    return rowptr, colidx, vals


def fd_pattern_rect(nx, ny, device="cpu"):
    """fd_pattern for an nx (fast) x ny grid."""
    N = nx * ny
    i = torch.arange(N, device=device, dtype=torch.int64)
    ix, iy = i % nx, i // nx
    cols = torch.stack([i - nx, i - 1, i, i + 1, i + nx], dim=1)
    valid = torch.stack([iy > 0, ix > 0, torch.ones_like(ix, dtype=torch.bool), ix < nx - 1, iy < ny - 1], dim=1)
    counts = valid.sum(dim=1)
    rowptr = torch.zeros(N + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr.to(torch.int32), cols[valid].to(torch.int32), valid


def tile_permutation(n, tile=16, seed=4321):
    """Renumbering of the n x n grid with the locality of an unstructured-mesh numbering but without the
    stencil structure: tiles of tile x tile nodes in row-major tile order, the nodes of a tile in random
    order.  perm[new] = old.  (Workload "c3u": the general sparse path on config-3 blocks.)"""
    rng = np.random.RandomState(seed)
    ids = np.arange(n * n).reshape(n, n)
    out = []
    for ty in range(0, n, tile):
        for tx in range(0, n, tile):
            blk = ids[ty:ty + tile, tx:tx + tile].ravel().copy()
            rng.shuffle(blk)
            out.append(blk)
    return np.concatenate(out)


def permute_blocks(rowptr, colidx, vals, perm):
    """P A P^T of all blocks on their shared pattern (row / column new = old perm[new]), column indices
    of a row ascending; runs where the tensors live (HBM for the benchmark sizes)."""
    dev = rowptr.device
    perm_t = torch.as_tensor(np.asarray(perm), dtype=torch.int64, device=dev)
    N = perm_t.numel()
    inv = torch.empty_like(perm_t)
    inv[perm_t] = torch.arange(N, device=dev)
    rp = rowptr.to(torch.int64)
    counts = (rp[1:] - rp[:-1])[perm_t]
    new_rp = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    new_rp[1:] = torch.cumsum(counts, 0)
    nnz = int(new_rp[-1])
    new_row = torch.repeat_interleave(torch.arange(N, device=dev), counts)
    within = torch.arange(nnz, device=dev) - new_rp[new_row]
    eidx = rp[perm_t][new_row] + within                       # entry of the old matrix
    new_col = inv[colidx.to(torch.int64)[eidx]]
    order = torch.argsort(new_row * N + new_col)              # rows are contiguous: a global sort sorts every row
    eidx = eidx[order]
    return new_rp.to(torch.int32), new_col[order].to(torch.int32), vals[:, eidx].contiguous()


def blocks_to_scipy(rowptr, colidx, vals, n_rows):
    """Host scipy CSR matrices of the blocks (for the oracle and for assemble callbacks)."""
    rp = rowptr.cpu().numpy().astype(np.int32)
    ci = colidx.cpu().numpy().astype(np.int32)
    v = vals.cpu().numpy()
    return [sps.csr_matrix((v[k].copy(), ci, rp), shape=(n_rows, n_rows)) for k in range(v.shape[0])]


# ---- index sets -----------------------------------------------------------------------------------
def _sorted(arr):
    arr = np.ascontiguousarray(arr, dtype=np.int32)
    return np.ascontiguousarray(arr[sort_permutation(arr)])


def complete_order_indices(M, p):
    """All mu in N^M with |mu| <= p, in active_indices() order."""
    return _sorted(MultiindexSet.createCompleteOrderSet(M, p).arr)


def truncated_complete_indices(M, p, L):
    """First L indices of the sorted complete-order set (a prefix of a graded order is
    downward closed)."""
    arr = complete_order_indices(M, p)
    assert L <= arr.shape[0]
    return np.ascontiguousarray(arr[:L])


def anisotropic_indices(M, T):
    """Downward-closed anisotropic set {mu : sum_m mu_m * 2 ln(m+2) <= T}."""
    w = [2.0 * math.log(m + 2.0) for m in range(M)]
    out = []
    cur = [0] * M

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

    rec(0, float(T))
    return _sorted(np.array(out, dtype=np.int32))


def make_rvs(kind, M, mu=0.0):
    if kind == "legendre":
        rv = UniformRV(-1, 1)
    elif kind == "hermite":
        rv = NormalRV(mu, 1)
    else:
        raise ValueError("unknown polynomial family %r" % kind)
    return [rv] * M


class FDDiscretisation:
    """Finite-difference stand-in for the reference's ``FEMDiscretisation``
    (spuq/application/egsz/fem_discretisation.py:228-352) on the synthetic blocks above: the
    methods ``prepare_rhs`` / ``pcg_solve`` call on it (adaptive_solver2.py:39-85), with the
    reference's names and argument meaning.  A coefficient is the token the coefficient field
    yields: ``"mean"`` or the term number m.  Unknowns are the interior nodes, so Dirichlet data
    enter only through the right-hand side (the coupling of an interior node to a boundary node
    times the boundary value) and ``set_dirichlet_bc_entries`` has nothing to overwrite.

    ``linalg`` is the module that provides ``ScipyOperator`` / ``ScipySolveOperator`` (this package,
    or the oracle's restatement in the tests): the discretisation is an INPUT of the path, the same
    object feeds both sides of a parity test."""

    def __init__(self, n, M, linalg, f=1.0, uD=None, gamma=0.9, decay=2.0, mean=1.0):
        self.n, self.M, self.h = int(n), int(M), 1.0 / (n + 1)
        self._la = linalg
        self.f, self.uD = f, uD
        _, _, self._valid = fd_pattern(n)
        rowptr, colidx, _ = fd_pattern(n)
        self._rowptr, self._colidx = rowptr.numpy().astype(np.int64), colidx.numpy().astype(np.int64)
        amps = kl_amplitudes(M, gamma, decay)
        self._nodal = {"mean": torch.full((n + 2, n + 2), float(mean), dtype=torch.float64)}
        for m, (k1, k2) in enumerate(kl_frequencies(M)):
            self._nodal[m] = _coefficient(n, k1, k2, amps[m], "cpu")
        self._lhs = {}
        x = np.linspace(0.0, 1.0, n + 2)
        self._x1, self._x2 = np.meshgrid(x, x)            # [iy, ix]: x1 along ix, as in _coefficient

    def _on_nodes(self, g):
        if g is None:
            return np.zeros((self.n + 2, self.n + 2))
        if callable(g):
            return np.asarray(g(self._x1, self._x2), dtype=np.float64)
        return np.full((self.n + 2, self.n + 2), float(g))

    def assemble_lhs(self, basis=None, coeff="mean", withDirichletBC=True):
        if coeff not in self._lhs:
            vals = fd_values(self._nodal[coeff], self._valid).numpy()
            N = self.n * self.n
            self._lhs[coeff] = sps.csr_matrix((vals, self._colidx, self._rowptr), shape=(N, N))
        return self._lhs[coeff]

    def assemble_operator(self, basis, coeff="mean", withDirichletBC=True):
        return self._la.ScipyOperator(self.assemble_lhs(basis, coeff), basis, basis)

    assemble_operator_inner_dofs = assemble_operator

    def assemble_solve_operator(self, basis, coeff="mean", withDirichletBC=True):
        return self._la.ScipySolveOperator(self.assemble_lhs(basis, coeff), basis, basis)

    def assemble_rhs(self, basis=None, coeff="mean", withDirichletBC=True, withNeumannBC=True, f=None):
        """h^2 f at the interior nodes plus, with Dirichlet conditions, the lifted boundary values
        (there is no Neumann boundary in this model problem)."""
        a = self._nodal[coeff].numpy()
        F = (self.h ** 2) * self._on_nodes(self.f if f is None else f)[1:-1, 1:-1].copy()
        if withDirichletBC and self.uD is not None:
            g = self._on_nodes(self.uD)
            c = a[1:-1, 1:-1]
            F[0, :] += 0.5 * (c[0, :] + a[0, 1:-1]) * g[0, 1:-1]
            F[-1, :] += 0.5 * (c[-1, :] + a[-1, 1:-1]) * g[-1, 1:-1]
            F[:, 0] += 0.5 * (c[:, 0] + a[1:-1, 0]) * g[1:-1, 0]
            F[:, -1] += 0.5 * (c[:, -1] + a[1:-1, -1]) * g[1:-1, -1]
        return F.reshape(-1)

    def set_dirichlet_bc_entries(self, u, homogeneous=False):
        return None

    def norm(self, v, kind="L2"):
        v = np.asarray(v)
        if kind == "L2":
            return float(self.h * np.linalg.norm(v))
        if kind == "l2":
            return float(np.linalg.norm(v))
        raise ValueError("unknown norm %r" % (kind,))

    @property
    def energy_norm(self):
        K0 = self.assemble_lhs(None, "mean")
        return lambda v: float(np.sqrt(max(float(np.dot(v, K0 @ v)), 0.0)))

    def num_cells(self, basis=None):
        return (self.n + 1) ** 2

    # what a slab-native caller may use instead of per-column norm calls
    @property
    def l2_weight(self):
        return self.h ** 2


class GridProlongation:
    """Bilinear interpolation from the n x n interior grid to the (2n+1) x (2n+1) one (homogeneous Dirichlet
    boundary): the ``prolongate`` callable a basis' ``refine`` returns in the reference
    (fem/fenics/fenics_basis.py), here with a matrix the slab path can apply to all columns at once."""

    def __init__(self, coarse, fine):
        self.coarse, self.fine = coarse, fine
        n, nf = coarse.n, fine.n
        one = sps.lil_matrix((nf, n))
        for i in range(n):
            one[2 * i + 1, i] = 1.0
            one[2 * i, i] = 0.5
            one[2 * i + 2, i] = 0.5
        one = sps.csr_matrix(one)
        self._mat = sps.csr_matrix(sps.kron(one, one))          # rows iy*nf + ix, columns jy*n + jx

    def as_csr(self):
        return self._mat

    def __call__(self, vec):
        return type(vec)(self._mat @ np.asarray(vec.coeffs, dtype=np.float64), self.fine)


class GridBasis:
    """Nodal basis of the n x n interior grid (stand-in for the reference's FEniCSBasis as far as
    ``MultiVectorSharedBasis.refine`` needs it): ``refine(cell_ids)`` -> (new_basis, prolongate, restrict)
    with uniform refinement (cell_ids are ignored: every cell is halved)."""

    def __init__(self, n):
        self.n = int(n)

    @property
    def dim(self):
        return self.n * self.n

    def __eq__(self, other):
        return type(self) is type(other) and self.n == other.n

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def copy(self):
        return GridBasis(self.n)

    def refine(self, cell_ids=None):
        fine = GridBasis(2 * self.n + 1)
        P = GridProlongation(self, fine)
        R = P.as_csr().T.tocsr() * 0.25
        return fine, P, (lambda vec: type(vec)(R @ np.asarray(vec.coeffs, dtype=np.float64), self))
