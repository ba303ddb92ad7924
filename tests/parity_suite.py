"""Parity checks of the product path against the oracle.  The SAME functions run on the host-
emulated build of the CUDA sources (tests/test_emu_parity.py, no GPU) and on the real library on a
B200 (tests/test_gpu_parity.py, -m gpu); which one is decided by the fixture that selected the
backend before the call.  Tolerance: 1e-12 relative in fp64 (BASELINE.json north_star)."""
import logging

import numpy as np
import pytest
import scipy.sparse as sps
import torch

import spuq_b200 as sp
from oracle import linalg as olin
from oracle import multiindex as omi
from oracle import polys as opolys
from oracle import sg as osg
from spuq_b200 import _abi, synthetic
from spuq_b200.multi_operator import DeviceBlocks, SGPlan
from tests.cases import fd_case, oracle_multivector, oracle_operator, random_sparse_case, relerr

TOL = 1e-12


def _dev(t):
    return t.to(_abi.device())


def device_blocks(c):
    if hasattr(c, "rowptr"):
        return DeviceBlocks(_dev(c.rowptr), _dev(c.colidx), _dev(c.vals), (c.N, c.N))
    return DeviceBlocks.from_scipy(c.mats)


def product_apply(c, path=_abi.PATH_AUTO, tune=None, m_range=None, include_mean=True):
    A = sp.MultiOperator.from_blocks(device_blocks(c), c.rvs, path=path, m_range=m_range,
                                     include_mean=include_mean)
    w = sp.SlabMultiVector.from_array(c.idx, c.W)
    before = w.slab.clone()
    if tune:
        A.plan_for(w).tune(*tune)
    y = A.apply(w)
    assert torch.equal(before, w.slab), "apply modified its argument"
    assert y is not w and y.slab.data_ptr() != w.slab.data_ptr()
    return y.to_array(), A.plan_for(w)


# ---- apply -------------------------------------------------------------------------------------
def check_apply_paths(n=12, M=4, p=2, kind="legendre", mu=0.0):
    idx = synthetic.complete_order_indices(M, p)
    c = fd_case(n, M, idx, kind, mu)
    seen = set()
    for path in (_abi.PATH_CSR, _abi.PATH_ELL, _abi.PATH_STENCIL, _abi.PATH_AUTO):
        Y, plan = product_apply(c, path)
        assert relerr(Y, c.Y) <= TOL, (path, relerr(Y, c.Y))
        seen.add(plan.kernel_name)
    assert {"k_apply_csr", "k_apply_ellblk<5>"} <= seen
    Y, plan = product_apply(c, _abi.PATH_ELL, (1, 0, 0, 0))          # one column at a time (cross-check kernel)
    assert plan.kernel_name == "k_apply_ell<8>" and relerr(Y, c.Y) <= TOL
    for tune in [(1, 0, 0, 0), (2, 8, 2, 0), (2, 4, 2, 0), (2, 16, 2, 0), (2, 8, 1, 0), (2, 16, 1, 0), (2, 4, 4, 0),
                 (4, 1, 0, 0), (4, 2, 0, 0), (4, 1, 1, 0), (4, 2, 1, 0), (4, 3, 1, 0), (4, 2, 3, 0)]:
        Y, plan = product_apply(c, _abi.PATH_STENCIL, tune)
        assert relerr(Y, c.Y) <= TOL, (tune, plan.kernel_name, relerr(Y, c.Y))
    info = plan.info()
    L, N, nnz = idx.shape[0], c.N, c.colidx.numel()
    assert info["nnz"] == nnz and info["path"] == "stencil"
    up, dn = sp.lambda_tables.neighbour_tables(idx)
    E2 = int((up >= 0).sum() + (dn >= 0).sum())
    D = L * M if (kind == "hermite" and mu != 0.0) else 0
    assert info["n_terms"] == L + E2 + D
    assert info["algorithmic_flops"] == 2 * nnz * (L + E2 + D)
    assert info["algorithmic_bytes"] == 16 * N * L + 8 * nnz * (M + 1) + 4 * nnz + 4 * (N + 1) + M * L * 32


def check_apply_c1_dense_callbacks():
    """Config C1: 32x32 grid, M=5, TD-2 (21 indices), blocks handed over as dense MatrixOperators
    through the reference constructor MultiOperator(coeff_field, assemble_0, assemble_m)."""
    n, M = 32, 5
    idx = synthetic.complete_order_indices(M, 2)
    c = fd_case(n, M, idx, "legendre", dense=True)
    basis = sp.CanonicalBasis(c.N)
    calls = []

    def assemble(b, f):
        calls.append(f)
        k = 0 if f == "mean" else 1 + f
        return sp.MatrixOperator(c.mats[k].toarray(), b, b)

    cf = sp.ListCoefficientField("mean", list(range(M)), [sp.UniformRV(-1, 1)] * M)
    A = sp.MultiOperator(cf, assemble)
    w = sp.MultiVector()
    for k, row in enumerate(idx):
        w[sp.Multiindex(row)] = sp.FlatVector(c.W[:, k].copy(), basis)
    y = A * w                                           # host container in -> host container out
    assert type(y) is sp.MultiVector and len(calls) == M + 1      # one assembly per block
    Y = np.stack([y[mu].coeffs for mu in y.active_indices()], axis=1)
    assert relerr(Y, c.Y) <= TOL
    y2 = A(sp.SlabMultiVector.from_multivector(w))      # slab in -> slab out, call syntax
    assert isinstance(y2, sp.SlabMultiVector) and relerr(y2.to_array(), c.Y) <= TOL
    assert len(calls) == M + 1                          # blocks and plan were cached
    with pytest.raises(TypeError):
        A.apply(sp.FlatVector(np.zeros(3)))
    with pytest.raises(TypeError):
        sp.MultiOperator(3, assemble)
    with pytest.raises(TypeError):
        sp.MultiOperator(cf, 7)


def check_apply_reference_known_answer():
    """application/egsz/tests/test_multi_operator.py:57-90 through the product path."""
    N = 4
    consts = {"mean": 2.0, 0: 3.0, 1: 4.0}
    cf = sp.ListCoefficientField("mean", [0, 1], [sp.UniformRV(), sp.NormalRV(mu=0.5)])

    def diag_assemble(basis, func):
        return sp.DiagonalMatrixOperator(np.full(basis.dim, abs(consts[func])), basis, basis)

    A = sp.MultiOperator(cf, diag_assemble)
    mis = [sp.Multiindex([0]), sp.Multiindex([1]), sp.Multiindex([0, 1]), sp.Multiindex([0, 2])]
    rng = np.random.RandomState(5)
    basis = sp.CanonicalBasis(N)
    vecs = [sp.FlatVector(rng.random_sample(N), basis) for _ in mis]
    w = sp.MultiVector()
    for mi, v in zip(mis, vecs):
        w[mi] = v
    v = A * w
    L = opolys.LegendrePolynomials(normalised=True)
    H = opolys.StochasticHermitePolynomials(mu=0.5, normalised=True)
    c = [x.coeffs for x in vecs]
    v0 = 2 * c[0] + 3 * (L.get_beta(0)[1] * c[1] - L.get_beta(0)[0] * c[0]) + 4 * (H.get_beta(0)[1] * c[2] - H.get_beta(0)[0] * c[0])
    v2 = 2 * c[2] + 4 * (H.get_beta(1)[1] * c[3] - H.get_beta(1)[0] * c[2] + H.get_beta(1)[-1] * c[0])
    assert np.allclose(v[mis[0]].coeffs, v0, rtol=1e-14, atol=0)
    assert np.allclose(v[mis[2]].coeffs, v2, rtol=1e-14, atol=0)


def check_apply_unstructured():
    idx = synthetic.anisotropic_indices(4, 6.0)
    c = random_sparse_case(57, 4, idx)                 # different pattern per block, beta0 != 0
    for path in (_abi.PATH_CSR, _abi.PATH_ELL, _abi.PATH_AUTO):
        Y, plan = product_apply(c, path)
        assert relerr(Y, c.Y) <= TOL, (path, relerr(Y, c.Y))
    assert plan.info()["path"] == "ell"
    with pytest.raises(_abi.SpuqError, match="not a 3x3-window grid stencil"):
        product_apply(c, _abi.PATH_STENCIL)
    # the SG operator on an "unstructured" numbering: the grid blocks of a diffusion problem with the dofs
    # renumbered tile by tile (oracle/synth.py), which the stencil recognition must reject; column-blocked
    # general kernel (row widths 5 -> KW = 8) against the oracle, odd row count, beta0 != 0
    from oracle import synth
    n, M = 13, 5
    idx = synthetic.anisotropic_indices(M, 7.0)
    perm = synth.tile_permutation(n, 4)
    mats = synth.permuted(synth.fd_matrices(n, M), perm)
    from tests.cases import Case, oracle_rv
    c = Case()
    c.N, c.M, c.idx, c.mats = n * n, M, idx, mats
    rng = np.random.RandomState(11)
    c.W = rng.random_sample((c.N, idx.shape[0]))
    c.rvs = synthetic.make_rvs("hermite", M, 0.5)
    c.A, c.cf = oracle_operator(mats, [oracle_rv("hermite", 0.5)] * M)
    c.w = oracle_multivector(idx, c.W)
    c.Y = (c.A * c.w).to_slab()
    Y, plan = product_apply(c)
    assert plan.kernel_name == "k_apply_ellblk<5>" and relerr(Y, c.Y) <= TOL, (plan.kernel_name, relerr(Y, c.Y))
    # wider rows (random sparse blocks, up to 16 entries per row): KW = 16
    c2 = random_sparse_case(40, 3, synthetic.anisotropic_indices(3, 5.0), density=0.2, seed=9)
    Y, plan = product_apply(c2, _abi.PATH_ELL)
    assert relerr(Y, c2.Y) <= TOL, (plan.kernel_name, relerr(Y, c2.Y))


def nine_point_case(nx, ny, M, idx, seed=3):
    """Random 9-point-window blocks on an nx x ny grid (non-square grid, full 3x3 coupling)."""
    rng = np.random.RandomState(seed)
    N = nx * ny
    rows, cols = [], []
    for iy in range(ny):
        for ix in range(nx):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    if 0 <= ix + dx < nx and 0 <= iy + dy < ny:
                        rows.append(iy * nx + ix); cols.append((iy + dy) * nx + ix + dx)
    mats = [sps.csr_matrix((rng.standard_normal(len(rows)), (rows, cols)), shape=(N, N)) for _ in range(M + 1)]
    from tests.cases import Case
    c = Case()
    c.N, c.M, c.idx, c.mats = N, M, idx, mats
    c.W = rng.random_sample((N, idx.shape[0]))
    c.rvs = synthetic.make_rvs("hermite", M, 0.5)
    c.orvs = [opolys.NormalRV(0.5, 1)] * M
    c.A, c.cf = oracle_operator(mats, c.orvs)
    c.w = oracle_multivector(idx, c.W)
    c.Y = (c.A * c.w).to_slab()
    return c


def check_apply_nine_point_and_odd_shapes():
    idx = synthetic.complete_order_indices(3, 2)
    c = nine_point_case(10, 7, 3, idx)
    for tune in [None, (1, 0, 0, 0), (2, 8, 2, 0), (2, 4, 2, 0), (2, 8, 1, 0), (4, 1, 0, 0), (4, 1, 1, 0), (4, 2, 1, 0)]:
        Y, plan = product_apply(c, _abi.PATH_STENCIL, tune)
        assert relerr(Y, c.Y) <= TOL, (tune, plan.kernel_name)
    assert "9pt" in product_apply(c)[1].kernel_name
    # odd grid width: vector loads impossible -> the simple stencil kernel is chosen
    c = nine_point_case(7, 5, 2, synthetic.complete_order_indices(2, 2))
    Y, plan = product_apply(c)
    assert plan.kernel_name == "k_apply_stencil_simple" and relerr(Y, c.Y) <= TOL
    with pytest.raises(_abi.SpuqError, match="even grid width"):
        product_apply(c, _abi.PATH_STENCIL, (2, 8, 2, 0))
    # 1-D problem (tridiagonal blocks): one grid row
    N, M = 40, 2
    idx = synthetic.complete_order_indices(M, 3)
    rng = np.random.RandomState(9)
    mats = [sps.diags([rng.standard_normal(N - 1), rng.standard_normal(N), rng.standard_normal(N - 1)],
                      [-1, 0, 1], format="csr") for _ in range(M + 1)]
    from tests.cases import Case
    c = Case(); c.N, c.M, c.idx, c.mats = N, M, idx, mats
    c.W = rng.random_sample((N, idx.shape[0]))
    c.rvs = synthetic.make_rvs("legendre", M); c.orvs = [opolys.UniformRV(-1, 1)] * M
    c.A, _ = oracle_operator(mats, c.orvs); c.w = oracle_multivector(idx, c.W); c.Y = (c.A * c.w).to_slab()
    for tune in [None, (1, 0, 0, 0), (2, 8, 2, 0)]:
        Y, plan = product_apply(c, _abi.PATH_AUTO, tune)
        assert plan.info()["path"] == "stencil" and relerr(Y, c.Y) <= TOL


def check_apply_edge_cases():
    # |Lambda| = 1 (mean only), tiny grid, L not a multiple of the column block
    c = fd_case(4, 0, np.zeros((1, 1), dtype=np.int32), "legendre")
    c.rvs = []
    Y, plan = product_apply(c)
    assert relerr(Y, c.Y) <= TOL and plan.info()["n_terms"] == 1
    idx = synthetic.complete_order_indices(2, 3)          # L = 10
    c = fd_case(6, 2, idx, "hermite", 0.5)
    for tune in [(2, 4, 2, 0), (2, 8, 2, 0), (2, 16, 1, 0)]:
        assert relerr(product_apply(c, _abi.PATH_STENCIL, tune)[0], c.Y) <= TOL
    # a set with missing neighbours (not downward closed) and unsorted input order
    idx = np.array([[0, 2], [0, 0], [1, 0], [0, 1], [3, 1]], dtype=np.int32)
    W = np.random.RandomState(2).random_sample((36, 5))
    c = fd_case(6, 2, idx, "hermite", 0.5, with_oracle=False)
    c.W = W
    c.orvs = [opolys.NormalRV(0.5, 1)] * 2
    c.A, _ = oracle_operator(c.mats, c.orvs)
    c.w = oracle_multivector(idx, W)
    c.Y = (c.A * c.w).to_slab()
    assert relerr(product_apply(c)[0], c.Y) <= TOL
    # coefficient field shorter than the longest index: maxm is capped (multi_operator2.py:45-48)
    idx = synthetic.complete_order_indices(3, 2)
    c = fd_case(5, 3, idx, "legendre", with_oracle=False)
    c.orvs = [opolys.UniformRV(-1, 1)] * 2
    c.A, _ = oracle_operator(c.mats[:3], c.orvs)
    c.w = oracle_multivector(idx, c.W)
    c.Y = (c.A * c.w).to_slab()
    blocks = device_blocks(c)
    short = DeviceBlocks(blocks.rowptr, blocks.colidx, blocks.vals[:3].contiguous(), blocks.shape)
    A = sp.MultiOperator.from_blocks(short, c.rvs[:2])
    y = A.apply(sp.SlabMultiVector.from_array(idx, c.W))
    assert relerr(y.to_array(), c.Y) <= TOL


def check_apply_m_shards():
    """M-sharded plans (one per rank in the multi-GPU path) sum to the full operator."""
    idx = synthetic.complete_order_indices(5, 2)
    c = fd_case(8, 5, idx, "hermite", 0.5)
    total = np.zeros_like(c.Y)
    for g, (lo, hi) in enumerate([(0, 2), (2, 3), (3, 5)]):
        Y, _ = product_apply(c, m_range=(lo, hi), include_mean=(g == 0))
        total += Y
    assert relerr(total, c.Y) <= TOL


def check_apply_multiblock():
    """More than one 64-column block: the TMEM-kernel schedule reorders Lambda (reconstructed from
    the neighbour tables), cuts it into blocks and register sets; anisotropic and isotropic sets,
    beta0 != 0, and M-shards that leave some output columns without any term."""
    cases = [(8, synthetic.complete_order_indices(6, 3), 6, "legendre", 0.0),          # L = 84
             (6, synthetic.anisotropic_indices(8, 9.0), 8, "hermite", 0.5),
             (6, synthetic.truncated_complete_indices(7, 3, 100), 7, "hermite", 0.0)]
    for n, idx, M, kind, mu in cases:
        assert idx.shape[0] > 64
        c = fd_case(n, M, idx, kind, mu)
        for tune in [(4, 1, 0, 0), (4, 3, 1, 0), (4, 2, 1, 0), (4, 2, 3, 0), None]:
            Y, plan = product_apply(c, _abi.PATH_STENCIL, tune)
            assert relerr(Y, c.Y) <= TOL, (tune, plan.kernel_name, relerr(Y, c.Y))
        total = np.zeros_like(c.Y)
        shards = [(0, 1), (1, 3), (3, M)]
        for g, (lo, hi) in enumerate(shards):
            Y, _ = product_apply(c, _abi.PATH_STENCIL, (4, 2, 0, 0), m_range=(lo, hi), include_mean=(g == 1))
            total += Y
        assert relerr(total, c.Y) <= TOL


def check_refine_and_table_growth():
    """f3: (i) MultiVectorSharedBasis.refine on the slab (one rectangular sparse product for all columns)
    against the oracle's per-index prolongation; (ii) the neighbour tables of a grown Lambda obtained by
    updating the old ones, bit-identical to a rebuild and to the oracle's look-ups."""
    from spuq_b200.lambda_tables import LambdaTables
    n, M = 7, 3
    idx = synthetic.anisotropic_indices(M, 6.0)
    rng = np.random.RandomState(4)
    W = rng.random_sample((n * n, idx.shape[0]))
    basis = synthetic.GridBasis(n)
    w = sp.SlabMultiVector.from_array(idx, W, basis=basis)
    fine = w.refine()
    assert isinstance(fine, sp.SlabMultiVector) and fine.spatial_basis == synthetic.GridBasis(2 * n + 1)
    w_o = oracle_multivector(idx, W, basis=basis)                  # the oracle works on the same basis object
    fine_o = osg.refine_multivector(w_o)
    assert np.max(np.abs(fine.to_array() - fine_o.to_slab())) <= 1e-15
    assert fine.active_indices() == [sp.Multiindex(np.asarray(m.as_array)) for m in fine_o.active_indices()]
    host = w.to_multivector(sp.MultiVectorSharedBasis).refine()     # host container: per-index path
    assert np.max(np.abs(sp.SlabMultiVector.from_multivector(host).to_array() - fine_o.to_slab())) <= 1e-15
    # (ii) growth: add the forward neighbours of a few indices, one of them twice, some already active
    rvs = synthetic.make_rvs("hermite", M, 0.5)
    t0 = LambdaTables(idx, rvs)
    have = {tuple(r) for r in idx.tolist()}
    add = []
    for r in idx[::3]:
        for m in range(M):
            c = list(r); c[m] += 1
            if tuple(c) not in have:
                have.add(tuple(c)); add.append(c)
    assert len(add) >= 3
    grown_idx = np.vstack([idx, np.array(add, dtype=np.int32)])
    t1 = t0.grown(grown_idx, rvs)
    ref = LambdaTables(grown_idx, rvs)
    assert t1 is not None
    for name in ("index_array", "nbr_up", "nbr_dn", "beta_up", "beta_dn", "beta_0", "perm"):
        assert np.array_equal(getattr(t1, name), getattr(ref, name)), name
    Lam = sorted(omi.Multiindex(np.asarray(r, dtype=np.int64)) for r in grown_idx)
    up_o, dn_o = osg.neighbour_tables(Lam, ref.maxm)
    assert np.array_equal(t1.nbr_up, up_o) and np.array_equal(t1.nbr_dn, dn_o)
    assert t0.grown(idx[1:], rvs) is None                           # not a superset -> rebuild
    # through the operator: a plan for the grown vector after one for the original (uses the update)
    c = fd_case(n, M, idx, "hermite", 0.5)
    A = sp.MultiOperator.from_blocks(device_blocks(c), c.rvs)
    ws = sp.SlabMultiVector.from_array(idx, c.W)
    A.apply(ws)
    sp.Marking.refine_y(ws, [sp.Multiindex(np.asarray(a)) for a in add])
    y = A.apply(ws)
    cg = fd_case(n, M, np.ascontiguousarray(ref.index_array), "hermite", 0.5, with_oracle=False)
    Wg = ws.to_array()
    wg_o = oracle_multivector(ref.index_array, Wg)
    A_o, _ = oracle_operator(c.mats, c.orvs)
    Yg = (A_o * wg_o).to_slab()
    assert relerr(y.to_array()[:, np.linalg.norm(Yg, axis=0) > 0], Yg[:, np.linalg.norm(Yg, axis=0) > 0]) <= TOL


def check_upper_tail_bound():
    """evaluateUpperTailBound (residual_estimator2.py:176-290) on the slab against the oracle's restatement:
    same boundary indices with the same multiplicities, zeta / zeta_bar / global_zeta to 1e-12."""
    from spuq_b200.residual import evaluate_upper_tail_bound, lambda_boundary
    for kind, mu_rv, T in (("hermite", 0.5, 6.0), ("legendre", 0.0, 7.5)):
        M = 4
        idx = synthetic.anisotropic_indices(M, T)
        c = fd_case(8, M, idx, kind, mu_rv)
        K0 = c.mats[0]
        ainfty = lambda m: 0.9 / (m + 2.0) ** 2                       # noqa: E731
        n_terms = M + 3                                               # a field longer than the vector's order
        rv_o = c.orvs[0]
        cf_o = osg.ListCoefficientField("mean", list(range(n_terms)), [rv_o] * n_terms)
        energy_o = lambda v: float(np.sqrt(max(float(np.dot(v.coeffs, K0 @ v.coeffs)), 0.0)))      # noqa: E731
        g_o, zeta_o, zbar_o, zm_o = osg.evaluate_upper_tail_bound(c.w, cf_o, energy_o, ainfty, add_maxm=2)
        cf_p = sp.ListCoefficientField("mean", list(range(n_terms)), synthetic.make_rvs(kind, n_terms, mu_rv))
        w = sp.SlabMultiVector.from_array(idx, c.W)
        g_p, zeta_p, zbar_p, zm_p = evaluate_upper_tail_bound(w, cf_p, K0, ainfty, add_maxm=2)
        width = max(len(k.as_array) for k in zeta_o) if zeta_o else 0
        as_t = lambda k, wd: tuple(int(v) for v in np.pad(np.asarray(k.as_array), (0, max(0, wd - len(k.as_array)))))  # noqa: E731
        wd = max(width, idx.shape[1], n_terms)
        zo = {as_t(k, wd): v for k, v in zeta_o.items()}
        zp = {as_t(k, wd): v for k, v in zeta_p.items()}
        assert set(zo) == set(zp) and len(zo) > 0
        for k in zo:
            assert abs(zo[k] - zp[k]) <= 1e-12 * max(abs(zo[k]), 1e-300), (k, zo[k], zp[k])
        bo = {as_t(k, wd): v for k, v in zbar_o.items()}
        bp = {as_t(k, wd): v for k, v in zbar_p.items()}
        assert set(bo) == set(bp)
        for k in bo:
            assert abs(bo[k] - bp[k]) <= 1e-12 * max(abs(bo[k]), 1e-300), (k, bo[k], bp[k])
        assert abs(g_o - g_p) <= 1e-12 * g_o
        mu_o = c.w.active_indices()[1]
        assert abs(zm_o(mu_o, 0) - zm_p(sp.Multiindex(np.asarray(mu_o.as_array)), 0)) <= 1e-12 * abs(zm_o(mu_o, 0))
        # multiplicities of the boundary generator
        assert any(v > 1 for v in lambda_boundary(idx).values())


def check_row_bands_manual_exchange(n=24, M=3, bands=((0, 7), (7, 16), (16, 24))):
    """The single-GPU kernel on row bands with halo rows, as RowShardedApply runs it, with all ranks
    simulated in one process (halo rows copied by hand): bands of odd and even height that are not
    multiples of the tile height must reproduce the global result."""
    from spuq_b200.sharding import RowShardedApply
    idx = synthetic.complete_order_indices(M, 2)
    c = fd_case(n, M, idx, "hermite", 0.5)
    L = idx.shape[0]
    Wg = torch.from_numpy(np.ascontiguousarray(c.W.T)).view(L, n, n)
    Yg = np.zeros((L, n, n))
    for y0, y1 in bands:
        nyl = y1 - y0 + 2
        rowptr, colidx, vals = synthetic.fd_blocks_rows(n, M, y0, y1)
        blocks = DeviceBlocks(_dev(rowptr), _dev(colidx), _dev(vals), (nyl * n, nyl * n))
        sh = RowShardedApply(blocks, c.rvs, n, y0, y1)
        W = torch.zeros((L, nyl, n), dtype=torch.float64)
        lo, hi = max(y0 - 1, 0), min(y1 + 1, n)                     # owned rows + the neighbours' boundary rows
        W[:, lo - (y0 - 1):hi - (y0 - 1), :] = Wg[:, lo:hi, :]
        W = _dev(W.view(L, nyl * n))
        w = sp.SlabMultiVector(idx, slab=W, _sorted=True)
        plan = sh.plan_for(w)
        Y = torch.empty_like(W)
        plan.apply(W, Y)
        Yg[:, y0:y1, :] = Y.view(L, nyl, n)[:, 1:-1, :].cpu().numpy()
    assert relerr(Yg.reshape(L, n * n).T, c.Y) <= TOL


def check_partitioned_mean_solver():
    """The partition method for the mean solve on row-sharded slabs (PartitionedMeanSolver), all ranks
    simulated in one process: forward halves, the all-gather done by hand, corrections and inverse
    halves must reproduce the global sparse direct solve; uneven partitions, even and odd nx."""
    import scipy.sparse.linalg as spla
    from spuq_b200.sharding import PartitionedMeanSolver
    lib = _abi.lib()
    rng = np.random.RandomState(3)
    for nx, ny, ranges, L in [(6, 11, [(0, 4), (4, 7), (7, 11)], 3), (9, 8, [(0, 4), (4, 8)], 2), (16, 16, [(0, 16)], 2)]:
        T = lambda n, o, dd: sps.diags([o, dd, o], [-1, 0, 1], shape=(n, n))      # noqa: E731
        A0 = sps.csr_matrix(sps.kron(sps.identity(ny), T(nx, -1.0, 2.0)) + sps.kron(T(ny, -1.0, 2.0), sps.identity(nx)))
        R = rng.random_sample((L, ny * nx))
        lu = spla.splu(A0.tocsc())
        ref = np.stack([lu.solve(R[k]) for k in range(L)])
        solvers = [PartitionedMeanSolver(nx, ny, ranges, p, (4.0, -1.0, -1.0)) for p in range(len(ranges))]
        Ts = []
        for s_, (r0, r1) in zip(solvers, ranges):
            Rc = _dev(torch.from_numpy(np.ascontiguousarray(R.reshape(L, ny, nx)[:, r0:r1, :]).reshape(L, -1)))
            Tt = torch.empty_like(Rc)
            work = torch.empty(int(lib.spuq_sg_precond_work_bytes(s_.local.handle, Rc.shape[1], L)) // 8 + 1,
                               dtype=torch.float64, device=Rc.device)
            _abi.check(lib.spuq_sg_precond_mean_fd_forward(s_.local.handle, Rc.shape[1], L, _abi.ptr(Rc), _abi.ptr(Tt),
                                                           _abi.ptr(work), _abi.stream_ptr()), "forward")
            Ts.append(Tt)
        g = torch.cat([torch.stack([t.view(L, r1 - r0, nx)[:, 0, :], t.view(L, r1 - r0, nx)[:, -1, :]])
                       for t, (r0, r1) in zip(Ts, ranges)])
        out = np.zeros((L, ny, nx))
        for s_, Tt, (r0, r1) in zip(solvers, Ts, ranges):
            zl = torch.einsum("jk,jlk->lk", s_.cz, g).contiguous()
            zr = torch.einsum("jk,jlk->lk", s_.ca, g).contiguous()
            _abi.check(lib.spuq_sg_spike_correct(L, r1 - r0, nx, _abi.ptr(Tt), _abi.ptr(zl), _abi.ptr(zr), _abi.ptr(s_.sl),
                                                 _abi.ptr(s_.sr), _abi.stream_ptr()), "correct")
            Sc = torch.empty_like(Tt)
            _abi.check(lib.spuq_sg_precond_mean_fd_inverse(s_.local.handle, Tt.shape[1], L, _abi.ptr(Tt), _abi.ptr(Sc),
                                                           _abi.stream_ptr()), "inverse")
            out[:, r0:r1, :] = Sc.view(L, r1 - r0, nx).cpu().numpy()
        err = np.max(np.abs(out.reshape(L, -1) - ref)) / np.max(np.abs(ref))
        assert err <= 1e-12, (nx, ny, ranges, err)
    with pytest.raises(ValueError):
        PartitionedMeanSolver(6, 5, [(0, 4), (4, 5)], 0, (4.0, -1.0, -1.0))


def check_row_sharded_pcg_single_rank(n=16, M=3, p=2):
    """RowShardedPCG with one rank (no process group): the halo'd local slab is the whole grid plus two
    empty rows; the solve must agree with the oracle's PCG on the global problem (the two-rank form
    of this check runs on gloo in tests/test_sharded_gloo.py)."""
    from spuq_b200.sharding import RowShardedApply, RowShardedPCG
    idx = synthetic.complete_order_indices(M, p)
    c = fd_case(n, M, idx, "legendre")
    L = idx.shape[0]
    rowptr, colidx, vals = synthetic.fd_blocks_rows(n, M, 0, n)
    nyl = n + 2
    blocks = DeviceBlocks(_dev(rowptr), _dev(colidx), _dev(vals), (nyl * n, nyl * n))
    sh = RowShardedApply(blocks, c.rvs, n, 0, n)
    F = torch.zeros((L, nyl, n), dtype=torch.float64)
    F[:, 1:-1, :] = torch.from_numpy(np.ascontiguousarray(c.Y.T)).view(L, n, n)
    F = _dev(F.view(L, nyl * n))
    w = sp.SlabMultiVector(idx, slab=torch.zeros_like(F), _sorted=True)
    plan = sh.plan_for(w)
    d = float(c.mats[0].diagonal()[0])
    solver = RowShardedPCG(sh, n, n, [(0, n)], mean_stencil=(d, -1.0, -1.0))
    X, zeta, it = solver.solve(plan, F, torch.zeros_like(F), eps=1e-9, maxiter=100)
    basis = olin.CanonicalBasis(c.N)
    lu = olin.ScipySolveOperator(c.mats[0], basis, basis, factor=True)
    oh = []
    x_o, zeta_o, it_o = osg.pcg(c.A, c.A * c.w, osg.PreconditioningOperator("mean", lambda b, f: lu), 0 * c.w,
                                eps=1e-9, maxiter=100, history=oh)
    assert abs(it - it_o) <= 1
    k = min(len(oh), len(solver.last_history), 6)
    assert np.allclose(solver.last_history[:k], oh[:k], rtol=1e-9)
    Xg = X.view(L, nyl, n)[:, 1:-1, :].reshape(L, n * n).cpu().numpy().T
    assert np.max(np.abs(Xg - x_o.to_slab())) <= 1e-8
    with pytest.raises(Exception, match="PCG did not converge"):
        solver.solve(plan, F, torch.zeros_like(F), eps=1e-14, maxiter=3)


# ---- the caller of pcg: right-hand side and solve --------------------------
def check_pcg_solve(n=10, M=3, p=2, kind="hermite", mu=0.5):
    """prepare_rhs / pcg_solve (adaptive_solver2.py:39-85) against the oracle's restatement on the
    same discretisation object: inhomogeneous Dirichlet data so that every term contributes, a
    non-symmetric random variable so that beta[0] != 0 (the mean column collects all terms)."""
    idx = synthetic.complete_order_indices(M, p)
    c = fd_case(n, M, idx, kind, mu)
    uD = lambda x1, x2: 1.0 + x1 + 0.5 * x2 * x2          # noqa: E731
    load = lambda x1, x2: 1.0 + np.sin(3.0 * x1) * x2     # noqa: E731
    pde_o = synthetic.FDDiscretisation(n, M, olin, f=load, uD=uD)
    pde_p = synthetic.FDDiscretisation(n, M, sp, f=load, uD=uD)
    w0_o = 0 * c.w
    b_o = osg.prepare_rhs(c.A, w0_o, c.cf, pde_o)
    stats_o = {}
    x_o, zeta_o = osg.pcg_solve(c.A, w0_o, c.cf, pde_o, stats_o, 1e-9, 100)

    A = sp.MultiOperator.from_blocks(device_blocks(c), c.rvs)
    cf = sp.ListCoefficientField("mean", list(range(M)), c.rvs)
    w0 = sp.SlabMultiVector.from_array(idx, np.zeros_like(c.W))
    b = sp.prepare_rhs(A, w0, cf, pde_p)
    assert isinstance(b, sp.SlabMultiVector) and b.active_indices() == w0.active_indices()
    B, B_o = b.to_array(), b_o.to_slab()
    assert np.max(np.abs(B - B_o)) <= 1e-14 * np.max(np.abs(B_o))
    nonzero = [k for k in range(B.shape[1]) if np.any(B[:, k] != 0.0)]
    assert len(nonzero) == M + 1                          # the mean column and one column per e_m
    stats = {}
    x, zeta = sp.pcg_solve(A, w0, cf, pde_p, stats, 1e-9, 100)
    assert isinstance(x, sp.SlabMultiVector)
    assert np.max(np.abs(x.to_array() - x_o.to_slab())) <= 1e-8 * np.max(np.abs(x_o.to_slab()))
    assert stats["DOFS"] == stats_o["DOFS"] == c.N * idx.shape[0]
    assert stats["CELLS"] == stats_o["CELLS"] == (n + 1) ** 2 * idx.shape[0]
    assert stats["TIME-PCG"] > 0.0 and stats["PCG-ITERATIONS"] >= 2
    for key in ("RESIDUAL-L2", "RESIDUAL-H1A"):
        # residuals of two converged runs: same size, both tiny against the right-hand side
        assert stats[key] <= 1e-6 and stats_o[key] <= 1e-6, (key, stats[key], stats_o[key])
    # the norms themselves, on a residual that is not tiny: one PCG pass only
    s1, s1_o = {}, {}
    with pytest.raises(Exception, match="PCG did not converge"):
        sp.pcg_solve(A, w0, cf, pde_p, s1, 1e-14, 2)
    r_p = sp.error_norm(b, A * (0.5 * b), "L2", pde_p)
    r_o = osg.error_norm(b_o, c.A * (0.5 * b_o), "L2", pde_o)
    assert abs(r_p - r_o) <= 1e-12 * r_o
    from spuq_b200.adaptive_solver import _EnergyNorm
    e_p = sp.error_norm(b, A * (0.5 * b), _EnergyNorm(pde_p.energy_norm, pde_p.assemble_lhs(None, "mean")), pde_p)
    e_o = osg.error_norm(b_o, c.A * (0.5 * b_o), pde_o.energy_norm, pde_o)
    assert abs(e_p - e_o) <= 1e-11 * e_o
    # a dictionary MultiVector in, as the reference's callers hold it
    x2, _ = sp.pcg_solve(A, w0.to_multivector(), cf, pde_p, {}, 1e-9, 100)
    assert np.max(np.abs(x2.to_array() - x.to_array())) <= 1e-12


# ---- neighbour combination of the residual estimator ------------------------
def check_neighbour_combination():
    """res = -beta0 w[mu] + beta1 w[mu+e_m] + beta-1 w[mu-e_m] (residual_estimator2.py:117-131) for all
    columns of a term at once: BIT-identical to the oracle's loop (same operation order, no fused
    multiply-add), for a symmetric and a shifted random variable; per-column energy norms."""
    for n, M, p, kind, mu0 in [(7, 3, 3, "hermite", 0.5), (8, 4, 2, "legendre", 0.0)]:
        idx = synthetic.complete_order_indices(M, p)
        c = fd_case(n, M, idx, kind, mu0)
        cf_p = sp.ListCoefficientField("mean", list(range(M)), c.rvs)
        w = sp.SlabMultiVector.from_array(idx, c.W)
        for m in range(M):
            res = sp.neighbour_combination(w, cf_p, m)
            ref = osg.neighbour_combination(c.w, c.cf, m).to_slab()
            assert np.array_equal(res.to_array(), ref), (kind, m, np.max(np.abs(res.to_array() - ref)))
        with pytest.raises(IndexError):
            sp.neighbour_combination(w, cf_p, M)
        norms = sp.column_energy_norms(w, c.mats[0]).cpu().numpy()
        ref_n = np.sqrt(np.einsum("ik,ik->k", c.W, c.mats[0] @ c.W))
        assert np.max(np.abs(norms - ref_n)) <= 1e-12 * np.max(ref_n)


# ---- growth of Lambda between solves ---------------------------------------
def check_lambda_growth():
    """Marking.refine_y (marking2.py:205-208) on a slab: the grown vector has the oracle's keys in the
    oracle's order, old columns untouched, new ones zero; the operator applied to it (fresh neighbour
    tables for the grown set) matches the oracle; batched insertion == one-by-one insertion."""
    M = 4
    idx_all = synthetic.complete_order_indices(M, 3)
    start = [k for k in range(idx_all.shape[0]) if idx_all[k].sum() <= 2]
    idx0 = idx_all[start]
    rest = [k for k in range(idx_all.shape[0]) if idx_all[k].sum() == 3]
    new_rows = [idx_all[k] for k in rest[::3]]                          # a scattered third of the degree-3 shell
    c = fd_case(8, M, idx0, "hermite", 0.5)
    new_o = [omi.Multiindex(np.asarray(r, dtype=np.int64)) for r in new_rows]
    new_p = [sp.Multiindex(r) for r in new_rows]
    w_o = c.w.copy()
    osg.refine_y(w_o, new_o)
    w = sp.SlabMultiVector.from_array(idx0, c.W)
    before = {mu: w[mu].as_array().copy() for mu in w.active_indices()}
    sp.Marking.refine_y(w, new_p)
    assert len(w) == len(w_o) == idx0.shape[0] + len(new_rows)
    assert [tuple(mu.as_array) for mu in w.active_indices()] == [tuple(mu.as_array) for mu in w_o.active_indices()]
    for mu in w.active_indices():
        col = w[mu].as_array()
        assert np.array_equal(col, before[mu]) if mu in before else not col.any()
    # one by one gives the same slab and the same order
    w1 = sp.SlabMultiVector.from_array(idx0, c.W)
    for mu in new_p:
        w1[mu] = sp.FlatVector(np.zeros(c.N), w1.spatial_basis)
    assert w1.active_indices() == w.active_indices() and torch.equal(w1.slab, w.slab)
    sp.Marking.refine_y(w, new_p[:2])                                   # already active: nothing changes
    assert torch.equal(w1.slab, w.slab)
    # the operator on the grown set: new columns receive contributions from their neighbours
    A = sp.MultiOperator.from_blocks(device_blocks(c), c.rvs)
    Y = A.apply(w).to_array()
    Y_o = (c.A * w_o).to_slab()
    assert relerr(Y, Y_o) <= TOL
    grown_cols = [k for k, mu in enumerate(w.active_indices()) if mu not in before]
    assert np.abs(Y[:, grown_cols]).max() > 0.0
    # a dictionary MultiVector grows the reference's way
    wd = sp.SlabMultiVector.from_array(idx0, c.W).to_multivector()
    sp.Marking.refine_y(wd, new_p)
    assert len(wd) == len(w)


# ---- evaluation at parameter samples ---------------------------------------
def check_parametric_samples():
    """compute_parametric_sample_solution (sampling.py:59-75) and its many-sample form against the
    oracle's restatement: odd and even n, |Lambda| above and below the kernel's coefficient chunk,
    sample counts that exercise every accumulator width."""
    rng = np.random.RandomState(5)
    for n, M, p, kind, mu, S in [(9, 3, 3, "hermite", 0.5, 11), (8, 4, 2, "legendre", 0.0, 2), (4, 7, 5, "legendre", 0.0, 5),
                                 (5, 2, 2, "hermite", 0.0, 1)]:
        idx = synthetic.complete_order_indices(M, p)
        c = fd_case(n, M, idx, kind, mu)
        cf_p = sp.ListCoefficientField("mean", list(range(M)), c.rvs)
        w = sp.SlabMultiVector.from_array(idx, c.W)
        samples = rng.uniform(-0.9, 0.9, size=(S, M))
        vals = sp.sample_solutions(samples, cf_p, w)
        assert tuple(vals.shape) == (S, c.N)
        vals = vals.cpu().numpy()
        for s_ in range(S):
            ref = osg.compute_parametric_sample_solution(list(samples[s_]), c.cf, c.w)
            assert np.max(np.abs(vals[s_] - ref.coeffs)) <= 1e-12 * max(1.0, np.max(np.abs(ref.coeffs))), (n, M, p, s_)
        one = sp.compute_parametric_sample_solution(list(samples[0]), cf_p, w)
        assert isinstance(one, sp.FlatVector) and np.array_equal(one.coeffs, vals[0])
        # the sample map itself: bit-exact products of the same polynomial values
        sm_p, _ = cf_p.sample_realization(w.active_indices(), list(samples[0]))
        sm_o, _ = c.cf.sample_realization(c.w.active_indices(), list(samples[0]))
        assert [sm_p[k] for k in w.active_indices()] == [sm_o[k] for k in c.w.active_indices()]
    # a dictionary MultiVector in
    one2 = sp.compute_parametric_sample_solution(list(samples[0]), cf_p, w.to_multivector())
    assert np.max(np.abs(one2.coeffs - one.coeffs)) <= 1e-14
    with pytest.raises(TypeError):
        sp.sample_solutions(samples, cf_p, 3)


# ---- vectors ------------------------------------------------------------------------------------
def check_slab_multivector_api():
    FV = sp.FlatVector
    mis = sp.MultiindexSet.createCompleteOrderSet(3, 4)
    a, b, c3 = sp.SlabMultiVector(), sp.SlabMultiVector(), sp.SlabMultiVector()
    a.set_defaults(mis, FV([3, 4, 5])); b.set_defaults(mis, FV([6, 8, 12])); c3.set_defaults(mis, FV([9, 12, 17]))
    assert a + b == c3 and len(a) == 35
    assert a[sp.Multiindex([1, 2, 1])] == FV([3, 4, 5]) and a[sp.Multiindex()] == FV([3, 4, 5])
    with pytest.raises(KeyError):
        a[sp.Multiindex([1, 2, 2])]
    d = sp.SlabMultiVector(); d.set_defaults(mis, FV([6, 8, 10]))
    assert 2 * a == d and a * 2.0 == d and (d - a) == a
    assert (-a)[sp.Multiindex(mis[-1])] == FV([-3, -4, -5])
    keys = a.active_indices()
    assert keys == sorted(keys) and a.max_order == 3
    a2 = a.copy()
    a[sp.Multiindex([1, 2, 1])] = FV([7, 7, 7])
    assert a2[sp.Multiindex([1, 2, 1])] == FV([3, 4, 5]) and a[sp.Multiindex([1, 2, 1])] == FV([7, 7, 7])
    # trailing zeros collide; new keys grow Lambda (zero-initialised column, sorted re-layout)
    mv = sp.SlabMultiVector()
    mv[sp.Multiindex([1, 2, 3])] = FV([3, 4, 5])
    mv[sp.Multiindex([0, 1, 2, 3])] = FV([3, 4, 6])
    mv[sp.Multiindex([1, 2, 3, 0])] = FV([3, 4, 7])
    assert len(mv) == 2 and mv[sp.Multiindex([1, 2, 3])] == FV([3, 4, 7])
    assert [list(k.as_array) for k in mv.active_indices()] == [[0, 1, 2, 3], [1, 2, 3]]
    assert str(mv) == "<SlabMultiVector keys=[<Multiindex inds=[0 1 2 3]>, <Multiindex inds=[1 2 3]>]>"
    # inner product, dict round trip, flatten layout, pickling
    rng = np.random.RandomState(4)
    idx = synthetic.complete_order_indices(3, 2)
    X, Y = rng.random_sample((17, 10)), rng.random_sample((17, 10))
    x, y = sp.SlabMultiVector.from_array(idx, X), sp.SlabMultiVector.from_array(idx, Y)
    assert abs(sp.inner(x, y) - float((X * Y).sum())) <= 1e-13 * abs((X * Y).sum())
    host = x.to_multivector()
    assert type(host) is sp.MultiVector and sp.SlabMultiVector.from_multivector(host) == x
    assert abs(sp.inner(host, y.to_multivector()) - sp.inner(x, y)) < 1e-12
    assert np.array_equal(x.flatten().coeffs, X.T.reshape(-1))         # column-major N x L slab
    import pickle
    assert pickle.loads(pickle.dumps(x)) == x
    with pytest.raises(AssertionError):
        x += sp.SlabMultiVector.from_array(idx[:5], X[:, :5])


def check_blas():
    lib = _abi.lib()
    rng = np.random.RandomState(8)
    for n in (1, 7, 1000, 100003):
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        tx, ty = _dev(torch.from_numpy(x)), _dev(torch.from_numpy(y))
        scratch = torch.zeros(int(lib.spuq_sg_reduce_scratch_bytes()) // 8 + 1, dtype=torch.float64, device=tx.device)
        out = torch.zeros(2, dtype=torch.float64, device=tx.device)
        _abi.check(lib.spuq_sg_dot2(n, _abi.ptr(tx), _abi.ptr(ty), _abi.ptr(out), _abi.ptr(scratch), _abi.stream_ptr()))
        got = out.cpu().numpy()
        assert abs(got[0] - x @ y) <= 1e-12 * max(1.0, np.abs(x * y).sum())
        assert abs(got[1] - x @ x) <= 1e-12 * (x @ x)
        r1 = got.copy()
        _abi.check(lib.spuq_sg_dot2(n, _abi.ptr(tx), _abi.ptr(ty), _abi.ptr(out), _abi.ptr(scratch), _abi.stream_ptr()))
        assert np.array_equal(out.cpu().numpy(), r1)            # deterministic
        num = _dev(torch.tensor([3.0], dtype=torch.float64)); den = _dev(torch.tensor([4.0], dtype=torch.float64))
        w, rho = tx.clone(), ty.clone()
        _abi.check(lib.spuq_sg_axpy2(n, _abi.ptr(num), _abi.ptr(den), _abi.ptr(ty), _abi.ptr(w), _abi.ptr(tx),
                                     _abi.ptr(rho), _abi.stream_ptr()))
        assert np.allclose(w.cpu().numpy(), x + 0.75 * y, rtol=1e-15) and np.allclose(rho.cpu().numpy(), y - 0.75 * x, rtol=1e-15)
        v = tx.clone()
        _abi.check(lib.spuq_sg_xpay(n, _abi.ptr(num), _abi.ptr(den), _abi.ptr(ty), _abi.ptr(v), _abi.stream_ptr()))
        assert np.allclose(v.cpu().numpy(), y + 0.75 * x, rtol=1e-15)


# ---- pcg ------------------------------------------------------------------------------------------
def check_pcg_reference_test():
    """application/egsz/tests/test_pcg.py:12-44 against the product classes."""
    rand = np.random.mtrand.RandomState(1234).random_sample
    N = 7
    M = rand((N, N))
    A = sp.MatrixOperator(np.dot(M, M.T))
    P = sp.MultiplicationOperator(1, sp.CanonicalBasis(N))
    x = sp.FlatVector(rand((N,)))
    b = A * x
    assert np.allclose(b.coeffs, np.dot(M, M.T) @ x.coeffs, rtol=1e-14)
    x_ap, zeta, it = sp.pcg(A, b, P, 0 * x, eps=0.00001)
    np.testing.assert_array_almost_equal(x.coeffs, x_ap.coeffs)
    assert it <= N + 3
    # oracle on the same numbers: same iteration count, same zeta trajectory
    oh, ph = [], []
    oA = olin.MatrixOperator(np.dot(M, M.T))
    ox, _, oit = osg.pcg(oA, oA * olin.FlatVector(x.coeffs.copy()), olin.MultiplicationOperator(1, olin.CanonicalBasis(N)),
                         olin.FlatVector(np.zeros(N)), eps=0.00001, history=oh)
    sp.pcg(A, b, P, 0 * x, eps=0.00001, history=ph)
    assert abs(it - oit) <= 1 and np.allclose(ph[:4], oh[:4], rtol=1e-9)
    P = sp.MatrixSolveOperator(np.dot(M, M.T))
    x_ap, zeta, it = sp.pcg(A, b, P, 0 * x, eps=0.00001)
    np.testing.assert_array_almost_equal(x.coeffs, x_ap.coeffs)
    assert it <= 2
    P = sp.MatrixOperator(np.asarray(P.as_matrix()))
    x_ap, zeta, it = sp.pcg(A, b, P, 0 * x, eps=0.00001)
    np.testing.assert_array_almost_equal(x.coeffs, x_ap.coeffs)
    P = sp.DiagonalMatrixOperator(np.diag(A.as_matrix())).inverse()
    x_ap, zeta, it = sp.pcg(A, b, P, 0 * x, eps=0.00001)
    np.testing.assert_array_almost_equal(x.coeffs, x_ap.coeffs)
    # leaf known answers, linalg/tests/test_operator.py:44-49,112-117
    assert sp.MatrixOperator(np.array([[1, 2, 4], [3, 4, 5]], dtype=float)) * sp.FlatVector([2, 3, 4]) == sp.FlatVector([24, 38])
    assert sp.DiagonalMatrixOperator(np.array([1, 2, 3])) * sp.FlatVector([3, 4, 5]) == sp.FlatVector([3, 8, 15])
    with pytest.raises(TypeError):
        sp.pcg(A, b, 3, 0 * x)


def check_pcg_failures():
    N = 4
    basis = sp.CanonicalBasis(N)
    I = sp.MultiplicationOperator(1, basis)
    x = sp.FlatVector(np.ones(N))
    with pytest.raises(Exception, match="Matrix for PCG is not positive definite"):
        sp.pcg(sp.MatrixOperator(-np.eye(N)), x, I, 0 * x)
    with pytest.raises(Exception, match="Matrix for PCG is singular"):
        sp.pcg(sp.MatrixOperator(np.zeros((N, N))), x, I, 0 * x)
    with pytest.raises(Exception, match="Preconditioner for PCG is not positive definite"):
        sp.pcg(sp.MatrixOperator(np.eye(N)), x, sp.MultiplicationOperator(-1, basis), 0 * x)
    with pytest.raises(Exception, match="PCG did not converge"):
        sp.pcg(sp.MatrixOperator(np.diag([1.0, 10.0, 100.0, 1000.0])), x, I, 0 * x, eps=1e-14, maxiter=3)
    # the messages carry the value the reference prints: alpha = <Av, v> for the matrix failures (pcg.py:42-45),
    # zeta for the preconditioner failure (:33-36); same text as the oracle's exceptions
    A2, x2 = np.diag([1.0, -2.0]), np.array([1.0, 1.0])
    with pytest.raises(Exception) as e_o:
        osg.pcg(olin.MatrixOperator(A2), olin.FlatVector(x2), olin.MultiplicationOperator(1, olin.CanonicalBasis(2)),
                olin.FlatVector(0 * x2))
    with pytest.raises(Exception) as e_p:
        sp.pcg(sp.MatrixOperator(A2), sp.FlatVector(x2), sp.MultiplicationOperator(1, sp.CanonicalBasis(2)), sp.FlatVector(0 * x2))
    assert str(e_p.value) == str(e_o.value) == "Matrix for PCG is not positive definite (-1.0)"
    # maxiter beyond the device history (4096 passes) is accepted, as in the reference
    xs, zeta, it = sp.pcg(sp.MatrixOperator(np.diag([1.0, 2.0, 3.0, 4.0])), x, I, 0 * x, eps=1e-10, maxiter=10000)
    assert it <= 6 and np.allclose(xs.coeffs, 1.0 / np.array([1.0, 2.0, 3.0, 4.0]))


def check_pcg_sg(n=8, M=3, p=2, kind="legendre", precond="jacobi", eps=1e-8, poll_every=3):
    """PCG on the SG system against the oracle's PCG on the same numbers: same status, iteration
    count within one, zeta trajectory to 1e-9 at fixed iteration index, same solution."""
    idx = synthetic.complete_order_indices(M, p)
    c = fd_case(n, M, idx, kind)
    basis_o = olin.CanonicalBasis(c.N)
    f_o = c.A * c.w                                     # right-hand side f = A x*, x* ~ U(0,1)
    w0_o = 0 * c.w
    d = c.mats[0].diagonal()
    if precond == "jacobi":
        Pleaf_o = olin.DiagonalMatrixOperator(1.0 / d, basis_o, basis_o)
        Pleaf_p = lambda b: sp.DiagonalMatrixOperator(1.0 / d, b, b)            # noqa: E731
    else:
        Pleaf_o = olin.MultiplicationOperator(1.0, basis_o)
        Pleaf_p = None
    P_o = osg.PreconditioningOperator("mean", lambda b, f: Pleaf_o) if precond == "jacobi" else Pleaf_o
    oh = []
    if precond == "jacobi":
        x_o, zeta_o, it_o = osg.pcg(c.A, f_o, P_o, w0_o, eps=eps, maxiter=200, history=oh)
    else:
        class _Id(olin.Operator):
            def apply(self, v):
                return v.copy()
        x_o, zeta_o, it_o = osg.pcg(c.A, f_o, _Id(), w0_o, eps=eps, maxiter=200, history=oh)

    A = sp.MultiOperator.from_blocks(device_blocks(c), c.rvs)
    f = A.apply(sp.SlabMultiVector.from_array(idx, c.W))
    w0 = 0 * f
    if precond == "jacobi":
        class _JacobiSolver:                                      # explicit-matrix mean preconditioner
            def __init__(self, b):
                self._d = sp.DiagonalMatrixOperator(1.0 / d, b, b)

            def as_matrix(self):
                return self._d.as_matrix()
        P = sp.PreconditioningOperator("mean", lambda b, f_: _JacobiSolver(b))
    else:
        P = sp.MultiplicationOperator(1.0, sp.CanonicalBasis(c.N))
    ph = []
    x, zeta, it = sp.pcg(A, f, P, w0, eps=eps, maxiter=200, poll_every=poll_every, history=ph)
    assert isinstance(x, sp.SlabMultiVector)
    assert abs(it - it_o) <= 1, (it, it_o)
    k = min(len(ph), len(oh)) - 1
    assert np.allclose(ph[:k], oh[:k], rtol=1e-9), (ph[:k], oh[:k])
    assert len(ph) == it and zeta == ph[-1] and zeta <= eps ** 2
    X = x.to_array()
    assert np.max(np.abs(X - x_o.to_slab())) <= 1e-6 * np.max(np.abs(c.W))
    # distance to the exact solution: no worse than the oracle's own (set by eps and the conditioning)
    assert np.max(np.abs(X - c.W)) <= 2.0 * np.max(np.abs(x_o.to_slab() - c.W)) + 1e-9
    # host containers in -> host container out
    xh, _, ith = sp.pcg(A, f.to_multivector(), P, w0.to_multivector(), eps=eps, maxiter=200)
    assert type(xh) is sp.MultiVector and ith == it
    # the mean preconditioner as an operator on its own
    s = P * f if precond == "jacobi" else None
    if s is not None:
        assert relerr(s.to_array(), f.to_array() / d[:, None]) <= 1e-14


# ---- K4: exact mean preconditioner ------------------------------------------------------------------
def check_mean_preconditioner():
    """S = A_0^{-1} R column by column against scipy's sparse direct solve (what the reference's
    ScipySolveOperator.apply calls, linalg/scipy_operator.py:45-49)."""
    import scipy.sparse.linalg as spla
    from spuq_b200.multi_operator import MeanSolveSlabOperator
    rng = np.random.RandomState(21)
    # even nx takes the parity-split transform (half sizes 6, 35, 65, 2, 3: even and odd; nx < 4 keeps the full transform), odd nx the full one
    # ny > 32 and nx > 31 exercise the two-level strip partition (several strips of two heights, x-chunks
    # of two lengths with their separator systems); the rest a single strip / the plain transform
    for nx, ny, L in [(12, 12, 5), (70, 9, 3), (9, 33, 4), (130, 2, 2), (17, 1, 3), (4, 3, 2), (6, 7, 3),
                      (70, 75, 3), (33, 67, 2), (95, 64, 2), (31, 100, 2), (64, 34, 2), (100, 133, 1),
                      # 16 x-chunks at pitch 33 with padded shared-memory rows (the layout of configurations 2 and 4)
                      (500, 40, 1), (512, 30, 2), (498, 29, 1), (66, 40, 1)]:
        T = lambda n, o, dd: sps.diags([o, dd, o], [-1, 0, 1], shape=(n, n))      # noqa: E731
        if ny > 1:
            A0 = sps.kron(sps.identity(ny), T(nx, -1.0, 2.0)) + sps.kron(T(ny, -1.0, 2.0), sps.identity(nx))
        else:
            A0 = T(nx, -1.0, 4.0)
        A0 = sps.csr_matrix(A0)
        N = nx * ny
        R = rng.random_sample((L, N))
        solver = MeanSolveSlabOperator(A0, L)
        Rt = _dev(torch.from_numpy(R.copy()))
        St = torch.empty_like(Rt)
        solver.apply_slab(Rt, St)
        lu = spla.splu(A0.tocsc())
        ref = np.stack([lu.solve(R[k]) for k in range(L)])
        err = np.max(np.linalg.norm(St.cpu().numpy() - ref, axis=1) / np.linalg.norm(ref, axis=1))
        assert err <= 1e-12, (nx, ny, err)
        assert torch.equal(Rt.cpu(), torch.from_numpy(R))
        solver.apply_slab(Rt, Rt)                                   # in place (input and output alias)
        err = np.max(np.linalg.norm(Rt.cpu().numpy() - ref, axis=1) / np.linalg.norm(ref, axis=1))
        assert err <= 1e-12, (nx, ny, err, "in place")
        if (nx, ny) in ((70, 75), (64, 34), (100, 133)):
            # strips + the separator system in ONE kernel: forward and final pass (main strips + remainder strip
            # each) and k_schur_fused
            assert _abi.lib().spuq_sg_precond_last_launches(solver.handle) <= 5, (nx, ny)
    with pytest.raises(_abi.SpuqError, match="constant-coefficient"):
        MeanSolveSlabOperator(sps.csr_matrix(sps.diags([np.arange(1.0, 10.0)], [0])), 1)


def check_sparse_lu_preconditioner():
    """Exact mean solve for blocks the strip solver does not recognise -- variable coefficients, a
    renumbered grid, an unsymmetric matrix -- against scipy's sparse direct solve
    (ScipySolveOperator.apply, linalg/scipy_operator.py:45-49)."""
    import scipy.sparse.linalg as spla
    from oracle import synth
    from spuq_b200.multi_operator import SparseLUSlabOperator, mean_solver_for
    rng = np.random.RandomState(5)
    n = 20
    mats = synth.fd_matrices(n, 2)
    var = sps.csr_matrix(mats[0] + 0.7 * mats[1] + 0.2 * mats[2])            # variable mean field
    perm = synth.tile_permutation(n, 5)
    cases = {"variable": var, "permuted": synth.permuted([mats[0]], perm)[0],
             "unsymmetric": sps.csr_matrix(var + sps.diags([0.3 * rng.random_sample(n * n - 1)], [1]))}
    for name, A0 in cases.items():
        N = A0.shape[0]
        for L in (1, 7, 40):
            R = rng.random_sample((L, N))
            solver = mean_solver_for(A0, L)
            assert isinstance(solver, SparseLUSlabOperator), name
            Rt = _dev(torch.from_numpy(R.copy()))
            St = torch.empty_like(Rt)
            solver.apply_slab(Rt, St)
            lu = spla.splu(A0.tocsc())
            ref = np.stack([lu.solve(R[k]) for k in range(L)])
            err = np.max(np.linalg.norm(St.cpu().numpy() - ref, axis=1) / np.linalg.norm(ref, axis=1))
            assert err <= 1e-12, (name, L, err)
            solver.apply_slab(Rt, Rt)                               # in place
            err = np.max(np.linalg.norm(Rt.cpu().numpy() - ref, axis=1) / np.linalg.norm(ref, axis=1))
            assert err <= 1e-12, (name, L, err, "in place")
    # a constant-coefficient grid block still takes the strip solver
    from spuq_b200.multi_operator import MeanSolveSlabOperator
    assert isinstance(mean_solver_for(mats[0], 3), MeanSolveSlabOperator)


def check_pcg_variable_mean(n=12, M=3, p=2, eps=1e-10):
    """PCG with the mean-based preconditioner when the mean field is NOT constant (the case round 1
    refused): zeta trajectory and solution against the oracle's PCG with a sparse direct solve; and a
    MULTIPLY-type operator returned by assemble_solver is applied as a matrix, as the reference's
    `A0 * w[mu]` does (multi_operator2.py:159), not inverted."""
    idx = synthetic.complete_order_indices(M, p)
    c = fd_case(n, M, idx, "legendre")
    # variable mean: A_0 <- A_0 + 0.6 A_1 (still SPD), the other blocks unchanged
    mats = list(c.mats)
    mats[0] = sps.csr_matrix(mats[0] + 0.6 * mats[1])
    A_o, _ = oracle_operator(mats, c.orvs)
    basis_o = olin.CanonicalBasis(c.N)
    solve_o = olin.ScipySolveOperator(mats[0], basis_o, basis_o, factor=True)
    P_o = osg.PreconditioningOperator("mean", lambda b, f: solve_o)
    f_o = A_o * c.w
    oh = []
    x_o, zeta_o, it_o = osg.pcg(A_o, f_o, P_o, 0 * c.w, eps=eps, maxiter=100, history=oh)

    A = sp.MultiOperator.from_blocks(DeviceBlocks.from_scipy(mats), c.rvs)
    f = A.apply(sp.SlabMultiVector.from_array(idx, c.W))
    assert relerr(f.to_array(), f_o.to_slab()) <= TOL
    P = sp.PreconditioningOperator("mean", lambda b, f_: sp.ScipySolveOperator(mats[0], b, b))
    ph = []
    x, zeta, it = sp.pcg(A, f, P, 0 * f, eps=eps, maxiter=100, history=ph)
    assert abs(it - it_o) <= 1, (it, it_o)
    k = min(len(ph), len(oh)) - 1
    assert np.allclose(ph[:k], oh[:k], rtol=1e-9), (ph, oh)
    assert np.max(np.abs(x.to_array() - x_o.to_slab())) <= 1e-8
    # multiply-type leaf as "preconditioner": P * f == A_0 f column by column
    Pm = sp.PreconditioningOperator("mean", lambda b, f_: sp.ScipyOperator(mats[0], b, b))
    s = (Pm * f).to_array()
    ref = mats[0] @ f.to_array()
    assert np.max(np.abs(s - ref)) <= 1e-12 * np.max(np.abs(ref))
    Pd = sp.PreconditioningOperator("mean", lambda b, f_: sp.DiagonalMatrixOperator(mats[0].diagonal(), b, b).inverse())
    s = (Pd * f).to_array()
    ref = f.to_array() / mats[0].diagonal()[:, None]
    assert np.max(np.abs(s - ref)) <= 1e-12 * np.max(np.abs(ref))


def check_pcg_sg_mean(n=12, M=3, p=2, eps=1e-10):
    """Config-4 style solve: PCG with the exact mean-based preconditioner against the oracle's PCG
    with a sparse direct solve per column."""
    idx = synthetic.complete_order_indices(M, p)
    c = fd_case(n, M, idx, "legendre")
    basis_o = olin.CanonicalBasis(c.N)
    solve_o = olin.ScipySolveOperator(c.mats[0], basis_o, basis_o, factor=True)
    P_o = osg.PreconditioningOperator("mean", lambda b, f: solve_o)
    f_o = c.A * c.w
    oh = []
    x_o, zeta_o, it_o = osg.pcg(c.A, f_o, P_o, 0 * c.w, eps=eps, maxiter=100, history=oh)

    A = sp.MultiOperator.from_blocks(device_blocks(c), c.rvs)
    f = A.apply(sp.SlabMultiVector.from_array(idx, c.W))
    P = sp.PreconditioningOperator("mean", lambda b, f_: sp.ScipySolveOperator(c.mats[0], b, b))
    ph = []
    x, zeta, it = sp.pcg(A, f, P, 0 * f, eps=eps, maxiter=100, history=ph)
    assert abs(it - it_o) <= 1, (it, it_o)
    k = min(len(ph), len(oh)) - 1
    assert np.allclose(ph[:k], oh[:k], rtol=1e-9), (ph, oh)
    assert np.max(np.abs(x.to_array() - x_o.to_slab())) <= 1e-8
    assert np.max(np.abs(x.to_array() - c.W)) <= 1e-7
    assert it <= 30                                       # mean-based preconditioning: few iterations
    # a pass: apply, <z,v>, the residual update, strips forward, separator system, strips final (+ <rho,s>), the
    # solution / direction update -- seven launches (five when the grid is a single strip); the solve's count
    # includes ~9 for the start-up
    launches = A.plan_for(f).last_launches
    assert launches <= 10 + 7 * it, (launches, it)
    # the preconditioner as a stand-alone operator: P * f, host containers too
    s = P * f
    lu_ref = (P_o * f_o).to_slab()
    assert relerr(s.to_array(), lu_ref) <= 1e-12
    sh = P * f.to_multivector()
    assert type(sh) is sp.MultiVector
