"""Parity tests proper: the CUDA path on a B200, called through the C ABI, against the oracle.

Small cases: the whole parity suite (same checks the emulated build passes on the CPU).
Full-size cases (BASELINE.json configs 2 and 3): the oracle cannot run L*(M+1) SpMVs of 10^6 rows in
seconds, so they are checked (i) against the oracle on a SAMPLE of output columns -- y[mu] only
needs w[mu] and its 2M neighbours, which scipy does in milliseconds -- and (ii) through
size-independent properties: linearity, symmetry of the SG operator, agreement of every kernel
path, and the M-shard sum."""
import numpy as np
import pytest
import torch

from tests import parity_suite as ps

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _backend(gpu):
    yield


def test_library_is_the_cuda_build(gpu):
    assert gpu.library_path().endswith("spuq_b200/libspuq_sg.so") and not gpu.is_emulated()
    assert torch.cuda.get_device_capability(0)[0] == 10


def test_apply_paths_legendre():
    ps.check_apply_paths(12, 4, 2, "legendre")


def test_apply_paths_hermite_shifted():
    ps.check_apply_paths(10, 3, 3, "hermite", 0.5)


def test_apply_paths_wide_grid():
    ps.check_apply_paths(70, 3, 2, "hermite", 0.5)        # x-tiles with a ragged tail, several y-tiles


@pytest.mark.parametrize("n,M,p,kind,mu", [(70, 3, 2, "hermite", 0.5), (132, 2, 3, "legendre", 0.0),
                                           (12, 4, 2, "legendre", 0.0)])
def test_apply_tma_variants(n, M, p, kind, mu):
    """The TMA-pipelined kernel (variant 3) in all instantiated shapes, incl. grids smaller than one
    TMA box (zero-filled out-of-bounds reads) and ragged tile tails."""
    from spuq_b200 import _abi, synthetic
    from tests.cases import fd_case, relerr
    idx = synthetic.complete_order_indices(M, p)
    c = fd_case(n, M, idx, kind, mu)
    for tune in [(3, 8, 2, 0), (3, 4, 2, 0), (3, 12, 2, 0), (3, 16, 2, 0), (3, 4, 4, 0), (3, 8, 1, 0), (3, 16, 1, 0),
                 (3, 8, 2, 1)]:
        Y, plan = ps.product_apply(c, _abi.PATH_STENCIL, tune)
        assert "tma" in plan.kernel_name
        assert relerr(Y, c.Y) <= ps.TOL, (tune, plan.kernel_name, relerr(Y, c.Y))


@pytest.mark.parametrize("n,M,p,kind,mu", [(70, 3, 2, "hermite", 0.5), (132, 2, 3, "legendre", 0.0),
                                           (12, 4, 2, "legendre", 0.0), (66, 5, 3, "legendre", 0.0)])
def test_apply_tmem_variants(n, M, p, kind, mu):
    """The TMEM-accumulator kernel (variant 4) for every register-set count, incl. grids smaller than
    one TMA box, ragged tile tails and (66, 5, 3): L = 56 columns on several tiles."""
    from spuq_b200 import _abi, synthetic
    from tests.cases import fd_case, relerr
    idx = synthetic.complete_order_indices(M, p)
    c = fd_case(n, M, idx, kind, mu)
    for tune in [(4, 1, 0, 0), (4, 2, 0, 0), (4, 1, 1, 0), (4, 2, 1, 0), (4, 3, 1, 0), (4, 2, 3, 0), (4, 1, 3, 0)]:
        Y, plan = ps.product_apply(c, _abi.PATH_STENCIL, tune)
        assert "tmem" in plan.kernel_name
        assert relerr(Y, c.Y) <= ps.TOL, (tune, plan.kernel_name, relerr(Y, c.Y))


def test_tmem_window_release_stress():
    """The W-ring windows of k_apply_tmem are released by lane 0's mbarrier arrive after the warp's loads were
    ISSUED (no __syncwarp, kernels_tmem.cuh `window_reads_issued`).  If the producer's TMA fill could overtake
    outstanding loads, the result would differ from run to run: 200 applications of a problem with many
    windows per tile (K = 2 and K = 1 register sets, ring lengths 2 and 3) must all be bit-identical, and equal
    the one-column-at-a-time cross-check kernel to the parity tolerance."""
    import torch
    import spuq_b200 as sp
    from spuq_b200 import _abi, synthetic
    from tests.cases import fd_case, relerr
    idx = synthetic.complete_order_indices(8, 3)                    # 165 columns
    c = fd_case(192, 8, idx, "hermite", 0.5, with_oracle=False)
    Yref, plan = ps.product_apply(c, _abi.PATH_ELL, (1, 0, 0, 0))
    A = sp.MultiOperator.from_blocks(ps.device_blocks(c), c.rvs, path=_abi.PATH_STENCIL)
    w = sp.SlabMultiVector.from_array(c.idx, c.W)
    plan = A.plan_for(w)
    Y0 = torch.empty_like(w.slab)
    Y = torch.empty_like(w.slab)
    for tune in [(4, 2, 2, 0), (4, 1, 2, 0), (4, 2, 3, 0)]:
        plan.tune(*tune)
        plan.apply(w.slab, Y0)
        assert "tmem" in plan.kernel_name
        assert relerr(Y0.cpu().numpy().T, Yref) <= ps.TOL
        for _ in range(200):
            Y.fill_(float("nan"))
            plan.apply(w.slab, Y)
            assert torch.equal(Y, Y0), tune


def test_apply_tmem_nine_point():
    from spuq_b200 import _abi, synthetic
    from tests.cases import relerr
    c = ps.nine_point_case(72, 40, 3, synthetic.complete_order_indices(3, 2))
    for tune in [(4, 1, 0, 0), (4, 1, 1, 0), (4, 2, 1, 0), (4, 1, 3, 0)]:
        Y, plan = ps.product_apply(c, _abi.PATH_STENCIL, tune)
        assert "tmem" in plan.kernel_name and "9pt" in plan.kernel_name
        assert relerr(Y, c.Y) <= ps.TOL, (tune, relerr(Y, c.Y))


def test_apply_multiblock():
    ps.check_apply_multiblock()


def test_apply_tma_nine_point():
    from spuq_b200 import _abi, synthetic
    from tests.cases import relerr
    c = ps.nine_point_case(72, 40, 3, synthetic.complete_order_indices(3, 2))
    for tune in [(3, 8, 2, 0), (3, 4, 2, 0), (3, 8, 1, 0)]:
        Y, plan = ps.product_apply(c, _abi.PATH_STENCIL, tune)
        assert "tma" in plan.kernel_name and "9pt" in plan.kernel_name
        assert relerr(Y, c.Y) <= ps.TOL, (tune, relerr(Y, c.Y))


def test_apply_c1_dense_callbacks():
    ps.check_apply_c1_dense_callbacks()


def test_apply_reference_known_answer():
    ps.check_apply_reference_known_answer()


def test_apply_unstructured():
    ps.check_apply_unstructured()


def test_apply_nine_point_and_odd_shapes():
    ps.check_apply_nine_point_and_odd_shapes()


def test_apply_edge_cases():
    ps.check_apply_edge_cases()


def test_apply_m_shards():
    ps.check_apply_m_shards()


def test_slab_multivector_api():
    ps.check_slab_multivector_api()


def test_blas():
    ps.check_blas()


def test_pcg_reference_test():
    ps.check_pcg_reference_test()


def test_pcg_failures():
    ps.check_pcg_failures()


@pytest.mark.parametrize("precond", ["jacobi", "identity"])
def test_pcg_sg(precond):
    ps.check_pcg_sg(precond=precond)


def test_pcg_sg_bigger():
    ps.check_pcg_sg(n=24, M=4, p=3, kind="hermite", precond="jacobi", eps=1e-4, poll_every=8)


def test_mean_preconditioner():
    ps.check_mean_preconditioner()


def test_pcg_sg_mean():
    ps.check_pcg_sg_mean()
    ps.check_pcg_sg_mean(n=40, M=2, p=2)          # several strips: fused separator system and <rho,s>


def test_sparse_lu_preconditioner():
    ps.check_sparse_lu_preconditioner()


def test_pcg_variable_mean():
    ps.check_pcg_variable_mean()


def test_row_sharded_pcg_single_rank():
    ps.check_row_sharded_pcg_single_rank()
    ps.check_row_sharded_pcg_single_rank(n=32, M=4, p=2)


def test_pcg_solve_and_prepare_rhs():
    ps.check_pcg_solve()
    ps.check_pcg_solve(n=16, M=4, p=2, kind="legendre", mu=0.0)


# ---- full-size configurations ---------------------------------------------------------------------
def _oracle_columns(n, M, idx, kind, mu, W, cols):
    """Oracle y[mu] for a few output columns: `oracle.sg.MultiOperator.apply(w, only=...)`, the unchanged
    loop body of multi_operator2.py:52-83, on inputs the ORACLE side builds for itself -- blocks from
    oracle/synth.py, index set, `in Lambda` look-ups and beta values from the oracle's own classes.
    Nothing of the product's plan (neighbour / beta tables) enters the expected values."""
    from oracle import synth
    from oracle.sampled import SampledApply
    oidx = synth.index_set(idx) if isinstance(idx, tuple) else idx
    mats = synth.fd_matrices(n, M)
    sa = SampledApply(mats, oidx, kind, mu, cols, lambda k: W[k].cpu().numpy())
    out, _ = sa.run()
    return out


def _full_size(n, M, idx, kind, mu, sample_cols, variants, lam_spec=None):
    import spuq_b200 as sp
    from spuq_b200 import _abi, synthetic
    from spuq_b200.multi_operator import DeviceBlocks
    dev = _abi.device()
    N, L = n * n, idx.shape[0]
    rowptr, colidx, vals = synthetic.fd_blocks(n, M, device=dev)
    blocks = DeviceBlocks(rowptr, colidx, vals, (N, N))
    rvs = synthetic.make_rvs(kind, M, mu)
    A = sp.MultiOperator.from_blocks(blocks, rvs)
    g = torch.Generator(device=dev); g.manual_seed(1234)
    W = torch.rand((L, N), dtype=torch.float64, device=dev, generator=g)
    w = sp.SlabMultiVector(idx, slab=W, _sorted=True)
    plan = A.plan_for(w)
    y = A.apply(w)
    # (i) sampled columns against the oracle formula
    if lam_spec is not None:
        from oracle import synth
        assert np.array_equal(synth.index_set(lam_spec), idx)      # the oracle's own construction of Lambda
    ref = _oracle_columns(n, M, idx, kind, mu, W, sample_cols)
    ynorm = float(torch.linalg.norm(y.slab))
    for mu_, r in ref.items():
        err = np.linalg.norm(y.slab[mu_].cpu().numpy() - r) / np.linalg.norm(r)
        assert err <= 1e-12, (mu_, err)
    # (ii) every kernel variant produces the same slab (to reassociation)
    for tune in variants:
        plan.tune(*tune)
        y2 = A.apply(w)
        d = float(torch.linalg.norm(y2.slab - y.slab)) / ynorm
        assert d <= 1e-13, (tune, plan.kernel_name, d)
    plan.tune(0, 0, 0, 0)
    # (iii) linearity and symmetry
    X = torch.rand((L, N), dtype=torch.float64, device=dev, generator=g)
    x = w.like(X)
    ax = A.apply(x)
    comb = A.apply(w.like(2.5 * W - 0.75 * X))
    lin = float(torch.linalg.norm(comb.slab - (2.5 * y.slab - 0.75 * ax.slab))) / ynorm
    assert lin <= 1e-13, lin
    s1, s2 = sp.inner(y, x), sp.inner(w, ax)
    assert abs(s1 - s2) <= 1e-12 * abs(s1), (s1, s2)
    del comb, ax, x, X
    # (iv) M-shard sum
    half = M // 2
    A1 = sp.MultiOperator.from_blocks(blocks, rvs, m_range=(0, half), include_mean=True)
    A2 = sp.MultiOperator.from_blocks(blocks, rvs, m_range=(half, M), include_mean=False)
    ysum = A1.apply(w)
    ysum += A2.apply(w)
    d = float(torch.linalg.norm(ysum.slab - y.slab)) / ynorm
    assert d <= 1e-13, d


def test_full_size_config2():
    """512 x 512 grid, M = 20, total degree 3 (1771 indices), Legendre: 3.7 GB per slab."""
    from spuq_b200 import synthetic
    idx = synthetic.complete_order_indices(20, 3)
    _full_size(512, 20, idx, "legendre", 0.0, [0, 1, 7, 230, 231, 1000, 1770],
               [(1, 0, 0, 0), (2, 4, 2, 0), (3, 8, 2, 0), (4, 3, 1, 0)], lam_spec=("td", 20, 3))


def test_full_size_config3_hermite_shifted():
    """1000 x 1000 grid, M = 50, anisotropic Lambda (988 indices), Hermite chaos with the
    reference's default NormalRV(mu=0.5) (beta0 != 0): 8 GB per slab."""
    from spuq_b200 import synthetic
    idx = synthetic.anisotropic_indices(50, 11.3)
    _full_size(1000, 50, idx, "hermite", 0.5, [0, 3, 50, 500, 987], [(2, 4, 2, 0), (3, 8, 2, 0), (4, 3, 1, 0), (4, 2, 3, 0)],
               lam_spec=("aniso", 50, 11.3))


def test_full_size_config3_benchmarked():
    """The workload bench.py times: config 3 with NormalRV(mu=0) (beta0 = 0), default kernel only."""
    from spuq_b200 import synthetic
    idx = synthetic.anisotropic_indices(50, 11.3)
    _full_size(1000, 50, idx, "hermite", 0.0, [0, 1, 17, 333, 612, 987], [], lam_spec=("aniso", 50, 11.3))


def test_full_size_config4_apply_and_pcg():
    """500 x 500 grid, M = 30, first 5000 indices of the degree-3 set, Legendre (10 GB per slab): the
    apply properties, then the solve of BASELINE config 4 -- mean-based preconditioner, eps = 1e-10 --
    checked through its manufactured solution and the residual of the returned iterate."""
    import spuq_b200 as sp
    from spuq_b200 import _abi, synthetic
    from spuq_b200.multi_operator import DeviceBlocks
    n, M = 500, 30
    idx = synthetic.truncated_complete_indices(30, 3, 5000)
    _full_size(n, M, idx, "legendre", 0.0, [0, 31, 2500, 4999], [(4, 3, 1, 0), (4, 2, 3, 0)], lam_spec=("prefix", 30, 3, 5000))
    torch.cuda.empty_cache()
    dev = _abi.device()
    N, L = n * n, idx.shape[0]
    rowptr, colidx, vals = synthetic.fd_blocks(n, M, device=dev)
    blocks = DeviceBlocks(rowptr, colidx, vals, (N, N))
    A = sp.MultiOperator.from_blocks(blocks, synthetic.make_rvs("legendre", M, 0.0))
    g = torch.Generator(device=dev); g.manual_seed(1234)
    X = torch.rand((L, N), dtype=torch.float64, device=dev, generator=g)
    f = A.apply(sp.SlabMultiVector(idx, slab=X, _sorted=True))
    A0 = synthetic.blocks_to_scipy(rowptr, colidx, vals[:1], N)[0]
    P = sp.PreconditioningOperator("mean", lambda b, f_: sp.ScipySolveOperator(A0, b, b))
    x, zeta, it = sp.pcg(A, f, P, 0 * f, eps=1e-10, maxiter=100)
    assert it <= 25 and zeta <= 1e-20
    assert float((x.slab - X).abs().max()) <= 1e-10
    r = A.apply(x)
    r -= f
    assert float(torch.linalg.norm(r.slab)) <= 1e-9 * float(torch.linalg.norm(f.slab))


def test_general_sparse_path_100k_rows():
    """The SG operator on a renumbered grid (N = 102400 rows, 21 blocks, 287 indices): the pattern is no
    stencil, so the column-blocked general kernel runs; sampled output columns against the oracle's apply
    on the oracle's own permuted blocks; the one-column-at-a-time kernel as a second opinion."""
    import spuq_b200 as sp
    from oracle import synth
    from oracle.sampled import SampledApply
    from spuq_b200 import _abi, synthetic
    from spuq_b200.multi_operator import DeviceBlocks
    dev = _abi.device()
    n, M = 320, 20
    N = n * n
    idx = synthetic.anisotropic_indices(M, 9.0)
    L = idx.shape[0]
    perm = synthetic.tile_permutation(n, 16)
    rowptr, colidx, vals = synthetic.permute_blocks(*synthetic.fd_blocks(n, M, device=dev), perm)
    A = sp.MultiOperator.from_blocks(DeviceBlocks(rowptr, colidx, vals, (N, N)), synthetic.make_rvs("hermite", M, 0.5))
    g = torch.Generator(device=dev); g.manual_seed(99)
    W = torch.rand((L, N), dtype=torch.float64, device=dev, generator=g)
    w = sp.SlabMultiVector(idx, slab=W, _sorted=True)
    plan = A.plan_for(w)
    y = A.apply(w)
    assert plan.kernel_name == "k_apply_ellblk<5>", plan.kernel_name
    mats = synth.permuted(synth.fd_matrices(n, M), synth.tile_permutation(n, 16))
    cols = [0, 1, 5, L // 2, L - 1]
    ref, _ = SampledApply(mats, synth.index_set(("aniso", M, 9.0)), "hermite", 0.5, cols, lambda k: W[k].cpu().numpy()).run()
    for k, r in ref.items():
        err = np.linalg.norm(y.slab[k].cpu().numpy() - r) / np.linalg.norm(r)
        assert err <= 1e-12, (k, err)
    plan.tune(1, 0, 0, 0)
    y2 = A.apply(w)
    assert plan.kernel_name == "k_apply_ell<8>"
    d = float(torch.linalg.norm(y2.slab - y.slab) / torch.linalg.norm(y.slab))
    assert d <= 1e-13, d


def test_parametric_samples():
    ps.check_parametric_samples()


def test_lambda_growth():
    ps.check_lambda_growth()


def test_partitioned_mean_solver():
    ps.check_partitioned_mean_solver()


def test_neighbour_combination():
    ps.check_neighbour_combination()


def test_upper_tail_bound():
    ps.check_upper_tail_bound()


def test_refine_and_table_growth():
    ps.check_refine_and_table_growth()


def test_row_bands_manual_exchange():
    ps.check_row_bands_manual_exchange()
    ps.check_row_bands_manual_exchange(n=128, M=3, bands=((0, 63), (63, 128)))       # 65 and 67 local rows, TMA-sized grid
