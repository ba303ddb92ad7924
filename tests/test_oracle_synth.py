"""The oracle-side input generators (oracle/synth.py) produce exactly the inputs the product builds
(spuq_b200/synthetic.py), and the sampled form of the oracle's apply (oracle/sampled.py) equals the full
oracle apply on the sampled columns."""
import numpy as np
import pytest

from oracle import synth
from oracle.sampled import SampledApply
from spuq_b200 import synthetic
from tests.cases import fd_case


@pytest.mark.parametrize("spec,builder", [
    (("td", 5, 2), lambda: synthetic.complete_order_indices(5, 2)),
    (("td", 7, 3), lambda: synthetic.complete_order_indices(7, 3)),
    (("aniso", 50, 11.3), lambda: synthetic.anisotropic_indices(50, 11.3)),
    (("aniso", 8, 9.0), lambda: synthetic.anisotropic_indices(8, 9.0)),
    (("prefix", 9, 3, 100), lambda: synthetic.truncated_complete_indices(9, 3, 100)),
])
def test_index_sets_identical(spec, builder):
    assert np.array_equal(synth.index_set(spec), builder())


def test_blocks_identical():
    for n, M in ((8, 3), (16, 6)):
        mats = synth.fd_matrices(n, M)
        rp, ci, v = synthetic.fd_blocks(n, M)
        ref = synthetic.blocks_to_scipy(rp, ci, v, n * n)
        assert len(mats) == len(ref) == M + 1
        for a, b in zip(mats, ref):
            assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            assert np.array_equal(a.data, b.data)


def test_tile_permutation_is_a_permutation_and_breaks_the_stencil():
    n = 24
    perm = synth.tile_permutation(n, 8)
    assert sorted(perm.tolist()) == list(range(n * n))
    m = synth.permuted(synth.fd_matrices(n, 1), perm)[0]
    offs = np.unique(m.tocoo().col - m.tocoo().row)
    assert offs.size > 9          # not a 3x3-window stencil any more


def test_permuted_blocks_identical():
    n, M = 12, 3
    perm = synth.tile_permutation(n, 5)
    assert np.array_equal(perm, synthetic.tile_permutation(n, 5))
    ref = synth.permuted(synth.fd_matrices(n, M), perm)
    rp, ci, v = synthetic.permute_blocks(*synthetic.fd_blocks(n, M), perm)
    got = synthetic.blocks_to_scipy(rp, ci, v, n * n)
    for a, b in zip(got, ref):
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
        assert np.array_equal(a.data, b.data)


@pytest.mark.parametrize("kind,mu", [("legendre", 0.0), ("hermite", 0.5)])
def test_sampled_apply_equals_full_apply(kind, mu):
    idx = synthetic.complete_order_indices(4, 3)
    c = fd_case(8, 4, idx, kind, mu)
    pick = [0, 3, 11, idx.shape[0] - 1]
    sa = SampledApply(c.mats, idx, kind, mu, pick, lambda k: c.W[:, k])
    out, _ = sa.run()
    for k in pick:
        assert np.array_equal(out[k], c.Y[:, k])
