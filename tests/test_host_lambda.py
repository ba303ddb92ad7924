"""Product host code (no GPU): Multiindex / MultiindexSet API, the K5 integer set-up in C++
(sort + neighbour tables, bit-exact against the oracle's Python lookups) and the beta tables."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import multiindex as omi
from oracle import polys as opolys
from oracle import sg as osg
from spuq_b200 import lambda_tables, polyquad, synthetic
from spuq_b200.multiindex import Multiindex, MultiindexSet
from spuq_b200.random_variable import NormalRV, UniformRV

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True)
def _lib(real_lib_path):
    # the K5 entry points are host code inside libspuq_sg.so: loadable without a device
    from spuq_b200 import _abi
    _abi.reset()
    yield


def test_set_goldens_product():
    with open(os.path.join(GOLD, "multiindex_sets.json")) as f:
        g = json.load(f)
    for key, rec in g.items():
        if "," not in key:
            continue
        m, p = map(int, key.split(","))
        arr = MultiindexSet.createCompleteOrderSet(m, p).arr
        assert list(arr.shape) == rec["shape"]
        assert hashlib.sha256(np.ascontiguousarray(arr, dtype=np.int32).tobytes()).hexdigest() == rec["sha256"]
    gen = MultiindexSet.createCompleteOrderSet(2)
    assert [next(gen).tolist() for _ in range(12)] == g["generator_2_first12"]
    assert MultiindexSet.createCompleteOrderSet(3, 2, reversed=True).arr.tolist() == g["reversed_3_2"]
    s = MultiindexSet.createCompleteOrderSet(7, 5)
    assert (s.m, s.p, s.count) == (7, 5, 792)
    assert repr(MultiindexSet.createCompleteOrderSet(2, 3))[:25] == g["repr_2_3_prefix"]


def test_multiindex_api():
    # the reference's test_multiindex.py, against the product class
    with pytest.raises(TypeError):
        Multiindex([0, 1.3, 3, 0])
    a = Multiindex(np.array([0, 1, 3, 0]))
    assert np.array_equal(a.as_array, [0, 1, 3]) and len(a) == 3 and a.order == 4
    b = Multiindex([2, 1, 3])
    assert (b[0], b[2], b[-1], b[3]) == (2, 3, 0, 0)
    arr = np.array([0, 1, 3, 0]); mi = Multiindex(arr); arr[2] = 4
    assert mi == Multiindex([0, 1, 3, 0])
    assert a == Multiindex([0, 1, 3, 0, 0, 0]) and hash(a) == hash(Multiindex([0, 1, 3, 0, 0, 0]))
    assert a != Multiindex([0, 1, 3, 0, 4, 0, 0])
    assert a.inc(4, 7) == Multiindex([0, 1, 3, 0, 7, 0]) and Multiindex([0, 1, 3, 0, 7]).dec(4, 7) == a
    assert a.dec(0, 1) is None and a.dec(8, 1) is None and Multiindex().inc(0) == Multiindex([1])
    assert list(Multiindex([0, 2, 0, 1]).supp) == [1, 3]
    a1, a2, a3, a4, a5 = (Multiindex(x) for x in ([0, 1, 3, 0], [0, 1, 3], [1, 3, 0], [0, 1, 4], [1, 4]))
    for lo, hi in [(a1, a2), (a2, a1), (a2, a3), (a2, a4), (a3, a4), (a3, a5), (a4, a5)]:
        assert lo <= hi and not lo > hi
    assert str(Multiindex([0, 1, 2, 3])) == "<Multiindex inds=[0 1 2 3]>"
    s = Multiindex.createCompleteOrderSet(3, 2)
    assert isinstance(s[4], Multiindex) and len(s) == 10


def _random_downward_closed(M, L, seed):
    """Random downward-closed set: grow from 0 by adding admissible neighbours."""
    rng = np.random.RandomState(seed)
    have = {tuple([0] * M)}
    frontier = [tuple([0] * M)]
    while len(have) < L:
        base = frontier[rng.randint(len(frontier))]
        m = rng.randint(M)
        cand = list(base); cand[m] += 1
        ok = all(tuple(cand[:j] + [cand[j] - 1] + cand[j + 1:]) in have for j in range(M) if cand[j] > 0)
        if ok and tuple(cand) not in have:
            have.add(tuple(cand)); frontier.append(tuple(cand))
    arr = np.array(sorted(have), dtype=np.int32)
    rng.shuffle(arr)
    return arr


@pytest.mark.parametrize("arr_fn", [
    lambda: MultiindexSet.createCompleteOrderSet(3, 4).arr,
    lambda: MultiindexSet.createCompleteOrderSet(5, 2).arr,
    lambda: MultiindexSet.createCompleteOrderSet(1, 6).arr,
    lambda: _random_downward_closed(6, 80, 3),
    lambda: _random_downward_closed(4, 50, 11)[5:],          # not downward closed: missing neighbours
    lambda: np.array([[0, 0], [1, 0], [0, 1], [0, 2]], dtype=np.int32),   # the reference's test set
])
def test_sort_and_neighbours_bit_exact(arr_fn):
    arr = np.asarray(arr_fn(), dtype=np.int32)
    perm = lambda_tables.sort_permutation(arr)
    # oracle: Python sorted() on Multiindex objects, `in` lookups
    objs = [omi.Multiindex(r.astype(np.int64)) for r in arr]
    order = sorted(range(len(objs)), key=lambda k: objs[k])
    assert [list(objs[k].as_array) for k in order] == [list(omi.Multiindex(arr[k].astype(np.int64)).as_array) for k in perm]
    assert np.array_equal(perm, np.array(order, dtype=np.int32))
    sorted_arr = arr[perm]
    Lam = sorted(objs)
    M = arr.shape[1]
    up_o, dn_o = osg.neighbour_tables(Lam, M)
    up, dn = lambda_tables.neighbour_tables(sorted_arr)
    assert np.array_equal(up, up_o) and np.array_equal(dn, dn_o)
    # product Multiindex sorts the same way
    assert [list(m.as_array) for m in sorted(Multiindex(r) for r in arr)] == [list(m.as_array) for m in Lam]


def test_sort_rejects_bad_input():
    from spuq_b200 import _abi
    with pytest.raises(_abi.SpuqError, match="duplicate"):
        lambda_tables.sort_permutation(np.array([[1, 0], [1, 0]], dtype=np.int32))
    with pytest.raises(_abi.SpuqError, match="negative"):
        lambda_tables.sort_permutation(np.array([[1, -1]], dtype=np.int32))


def test_beta_goldens_product():
    with open(os.path.join(GOLD, "beta.json")) as f:
        g = json.load(f)
    fams = {
        "legendre[-1,1]": polyquad.LegendrePolynomials(),
        "legendre[0,1]": polyquad.LegendrePolynomials(0.0, 1.0),
        "hermite(0,1)": polyquad.StochasticHermitePolynomials(),
        "hermite(0.5,1)": polyquad.StochasticHermitePolynomials(0.5, 1.0),
        "hermite(0,2)": polyquad.StochasticHermitePolynomials(0.0, 2.0),
        "jacobi(0.5,1;-2,3)": polyquad.JacobiPolynomials(0.5, 1.0, -2.0, 3.0),
    }
    for name, fam in fams.items():
        for n in range(21):
            b = fam.get_beta(n)
            assert [float(b[0]).hex(), float(b[1]).hex(), float(b[2]).hex()] == g[name][n], (name, n)
    assert UniformRV().orth_polys.normalised and NormalRV(0.5).orth_polys.get_beta(2)[0] == -0.5


def test_beta_tables_match_oracle_per_index():
    idx = synthetic.anisotropic_indices(6, 7.0)
    rvs = [UniformRV(-1, 1), NormalRV(0.5, 1), UniformRV(0, 1), NormalRV(0, 2), UniformRV(-1, 1), NormalRV(0, 1)]
    orvs = [opolys.UniformRV(-1, 1), opolys.NormalRV(0.5, 1), opolys.UniformRV(0, 1), opolys.NormalRV(0, 2),
            opolys.UniformRV(-1, 1), opolys.NormalRV(0, 1)]
    t = lambda_tables.LambdaTables(idx, rvs)
    assert t.maxm == 6 and t.L == idx.shape[0]
    for m in range(6):
        for k in range(t.L):
            b = orvs[m].orth_polys.get_beta(int(t.index_array[k, m]))
            assert (t.beta_0[m, k], t.beta_up[m, k], t.beta_dn[m, k]) == (float(b[0]), float(b[1]), float(b[-1]))


def test_synthetic_sets():
    assert synthetic.complete_order_indices(5, 2).shape == (21, 5)
    assert synthetic.complete_order_indices(20, 3).shape == (1771, 20)
    c3 = synthetic.anisotropic_indices(50, 11.3)
    assert c3.shape[0] == 988
    up, dn = lambda_tables.neighbour_tables(c3)
    assert int((up >= 0).sum()) == 2244 == int((dn >= 0).sum())
    assert int(c3.max()) == 8
    c4 = synthetic.truncated_complete_indices(30, 3, 5000)
    assert c4.shape == (5000, 30)
    up, dn = lambda_tables.neighbour_tables(c4)
    # prefix of a graded order is downward closed: every mu with mu_m > 0 has its down-neighbour
    assert np.array_equal(dn >= 0, c4.T > 0)
    assert synthetic.kl_frequencies(5) == [(0, 1), (1, 0), (0, 2), (1, 1), (2, 0)]
    amps = synthetic.kl_amplitudes(3)
    assert abs(amps[0] - 0.9 / 0.6449340668482264 / 4) < 1e-12
