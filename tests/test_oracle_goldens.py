"""The oracle against the reference's own known answers and the committed golden fixtures.

Restated reference tests (paths relative to the reference root, src/spuq/...):
  math_utils/tests/test_multiindex.py, math_utils/tests/test_multiindex_set.py,
  polyquad/tests/test_polynomials.py:208-223, polyquad/tests/test__polynomials.py:68-102,145-162,
  application/egsz/tests/test_multi_vector.py, linalg/tests/test_operator.py:44-49,112-117,
  application/egsz/tests/test_multi_operator.py:57-90, application/egsz/tests/test_pcg.py:12-44.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import linalg as olin
from oracle import polys as op
from oracle import sg as osg
from oracle.multiindex import Multiindex, MultiindexSet

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


# ---- multi-index sets ------------------------------------------------------------------------
def test_set_goldens():
    g = _gold("multiindex_sets.json")
    for key, rec in g.items():
        if "," not in key:
            continue
        m, p = map(int, key.split(","))
        arr = MultiindexSet.createCompleteOrderSet(m, p).arr
        assert list(arr.shape) == rec["shape"]
        a32 = np.ascontiguousarray(arr, dtype=np.int32)
        assert hashlib.sha256(a32.tobytes()).hexdigest() == rec["sha256"]
        if "rows" in rec:
            assert arr.tolist() == rec["rows"]
    gen = MultiindexSet.createCompleteOrderSet(2)
    assert [next(gen).tolist() for _ in range(12)] == g["generator_2_first12"]
    assert MultiindexSet.createCompleteOrderSet(3, 2, reversed=True).arr.tolist() == g["reversed_3_2"]
    assert repr(MultiindexSet.createCompleteOrderSet(2, 3))[:25] == g["repr_2_3_prefix"]


def test_set_counts_reference_tests():
    # test_multiindex_set.py:13-35
    for (m, p, cnt) in [(1, 1, 2), (2, 3, 10), (0, 2, 1), (5, 0, 1), (7, 5, 792)]:
        s = MultiindexSet.createCompleteOrderSet(m, p)
        assert s.count == cnt and len(s) == cnt
    # SURVEY 8c hashes
    a = np.ascontiguousarray(MultiindexSet.createCompleteOrderSet(5, 2).arr, dtype=np.int32)
    assert hashlib.sha256(a.tobytes()).hexdigest() == \
        "bdfa696907ae0452ca87b548fa6f8c471f5ab636ae7e9ad80d665aa4a56c297c"
    a = np.ascontiguousarray(MultiindexSet.createCompleteOrderSet(20, 3).arr, dtype=np.int32)
    assert hashlib.sha256(a.tobytes()).hexdigest() == \
        "23a3835ddcb1f56b8d4cf45b80236ed485ac52b4bf0f03a34e42df69cb486a93"


# ---- Multiindex: test_multiindex.py ------------------------------------------------------------
def test_multiindex_reference_tests():
    Multiindex([0, 1, 3, 0]); Multiindex(np.array([0, 1, 3, 0])); Multiindex(); Multiindex([0, 0])
    with pytest.raises(TypeError):
        Multiindex([0, 1.3, 3, 0])
    with pytest.raises(TypeError):
        Multiindex(np.array([0, 1.3, 3, 0]))
    a = Multiindex(np.array([0, 1, 3, 0]))
    assert np.array_equal(a.as_array, [0, 1, 3])
    b = Multiindex(np.array([2, 1, 3]))
    assert (b[0], b[2], b[-1], b[3]) == (2, 3, 0, 0)
    arr = [0, 1, 3, 0]
    mi = Multiindex(arr); arr[2] = 4
    assert mi == Multiindex([0, 1, 3, 0])
    assert len(Multiindex([0, 1, 3, 0])) == 3 and len(Multiindex([0, 1, 3, 0, 4, 0, 0])) == 5
    alpha, beta, gamma = Multiindex([0, 1, 3, 0]), Multiindex([0, 1, 3, 0, 0, 0]), Multiindex([0, 1, 3, 0, 4, 0, 0])
    assert alpha == beta and alpha != gamma and hash(alpha) == hash(beta)
    assert alpha.order == 4 and Multiindex([0, 1, 3, 0, 4, 0]).order == 8
    assert alpha.inc(4, 7) == Multiindex([0, 1, 3, 0, 7, 0])
    assert Multiindex([0, 1, 3, 0, 7, 0]).dec(4, 7) == alpha
    assert alpha.dec(0, 1) is None and alpha.dec(8, 1) is None
    assert Multiindex().inc(0) == Multiindex([1])
    a1, a2, a3, a4, a5 = (Multiindex(x) for x in ([0, 1, 3, 0], [0, 1, 3], [1, 3, 0], [0, 1, 4], [1, 4]))
    for lo, hi in [(a1, a2), (a2, a1), (a2, a3), (a2, a4), (a3, a4), (a3, a5), (a4, a5)]:
        assert lo <= hi and not lo > hi
    assert Multiindex([0, 1]) in [Multiindex([0, 1])]
    # sorted order differs from the set's row order from m >= 3 on
    assert Multiindex([0, 2, 0]) < Multiindex([1, 0, 1])


# ---- polynomials ---------------------------------------------------------------------------------
def test_beta_goldens_bit_exact():
    g = _gold("beta.json")
    fams = {
        "legendre[-1,1]": op.LegendrePolynomials(),
        "legendre[0,1]": op.LegendrePolynomials(0.0, 1.0),
        "hermite(0,1)": op.StochasticHermitePolynomials(),
        "hermite(0.5,1)": op.StochasticHermitePolynomials(0.5, 1.0),
        "hermite(0,2)": op.StochasticHermitePolynomials(0.0, 2.0),
        "jacobi(0.5,1;-2,3)": op.JacobiPolynomials(0.5, 1.0, -2.0, 3.0),
    }
    for name, fam in fams.items():
        for n in range(21):
            b = fam.get_beta(n)
            got = [float(b[0]).hex(), float(b[1]).hex(), float(b[-1]).hex()]
            assert got == g[name][n], (name, n)


def test_beta_survey_values():
    L = op.LegendrePolynomials()
    assert tuple(map(float, L.get_beta(0))) == (0.0, 0.5773502691896257, 0.0)
    assert tuple(map(float, L.get_beta(1))) == (0.0, 0.5163977794943223, 0.5773502691896258)
    assert tuple(map(float, L.get_beta(3))) == (0.0, 0.5039526306789697, 0.5070925528371101)
    H = op.StochasticHermitePolynomials()
    assert tuple(map(float, H.get_beta(2))) == (0.0, 1.732050807568877, 1.4142135623730951)
    assert tuple(map(float, H.get_beta(3))) == (0.0, 2.0, 1.7320508075688776)
    H5 = op.StochasticHermitePolynomials(0.5, 1.0)
    assert all(float(H5.get_beta(n)[0]) == -0.5 for n in range(6))
    L01 = op.LegendrePolynomials(0.0, 1.0)
    assert float(L01.get_beta(0)[0]) == -0.5 and float(L01.get_beta(0)[1]) == 0.28867513459481287


def test_beta_identity():
    # test_polynomials.py:208-223: x p_n = beta[1] p_{n+1} - beta[0] p_n + beta[-1] p_{n-1}
    x = np.poly1d([1.0, 0.0])
    for fam, ns in [(op.JacobiPolynomials(0.5, 1.0, -2.0, 3.0), (4, 0)), (op.LegendrePolynomials(), (3, 0)),
                    (op.LegendrePolynomials(2.0, 4.0, normalised=False), (3, 0)),
                    (op.StochasticHermitePolynomials(normalised=False), (5, 0))]:
        for n in ns:
            p = fam.eval(n + 1, x, all_degrees=True)
            b = fam.get_beta(n)
            lhs = x * p[n]
            rhs = b[1] * p[n + 1] - b[0] * p[n] + (b[-1] * p[n - 1] if n > 0 else 0)
            assert np.allclose((lhs - rhs).coeffs, 0.0, atol=1e-12)


def test_sqnorm_and_normalised_legendre():
    # test__polynomials.py:68-102, 145-162
    assert [op.sqnorm_legendre(n) for n in range(5)] == [1, 1 / 3, 1 / 5, 1 / 7, 1 / 9]
    assert [op.sqnorm_stoch_hermite(n) for n in range(5)] == [1, 1, 2, 6, 24]
    rc = op.normalise_rc(op.rc_legendre, op.sqnorm_legendre)
    for n in range(5):
        al0, be0 = op.rc_norm_legendre(n)
        _, be1 = op.rc_norm_legendre(n + 1)
        want = (-al0 / be1, 1.0 / be1, be0 / be1)
        assert np.allclose(rc(n), want, rtol=1e-14)
    for n in range(5):
        assert np.isclose(op.sqnorm_from_rc(op.rc_legendre, n), op.sqnorm_legendre(n), rtol=1e-14)


# ---- leaf operators and multi-vector algebra -----------------------------------------------------
def test_leaf_known_answers():
    A = olin.MatrixOperator(np.array([[1, 2, 4], [3, 4, 5]], dtype=float))
    assert A * olin.FlatVector([2, 3, 4]) == olin.FlatVector([24, 38])
    D = olin.DiagonalMatrixOperator(np.array([1, 2, 3]))
    assert D * olin.FlatVector([3, 4, 5]) == olin.FlatVector([3, 8, 15])


def test_multivector_reference_tests():
    FV, MV = olin.FlatVector, osg.MultiVector
    mv = MV()
    mv[Multiindex([1, 2, 3])] = FV([3, 4, 5])
    mv[Multiindex([0, 1, 2, 3])] = FV([3, 4, 6])
    mv[Multiindex([1, 2, 3, 0])] = FV([3, 4, 7])
    assert mv[Multiindex([1, 2, 3])] == FV([3, 4, 7]) and mv[Multiindex([0, 1, 2, 3])] == FV([3, 4, 6])
    assert [list(k.as_array) for k in mv.active_indices()] == [[0, 1, 2, 3], [1, 2, 3]]
    mv2 = mv.copy()
    mv[Multiindex([1, 2, 3])].coeffs[2] = 8
    assert mv2[Multiindex([1, 2, 3])] == FV([3, 4, 7])
    mis = MultiindexSet.createCompleteOrderSet(3, 4)
    a, b, c = MV(), MV(), MV()
    a.set_defaults(mis, FV([3, 4, 5])); b.set_defaults(mis, FV([6, 8, 12])); c.set_defaults(mis, FV([9, 12, 17]))
    assert a + b == c and a[Multiindex()] == FV([3, 4, 5])
    with pytest.raises(KeyError):
        a[Multiindex([1, 2, 2])]
    d = MV(); d.set_defaults(mis, FV([6, 8, 10]))
    assert 2 * a == d and a * 2.0 == d
    assert (-a)[Multiindex(mis[-1])] == FV([-3, -4, -5])
    assert (d - a) == a


# ---- apply: the reference's closed-form known answer ------------------------------------------------
def test_apply_closed_form():
    # test_multi_operator.py:57-90
    N = 4
    rvs = [op.UniformRV(), op.NormalRV(mu=0.5)]
    consts = {"mean": 2.0, 0: 3.0, 1: 4.0}
    cf = osg.ListCoefficientField("mean", [0, 1], rvs)

    def diag_assemble(basis, func):
        return olin.DiagonalMatrixOperator(np.full(basis.dim, abs(consts[func])), basis, basis)

    A = osg.MultiOperator(cf, diag_assemble)
    mis = [Multiindex([0]), Multiindex([1]), Multiindex([0, 1]), Multiindex([0, 2])]
    rng = np.random.RandomState(0)
    basis = olin.CanonicalBasis(N)
    vecs = [olin.FlatVector(rng.random_sample(N), basis) for _ in mis]
    w = osg.MultiVector()
    for mi, v in zip(mis, vecs):
        w[mi] = v
    v = A * w
    L = op.LegendrePolynomials(normalised=True)
    H = op.StochasticHermitePolynomials(mu=0.5, normalised=True)
    v0 = (2 * vecs[0] + 3 * (L.get_beta(0)[1] * vecs[1] - L.get_beta(0)[0] * vecs[0])
          + 4 * (H.get_beta(0)[1] * vecs[2] - H.get_beta(0)[0] * vecs[0]))
    v2 = (2 * vecs[2] + 4 * (H.get_beta(1)[1] * vecs[3] - H.get_beta(1)[0] * vecs[2]
                             + H.get_beta(1)[-1] * vecs[0]))
    assert np.allclose(v[mis[0]].coeffs, v0.coeffs, rtol=1e-15, atol=0)
    assert np.allclose(v[mis[2]].coeffs, v2.coeffs, rtol=1e-15, atol=0)
    with pytest.raises(TypeError):
        osg.MultiOperator(3, diag_assemble)
    with pytest.raises(TypeError):
        osg.MultiOperator(cf, 7)


def test_apply_loop_equals_batched_and_triples():
    from tests.cases import fd_case
    from spuq_b200 import synthetic
    idx = synthetic.complete_order_indices(3, 2)
    c = fd_case(5, 3, idx, kind="hermite", mu=0.5)
    Lam = c.w.active_indices()
    Ks = osg.stochastic_matrices(Lam, c.orvs, 3)
    Yb = osg.apply_batched(c.mats[0], c.mats[1:], Ks, c.W)
    assert np.max(np.abs(Yb - c.Y)) <= 1e-13 * np.max(np.abs(c.Y))
    # structural cross-check against the reference's Delta matrices: same
    # sparsity pattern; off-diagonals agree up to the row/column-degree ulp asymmetry
    T = osg.evaluate_triples([rv.orth_polys for rv in c.orvs], idx.astype(np.int64), idx.astype(np.int64))
    assert abs(T[0] - np.eye(len(Lam))).max() == 0
    for m in range(3):
        K = Ks[m].toarray(); D = T[m + 1].toarray()
        off = ~np.eye(len(Lam), dtype=bool)
        assert np.array_equal(K[off] != 0, D[off] != 0)
        assert np.allclose(K[off], D[off], rtol=1e-14)
        assert np.allclose(np.diag(K), -np.diag(D))      # sign convention differs on the diagonal


# ---- pcg: test_pcg.py ----------------------------------------------------------------------------
def test_pcg_reference_test():
    rand = np.random.mtrand.RandomState(1234).random_sample
    N = 7
    M = rand((N, N))
    A = olin.MatrixOperator(np.dot(M, M.T))
    P = olin.MultiplicationOperator(1, olin.CanonicalBasis(N))
    x = olin.FlatVector(rand((N,)))
    b = A * x
    x_ap, zeta, it = osg.pcg(A, b, P, 0 * x, eps=0.00001)
    np.testing.assert_array_almost_equal(x.coeffs, x_ap.coeffs)
    assert it <= N + 3
    P = olin.MatrixSolveOperator(np.dot(M, M.T))
    x_ap, zeta, it = osg.pcg(A, b, P, 0 * x, eps=0.00001)
    np.testing.assert_array_almost_equal(x.coeffs, x_ap.coeffs)
    assert it <= 2
    P = olin.MatrixOperator(P.as_matrix())
    x_ap, zeta, it = osg.pcg(A, b, P, 0 * x, eps=0.00001)
    np.testing.assert_array_almost_equal(x.coeffs, x_ap.coeffs)
    P = olin.DiagonalMatrixOperator(np.diag(A.as_matrix())).inverse()
    x_ap, zeta, it = osg.pcg(A, b, P, 0 * x, eps=0.00001)
    np.testing.assert_array_almost_equal(x.coeffs, x_ap.coeffs)


def test_pcg_failure_messages():
    N = 4
    basis = olin.CanonicalBasis(N)
    I = olin.MultiplicationOperator(1, basis)
    x = olin.FlatVector(np.ones(N))
    with pytest.raises(Exception, match="Matrix for PCG is not positive definite"):
        osg.pcg(olin.MatrixOperator(-np.eye(N)), x, I, 0 * x)
    with pytest.raises(Exception, match="Matrix for PCG is singular"):
        osg.pcg(olin.MatrixOperator(np.zeros((N, N))), x, I, 0 * x)
    with pytest.raises(Exception, match="Preconditioner for PCG is not positive definite"):
        osg.pcg(olin.MatrixOperator(np.eye(N)), x, olin.MultiplicationOperator(-1, basis), 0 * x)
    with pytest.raises(Exception, match="PCG did not converge"):
        osg.pcg(olin.MatrixOperator(np.diag([1.0, 10.0, 100.0, 1000.0])), x, I, 0 * x, eps=1e-14, maxiter=3)
