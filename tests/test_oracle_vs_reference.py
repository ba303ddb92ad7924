"""Oracle and product host code against the LIVE reference files (skipped where /root/reference is
absent, e.g. on the GPU box; the committed goldens in tests/golden cover that case)."""
import warnings

import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ref_shim.load()


@pytest.mark.parametrize("m,p", [(1, 1), (2, 3), (3, 2), (4, 4), (5, 2), (7, 5), (20, 3), (0, 2), (6, 0)])
def test_sets_bit_exact(ref, m, p):
    from oracle.multiindex import MultiindexSet as O
    from spuq_b200.multiindex import MultiindexSet as P
    want = ref[0].MultiindexSet.createCompleteOrderSet(m, p).arr
    assert np.array_equal(O.createCompleteOrderSet(m, p).arr, want)
    assert np.array_equal(P.createCompleteOrderSet(m, p).arr, want)
    assert np.array_equal(P.createCompleteOrderSet(m, p, reversed=True).arr,
                          ref[0].MultiindexSet.createCompleteOrderSet(m, p, reversed=True).arr)


def test_generator(ref):
    from spuq_b200.multiindex import MultiindexSet as P
    g1, g2 = ref[0].MultiindexSet.createCompleteOrderSet(3), P.createCompleteOrderSet(3)
    for _ in range(60):
        assert np.array_equal(next(g1), next(g2))


def test_beta_bit_exact(ref):
    from oracle import polys as O
    from spuq_b200 import polyquad as P
    pol = ref[1]
    cases = [
        (pol.LegendrePolynomials(), O.LegendrePolynomials(), P.LegendrePolynomials()),
        (pol.LegendrePolynomials(0.0, 1.0), O.LegendrePolynomials(0.0, 1.0), P.LegendrePolynomials(0.0, 1.0)),
        (pol.LegendrePolynomials(-3.0, 0.5), O.LegendrePolynomials(-3.0, 0.5), P.LegendrePolynomials(-3.0, 0.5)),
        (pol.StochasticHermitePolynomials(), O.StochasticHermitePolynomials(), P.StochasticHermitePolynomials()),
        (pol.StochasticHermitePolynomials(0.5, 1.0), O.StochasticHermitePolynomials(0.5, 1.0),
         P.StochasticHermitePolynomials(0.5, 1.0)),
        (pol.StochasticHermitePolynomials(-1.0, 2.5), O.StochasticHermitePolynomials(-1.0, 2.5),
         P.StochasticHermitePolynomials(-1.0, 2.5)),
        (pol.JacobiPolynomials(0.5, 1.0, -2.0, 3.0), O.JacobiPolynomials(0.5, 1.0, -2.0, 3.0),
         P.JacobiPolynomials(0.5, 1.0, -2.0, 3.0)),
        (pol.LegendrePolynomials(normalised=False), O.LegendrePolynomials(normalised=False),
         P.LegendrePolynomials(normalised=False)),
    ]
    for r, o, p in cases:
        for n in range(40):
            want = tuple(float(v) for v in r.get_beta(n))
            assert tuple(float(v) for v in o.get_beta(n)) == want
            assert tuple(float(v) for v in p.get_beta(n)) == want
