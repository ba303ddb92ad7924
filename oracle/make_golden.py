"""Generate tests/golden/*.json from the UNMODIFIED reference files that load under Python 3.

Run in the development container (needs /root/reference):   python -m oracle.make_golden
Test infrastructure only.  The fixtures travel to the GPU box; the reference tree does not.

What is recorded
  multiindex_sets.json   createCompleteOrderSet(m, p) for the shapes the reference tests and
                         SURVEY.md section 8c name: full arrays for small sets, shape + SHA-256 of
                         the int32 C-contiguous bytes for large ones; the first rows of the
                         generator form; the reversed form.
  beta.json              get_beta(n), n = 0..30, as exact hex floats, for Legendre[-1,1],
                         Legendre[0,1], Hermite N(0,1), N(0.5,1), N(0,2), Jacobi(0.5,1;-2,3), all
                         normalised, from reference polyquad/polynomials.py via the shim.
"""
import hashlib
import json
import os
import warnings

import numpy as np

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mis, pol = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)

    sets = {}
    for m, p in [(1, 1), (2, 2), (2, 3), (3, 2), (0, 2), (5, 0), (3, 4), (5, 2), (7, 5), (20, 3), (30, 3), (4, 6)]:
        arr = mis.MultiindexSet.createCompleteOrderSet(m, p).arr
        a32 = np.ascontiguousarray(arr, dtype=np.int32)
        rec = {"shape": list(arr.shape), "dtype": str(arr.dtype),
               "sha256": hashlib.sha256(a32.tobytes()).hexdigest()}
        if arr.size <= 400:
            rec["rows"] = arr.tolist()
        sets["%d,%d" % (m, p)] = rec
    gen = mis.MultiindexSet.createCompleteOrderSet(2)
    sets["generator_2_first12"] = [next(gen).tolist() for _ in range(12)]
    sets["reversed_3_2"] = mis.MultiindexSet.createCompleteOrderSet(3, 2, reversed=True).arr.tolist()
    sets["repr_2_3_prefix"] = repr(mis.MultiindexSet.createCompleteOrderSet(2, 3))[:25]
    with open(os.path.join(OUT, "multiindex_sets.json"), "w") as f:
        json.dump(sets, f, indent=1, sort_keys=True)

    fams = {
        "legendre[-1,1]": pol.LegendrePolynomials(),
        "legendre[0,1]": pol.LegendrePolynomials(0.0, 1.0),
        "hermite(0,1)": pol.StochasticHermitePolynomials(),
        "hermite(0.5,1)": pol.StochasticHermitePolynomials(0.5, 1.0),
        "hermite(0,2)": pol.StochasticHermitePolynomials(0.0, 2.0),
        "jacobi(0.5,1;-2,3)": pol.JacobiPolynomials(0.5, 1.0, -2.0, 3.0),
    }
    beta = {}
    for name, fam in fams.items():
        rows = []
        for n in range(31):
            b = fam.get_beta(n)
            rows.append([float(b[0]).hex(), float(b[1]).hex(), float(b[-1]).hex()])
        beta[name] = rows
    with open(os.path.join(OUT, "beta.json"), "w") as f:
        json.dump(beta, f, indent=1, sort_keys=True)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
