"""CPU oracle for the spuq stochastic-Galerkin hot path.  TEST INFRASTRUCTURE ONLY.

This package is a Python-3 / numpy / scipy restatement of the reference's algorithm for
``MultiOperator.apply`` and ``pcg`` (and the data model underneath them).  It exists so the CUDA
path in ``spuq_b200`` can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
``spuq_b200`` never does (tests/test_no_oracle_in_product.py enforces that).

Why a restatement: the reference is Python 2 and cannot be imported under the only interpreter in
the image.  Two reference files do load unmodified through a shim
(``oracle/ref_shim.py``): ``math_utils/multiindex_set.py`` and ``polyquad/{_polynomials,polynomials}.py``.
``oracle/make_golden.py`` runs those to pin this restatement (multi-index sets bit-exact, beta
weights bit-exact) and writes the fixtures in ``tests/golden/``.  The remaining functions are pinned
by the reference's own known-answer tests (cited in each docstring), restated in
``tests/test_oracle_*.py``.

Parity status:
  * multi-index construction / ordering / neighbours : pinned (live reference + goldens)
  * beta recurrence weights (Legendre, Hermite)      : pinned (live reference + goldens)
  * MultiOperator.apply arithmetic                   : pinned by the reference's closed-form test
    (application/egsz/tests/test_multi_operator.py:57-90)
  * pcg                                              : pinned by application/egsz/tests/test_pcg.py:12-44
  * CSR leaf operator (ScipyOperator)                : parity unpinned in the reference (its only
    exercisers need dolfin); pinned here by scipy 1.18.1 ``csr_matrix @ ndarray``.
"""
