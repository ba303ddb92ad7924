"""Growth of the active index set between solves.

``Marking.refine_y(w, new_mi)`` of spuq/application/egsz/marking2.py:205-208:
every new multi-index gets a zero vector on the shared basis.  On a ``SlabMultiVector`` all new
columns are inserted with one sorted re-layout on the device (``insert_zero_columns``); the next
``MultiOperator.apply`` / ``pcg`` builds the neighbour and beta tables of the grown set from scratch
(`LambdaTables`, keyed by the index set), so they are bit-identical to those of a fresh vector.
Choosing WHICH indices to add (``mark_y``, marking2.py:69-201) needs the residual estimators and is
out of scope.
"""
from .linalg import FlatVector
from .multi_vector import MultiVector, SlabMultiVector

import numpy as np

__all__ = ["Marking"]


class Marking:
    @classmethod
    def refine_y(cls, w, new_mi):
        if not isinstance(w, MultiVector) or not isinstance(new_mi, (list, tuple)):
            raise TypeError("refine_y(MultiVector, list of Multiindex)")
        if isinstance(w, SlabMultiVector):
            w.insert_zero_columns(new_mi)
            return
        some = w[w.active_indices()[0]]
        for mu in new_mi:                                   # marking2.py:207-208
            w[mu] = FlatVector(np.zeros(some.dim), some.basis)
