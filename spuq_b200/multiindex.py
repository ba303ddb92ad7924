"""Multi-indices and multi-index sets (host side, integer, bit-exact with the reference).

Mirrors the reference API
  * ``Multiindex``      spuq/math_utils/multiindex.py:13-138
  * ``MultiindexSet``   spuq/math_utils/multiindex_set.py:9-73
with a different representation: an index is an immutable tuple with the trailing zeros removed
(the reference keeps a numpy array and hashes its bytes with sha1).  Bulk work on whole sets --
sorting into ``active_indices()`` order and the mu +/- e_m neighbour tables -- is not done here but
in the C++ set-up code (``spuq_sg_lambda_sort`` / ``spuq_sg_lambda_neighbours``, see
``lambda_tables.py``).
"""
from itertools import count

import numpy as np

__all__ = ["Multiindex", "MultiindexSet"]


def _strip(values):
    n = len(values)
    while n and values[n - 1] == 0:
        n -= 1
    return tuple(values[:n])


class Multiindex:
    """Finitely supported multi-index mu in N_0^infinity.

    Same observable behaviour as the reference class: trailing zeros are irrelevant for equality
    and hashing (multiindex.py:26-37), ``mu[i]`` is 0 outside the stored tail (:85-89),
    ``inc``/``dec`` return a new index or ``None`` when an entry would become negative (:103-127),
    ordering is total degree first, then entries left to right with the lower entry first (:42-71).
    """

    __slots__ = ("_t", "_deg")

    def __init__(self, arr=None):
        if arr is None or len(arr) == 0:
            vals = ()
        else:
            a = np.asarray(arr)
            if not np.issubdtype(a.dtype, np.integer):
                raise TypeError("multi-index entries must be integers, got dtype %s" % a.dtype)
            if a.ndim != 1:
                raise TypeError("multi-index must be one-dimensional")
            vals = _strip([int(v) for v in a])
        self._t = vals
        self._deg = sum(vals)

    @classmethod
    def _from_tuple(cls, t):
        self = object.__new__(cls)
        self._t = t
        self._deg = sum(t)
        return self

    # identity
    def __eq__(self, other):
        return type(self) is type(other) and self._t == other._t

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash(self._t)

    # order: (degree, entries); tuples compare left to right and two different stripped tuples of
    # equal degree always differ inside their common prefix, so this is cmp_by_order exactly
    def _key(self):
        return (self._deg, self._t)

    def cmp_by_order(self, other):
        assert type(self) is type(other)
        a, b = self._key(), other._key()
        return (a > b) - (a < b)

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self._key() <= other._key()

    def __gt__(self, other):
        return self._key() > other._key()

    def __ge__(self, other):
        return self._key() >= other._key()

    # access
    def __len__(self):
        return len(self._t)

    def __getitem__(self, i):
        return self._t[i] if 0 <= i < len(self._t) else 0

    def __repr__(self):
        return "<Multiindex inds=%s>" % (self.as_array,)

    @property
    def order(self):
        return self._deg

    @property
    def as_array(self):
        return np.array(self._t, dtype=np.int64)

    @property
    def supp(self):
        return np.flatnonzero(np.array(self._t, dtype=np.int64))

    def padded(self, width):
        """Entries as a list of exactly ``width`` integers (zeros appended)."""
        if len(self._t) > width:
            raise ValueError("multi-index longer than %d" % width)
        return list(self._t) + [0] * (width - len(self._t))

    def inc(self, pos, by=1):
        assert pos >= 0
        new = self[pos] + by
        if new < 0:
            return None
        vals = list(self._t) + [0] * max(0, pos + 1 - len(self._t))
        vals[pos] = new
        return Multiindex._from_tuple(_strip(vals))

    def dec(self, pos, by=1):
        return self.inc(pos, -by)

    @staticmethod
    def createCompleteOrderSet(m, p):
        """Set whose items are Multiindex objects (multiindex.py:132-138)."""
        base = MultiindexSet.createCompleteOrderSet(m, p)

        class _Typed(MultiindexSet):
            def __getitem__(self, i):
                return Multiindex(MultiindexSet.__getitem__(self, i))
        return _Typed(base.arr)


class MultiindexSet:
    """Rows of an integer array as a set of multi-indices; multiindex_set.py:9-42."""

    def __init__(self, arr):
        self.arr = arr

    @property
    def m(self):
        return self.arr.shape[1]

    @property
    def p(self):
        return max(self.arr.sum(1))

    @property
    def count(self):
        return len(self)

    def __len__(self):
        return self.arr.shape[0]

    def __getitem__(self, i):
        return self.arr[i]

    def __repr__(self):
        return "<MISet m={0}, p={1}, arr={2}>".format(self.m, self.p, self.arr)

    def power(self, vec):
        assert vec.size == self.m
        return np.prod(np.asarray(vec)[None, :] ** self.arr, axis=1)

    @classmethod
    def createCompleteOrderSet(cls, m, p=None, reversed=False):
        """All mu in N_0^m with |mu| <= p in the reference's row order (multiindex_set.py:55-73):
        graded; within degree q the leading m-1 components enumerate the (m-1)-dimensional complete
        set of degree <= q in that same order and the last component is the remainder.

        Built iteratively per dimension (the reference recurses and re-stacks); ``p=None`` gives the
        endless generator of :44-53.
        """
        if p is None:
            return cls._generate(m)
        if m == 0:
            arr = np.zeros((1, 0), dtype=np.int64)
            return cls(arr)
        # level[q] = rows of the d-dimensional set of EXACT degree q, in reference order
        # d = 1: single row [q]
        level = [np.array([[q]], dtype=np.int64) for q in range(p + 1)]
        for _ in range(2, m + 1):
            # exact degree q in d dims: for r = 0..q, heads of exact degree r (all lower degrees
            # first, reference order inside) with tail q - r
            new = []
            for q in range(p + 1):
                parts = [np.hstack([level[r], np.full((level[r].shape[0], 1), q - r, dtype=np.int64)])
                         for r in range(q + 1)]
                new.append(np.vstack(parts))
            level = new
        arr = np.vstack(level)
        if reversed:
            arr = arr[:, ::-1]
        return cls(arr)

    @classmethod
    def _generate(cls, m):
        done = 0
        for p in count():
            arr = cls.createCompleteOrderSet(m, p).arr
            for i in range(done, arr.shape[0]):
                yield arr[i]
            done = arr.shape[0]
