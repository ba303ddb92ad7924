"""Oracle: multi-indices and complete-order multi-index sets (integer, bit-exact).

Restates
  * ``Multiindex``      reference src/spuq/math_utils/multiindex.py:13-138
  * ``MultiindexSet``   reference src/spuq/math_utils/multiindex_set.py:9-73
Test infrastructure only (see oracle/__init__.py).
"""
from hashlib import sha1

import numpy as np


class Multiindex:
    """Hashable multi-index with the trailing zeros stripped.

    multiindex.py:15-24 (constructor: copy, reject non-integer dtype), :26-32 (strip trailing
    zeros), :34-37 (equality), :42-53 + :64-71 (graded order), :73-76 (sha1 hash),
    :85-89 (item access returns 0 outside the tail), :103-127 (inc/dec, None when negative).
    """

    def __init__(self, arr=None):
        if arr is None or len(arr) == 0:
            arr = [0]
        arr = np.array(arr)  # always a private copy
        # the reference tests ``issubclass(dtype.type, int)`` which under Python 2 accepts the
        # platform integer; the intent (test_multiindex.py:13-14) is "integers yes, floats no".
        if not np.issubdtype(arr.dtype, np.integer):
            raise TypeError("multi-index entries must be integers")
        nz = np.nonzero(arr)[0]
        self._arr = arr[: nz.max() + 1].copy() if len(nz) else np.resize(arr, 0)
        self._hash = None

    # -- identity -------------------------------------------------------------------------
    def __eq__(self, other):
        return (type(self) is type(other) and hash(self) == hash(other)
                and np.array_equal(self._arr, other._arr))

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        # multiindex.py:73-76 hashes the raw array bytes, which makes the value depend on the
        # dtype; normalising to int64 keeps Multiindex([1]) == Multiindex(np.int8([1])) hash-equal,
        # which is what the reference relies on when it builds indices from int8 set rows.
        if self._hash is None:
            self._hash = int(sha1(np.ascontiguousarray(self._arr, dtype=np.int64)).hexdigest(), 16)
        return self._hash

    # -- ordering: total degree first, then entries left to right (lower entry first) ---------
    def cmp_by_order(self, other):
        assert type(self) is type(other)
        a, b = self._arr, other._arr
        sa, sb = int(a.sum()), int(b.sum())
        c = (sa > sb) - (sa < sb)
        if c == 0:
            for i in range(min(len(a), len(b))):
                c = (int(a[i]) > int(b[i])) - (int(a[i]) < int(b[i]))
                if c:
                    break
        return c

    def __le__(self, other):
        return self.cmp_by_order(other) <= 0

    def __gt__(self, other):
        return not self <= other

    # utils/decorators.py:76-110 derives the other two from __le__
    def __ge__(self, other):
        return other <= self

    def __lt__(self, other):
        return not other <= self

    # -- access ---------------------------------------------------------------------------
    def __len__(self):
        return len(self._arr)

    def __getitem__(self, i):
        return self._arr[i] if 0 <= i < len(self._arr) else 0

    def __repr__(self):
        return "<Multiindex inds=%s>" % (self._arr,)

    @property
    def order(self):
        return sum(self._arr)

    @property
    def as_array(self):
        return self._arr

    @property
    def supp(self):
        return np.flatnonzero(self._arr)

    def inc(self, pos, by=1):
        assert pos >= 0
        cur = self._arr[pos] if pos < len(self._arr) else 0
        new = cur + by
        if new < 0:
            return None
        n = len(self._arr)
        out = np.zeros(max(n, pos + 1), dtype=np.int64)
        out[:n] = self._arr
        out[pos] = new
        return type(self)(out)

    def dec(self, pos, by=1):
        return self.inc(pos, -by)


class MultiindexSet:
    """Array-backed set of multi-indices; multiindex_set.py:9-42."""

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
        out = vec[0] ** self.arr[:, 0]
        for i in range(1, vec.size):
            out = out * vec[i] ** self.arr[:, i]
        return out

    @staticmethod
    def _grow(m, func):
        """multiindex_set.py:44-53: endless generator in graded order."""
        p, done = 0, 0
        while True:
            mis = func(m, p)
            for i in range(done, len(mis)):
                yield mis[i]
            done = len(mis)
            p += 1

    @classmethod
    def createCompleteOrderSet(cls, m, p=None, reversed=False):
        """multiindex_set.py:55-73.  All mu in N^m with |mu| <= p.

        Built degree by degree; inside a degree the first m-1 components run through the
        (m-1)-dimensional set of the same recursion and the last component takes the remainder.
        """
        if p is None:
            return cls._grow(m, cls.createCompleteOrderSet)

        def build(dim, deg):
            if dim == 0:
                return np.zeros((1, 0), np.int8)
            rows = np.zeros((0, dim), np.int8)
            for q in range(deg + 1):
                head = build(dim - 1, q)
                tail = q - head.sum(1).reshape((head.shape[0], 1))
                rows = np.vstack((rows, np.hstack((head, tail))))
            return rows

        arr = build(m, p)
        if reversed:
            arr = arr[:, ::-1]
        return cls(arr)
