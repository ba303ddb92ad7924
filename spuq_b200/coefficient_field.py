"""Coefficient-field containers ``(mean_func, [(a_m, rv_m)])``.

Mirrors spuq/application/egsz/coefficient_field.py:14-147.  The functions a_m are opaque to this
package: they are only handed back to the ``assemble`` callbacks of ``MultiOperator``.
"""
import sys

from .random_variable import RandomVariable

__all__ = ["CoefficientField", "ListCoefficientField", "ParametricCoefficientField"]


class CoefficientField:
    @property
    def mean_func(self):  # pragma: no cover - abstract
        raise NotImplementedError

    def __len__(self):  # pragma: no cover - abstract
        raise NotImplementedError

    def __getitem__(self, i):  # pragma: no cover - abstract
        raise NotImplementedError

    def sample_realization(self, Lambda, RV_samples):
        """``sample_map[mu] = prod_m P_{mu_m}(y_m)`` over the stored entries of mu, and the samples
        back (coefficient_field.py:45-52; the reference's ``sample_rvs`` default is broken there,
        :103-104, so the samples must be given)."""
        sample_map = {}
        for mu in Lambda:
            vals = [self[m][1].orth_polys[mu[m]](RV_samples[m]) for m in range(len(mu))]
            p = 1.0
            for v in vals:
                p = p * v
            sample_map[mu] = float(p)
        return sample_map, RV_samples


class ListCoefficientField(CoefficientField):
    def __init__(self, mean_func, funcs, rvs):
        funcs, rvs = list(funcs), list(rvs)
        if not all(isinstance(rv, RandomVariable) for rv in rvs):
            raise TypeError("rvs must be RandomVariable instances")
        assert len(funcs) == len(rvs)
        self._mean_func, self._funcs, self._rvs = mean_func, funcs, rvs

    @classmethod
    def create_with_iid_rvs(cls, mean_func, funcs, rv):
        funcs = list(funcs)
        return cls(mean_func, funcs, [rv] * len(funcs))

    @property
    def mean_func(self):
        return self._mean_func

    @property
    def funcs(self):
        return self._funcs

    @property
    def rvs(self):
        return self._rvs

    def __len__(self):
        return len(self._funcs)

    def __getitem__(self, i):
        assert i < len(self._funcs), "invalid index"
        return self._funcs[i], self._rvs[i]

    def __repr__(self):
        return "<%s mean=%s funcs=%s, rvs=%s>" % (type(self).__name__, self.mean_func,
                                                  self._funcs, self._rvs)


class _Lazy:
    """Memoising view of a function or generator of terms (utils/parametric_array.py)."""

    def __init__(self, source):
        self._vals = []
        self._source = source
        self._it = None if callable(source) else iter(source)

    def __getitem__(self, i):
        while len(self._vals) <= i:
            k = len(self._vals)
            self._vals.append(self._source(k) if self._it is None else next(self._it))
        return self._vals[i]


class ParametricCoefficientField(CoefficientField):
    """Unbounded expansion given by generators; ``len`` is "infinite" (coefficient_field.py:139-141),
    so ``maxm`` of the operator is decided by the multi-vector alone."""

    def __init__(self, mean_func, func_func, rv_func):
        self._mean_func = mean_func
        self._funcs = _Lazy(func_func)
        self._rvs = _Lazy(rv_func)

    @classmethod
    def create_with_iid_rvs(cls, mean_func, func_func, rv):
        return cls(mean_func, func_func, lambda i: rv)

    @property
    def mean_func(self):
        return self._mean_func

    @property
    def funcs(self):
        return self._funcs

    @property
    def rvs(self):
        return self._rvs

    def __len__(self):
        return sys.maxsize

    def __getitem__(self, i):
        return self._funcs[i], self._rvs[i]
