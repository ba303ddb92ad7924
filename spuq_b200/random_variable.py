"""Random variables as far as the SG operator needs them: the ``orth_polys`` hook.

Mirrors spuq/stochastics/random_variable.py:138-182 (``NormalRV`` -> stochastic Hermite,
``UniformRV`` -> Legendre on [a, b]).  pdf / cdf / sampling of the reference class are outside the
hot path and not provided.  Unlike the reference, which builds a new polynomial-family object on
every ``rv.orth_polys`` access (and so throws its memo away L*M times per apply), the family is
created once per random variable.
"""
from . import polyquad

__all__ = ["RandomVariable", "NormalRV", "UniformRV"]


class RandomVariable:
    _polys = None

    @property
    def orth_polys(self):
        if self._polys is None:
            self._polys = self._make_polys()
        return self._polys

    def _make_polys(self):  # pragma: no cover - abstract
        raise NotImplementedError


class NormalRV(RandomVariable):
    def __init__(self, mu=0, sigma=1):
        self.mu = float(mu)
        self.sigma = float(sigma)

    def shift(self, delta):
        return NormalRV(self.mu + delta, self.sigma)

    def scale(self, scale):
        return NormalRV(self.mu, self.sigma * scale)

    def _make_polys(self):
        return polyquad.StochasticHermitePolynomials(self.mu, self.sigma, normalised=True)

    def __repr__(self):
        return "<NormalRV mu=%s sigma=%s>" % (self.mu, self.sigma)


class UniformRV(RandomVariable):
    def __init__(self, a=-1, b=1):
        self.a = float(min(a, b))
        self.b = float(max(a, b))

    def shift(self, delta):
        return UniformRV(a=self.a + delta, b=self.b + delta)

    def scale(self, scale):
        m = 0.5 * (self.a + self.b)
        d = scale * 0.5 * (self.b - self.a)
        return UniformRV(a=m - d, b=m + d)

    def _make_polys(self):
        return polyquad.LegendrePolynomials(self.a, self.b, normalised=True)

    def __repr__(self):
        return "<UniformRV a=%s b=%s>" % (self.a, self.b)
