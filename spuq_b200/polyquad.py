"""Orthogonal polynomial families and the beta weights of the SG operator (host side, fp64).

Mirrors the reference API ``family.get_beta(n)`` / ``recurrence_coefficients(n)`` /
``normalised`` (spuq/polyquad/polynomials.py:10-150, spuq/polyquad/_polynomials.py) for the
families the SG operator meets: Legendre (UniformRV), stochastic Hermite (NormalRV), plus Jacobi and
Chebyshev for completeness.  A family here is a small table-driven object: recurrence triples
(a_n, b_n, c_n) with  p_{n+1} = (a_n + b_n x) p_n - c_n p_{n-1}  are produced by a ``Recurrence``
and memoised per degree.

Bit-exactness.  The weights end up in the kernels' coefficient tables, and the parity bar for the
tables is bit-exact, so every formula keeps the reference's operation order:
  normalisation   (a h1/h2, b h1/h2, c h0/h2), h_k = sqnorm(k) ** 0.5, h0 = 0 at n = 0
                  (_polynomials.py:103-119)
  beta            (a/b, 1/b, c/b)                            (polynomials.py:44-48)
  affine map      (a - b*shift/scale, b/scale, c)            (_polynomials.py:146-164)
``** 0.5`` is Python's pow on floats, as in the reference (not math.sqrt).
"""
import math

import numpy as np
import scipy.special

__all__ = ["PolynomialFamily", "LegendrePolynomials", "StochasticHermitePolynomials",
           "JacobiPolynomials", "ChebyshevT", "ChebyshevU", "beta_table"]


class Recurrence:
    """Raw three-term recurrence n -> (a, b, c) with an optional closed-form squared norm."""

    def __init__(self, triple, sqnorm=None):
        self._triple = triple
        self._sqnorm = sqnorm
        self._memo = {}

    def __call__(self, n):
        n = int(n)
        hit = self._memo.get(n)
        if hit is None:
            hit = self._memo[n] = self._triple(n)
        return hit

    def sqnorm(self, n):
        """h_n with h_0 = 1; closed form if known, else Law & Sledd's product
        h_n = b_0/b_n * c_1...c_n (_polynomials.py:122-143)."""
        if self._sqnorm is not None:
            return self._sqnorm(n)
        h = self(0)[1] / self(n)[1]
        for i in range(1, n + 1):
            h *= self(i)[2]
        return h

    def affine(self, shift, scale):
        """Recurrence for the weight w((x - shift)/scale); the closed-form norm is dropped, as in
        the reference (polynomials.py:129-131,143-145)."""
        def triple(n):
            a, b, c = self(n)
            return (a - b * shift / scale, b / scale, c)
        return Recurrence(triple)

    def window(self, old, new):
        (a0, b0), (a1, b1) = old, new
        shift = 0.5 * ((a1 + b1) - (a0 + b0))
        scale = float(a1 - b1) / (a0 - b0)
        return self.affine(shift, scale)

    def normalised(self):
        def triple(n):
            a, b, c = self(n)
            h2 = self.sqnorm(n + 1) ** 0.5
            h1 = self.sqnorm(n) ** 0.5
            h0 = self.sqnorm(n - 1) ** 0.5 if n > 0 else 0
            return (a * h1 / h2, b * h1 / h2, c * h0 / h2)
        return Recurrence(triple, sqnorm=lambda n: 1.0)


def _four_to_three(a1, a2, a3, a4):
    """Abramowitz-Stegun four-coefficient form a1 p_{n+1} = (a2 + a3 x) p_n - a4 p_{n-1}."""
    lead = float(a1)
    return (a2 / lead, a3 / lead, a4 / lead)


def _legendre(n):
    n = float(n)
    return _four_to_three(max(0.0, n) + 1, 0.0, 2 * n + 1, n)


def _hermite(n):
    return (0.0, 1.0, float(n))


def _jacobi(alpha, beta):
    def triple(n):
        n = float(n)
        if n == 0:
            return ((alpha - beta) / 2.0, (alpha + beta + 2) / 2.0, 0)
        s = 2 * n + alpha + beta
        return _four_to_three(2 * (n + 1) * (n + alpha + beta + 1) * s,
                              (s + 1) * (alpha ** 2 - beta ** 2),
                              (s + 1) * s * (s + 2),
                              2 * (n + alpha) * (n + beta) * (s + 2))
    return triple


class PolynomialFamily:
    def __init__(self, recurrence, normalised):
        self._raw = recurrence
        self._rc = recurrence.normalised() if normalised else recurrence
        self._normalised = bool(normalised)

    def recurrence_coefficients(self, n):
        return self._rc(n)

    def get_beta(self, n):
        """(beta[0], beta[1], beta[-1]) with x p_n = beta[1] p_{n+1} - beta[0] p_n + beta[-1] p_{n-1}."""
        a, b, c = self._rc(n)
        return (a / b, 1 / b, c / b)

    def eval(self, n, x, all_degrees=False):
        vals = [0 * x, 0 * x + 1]
        for i in range(n):
            a, b, c = self._rc(i)
            vals.append((x * b + a) * vals[i + 1] - c * vals[i])
        return vals[1:] if all_degrees else vals[-1]

    def __getitem__(self, n):
        return self.eval(n, np.poly1d([1, 0]))

    def norm(self, n, sqrt=True):
        if self._normalised:
            return 1.0
        h = self._raw.sqnorm(n)
        return math.sqrt(h) if sqrt else h

    @property
    def normalised(self):
        return self._normalised


class LegendrePolynomials(PolynomialFamily):
    """Legendre on [a, b] (uniform weight); polynomials.py:125-136."""

    def __init__(self, a=-1.0, b=1.0, normalised=True):
        rec = Recurrence(_legendre, sqnorm=lambda n: 1.0 / (2 * float(n) + 1))
        if a != -1.0 or b != 1.0:
            rec = rec.window((-1, 1), (a, b))
        super().__init__(rec, normalised)


class StochasticHermitePolynomials(PolynomialFamily):
    """Probabilists' Hermite for N(mu, sigma); polynomials.py:139-150."""

    def __init__(self, mu=0.0, sigma=1.0, normalised=True):
        # the reference takes n! from scipy's (inexact, gamma based) factorial, _polynomials.py:238-240
        rec = Recurrence(_hermite, sqnorm=lambda n: float(scipy.special.factorial(n)))
        if mu != 0.0 or sigma != 1.0:
            rec = rec.affine(mu, sigma)
        super().__init__(rec, normalised)


class JacobiPolynomials(PolynomialFamily):
    def __init__(self, alpha, beta, a=-1.0, b=1.0, normalised=True):
        if alpha <= -1 or beta <= -1:
            raise TypeError("alpha and beta must be larger than -1")
        rec = Recurrence(_jacobi(alpha, beta))
        if a != -1.0 or b != 1.0:
            rec = rec.window((-1.0, 1.0), (a, b))
        super().__init__(rec, normalised)


class ChebyshevT(PolynomialFamily):
    def __init__(self, a=-1.0, b=1.0, normalised=True):
        rec = Recurrence(lambda n: (0.0, 1.0, 0.0) if n == 0 else (0.0, 2.0, 1.0))
        if a != -1.0 or b != 1.0:
            rec = rec.window((-1, 1), (a, b))
        super().__init__(rec, normalised)


class ChebyshevU(PolynomialFamily):
    def __init__(self, a=-1.0, b=1.0, normalised=True):
        rec = Recurrence(lambda n: (0.0, 2.0, 0.0) if n == 0 else (0.0, 2.0, 1.0))
        if a != -1.0 or b != 1.0:
            rec = rec.window((-1, 1), (a, b))
        super().__init__(rec, normalised)


def beta_table(family, max_degree):
    """(max_degree+1) x 3 array of get_beta(n) = (beta0, beta_up, beta_dn) for n = 0..max_degree."""
    out = np.empty((max_degree + 1, 3), dtype=np.float64)
    for n in range(max_degree + 1):
        b = family.get_beta(n)
        out[n, 0], out[n, 1], out[n, 2] = b[0], b[1], b[2]
    return out
