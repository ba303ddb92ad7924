"""Oracle: three-term recurrences of orthogonal polynomial families and the beta weights.

Restates (same floating-point operation order, so results are bit-identical):
  * reference src/spuq/polyquad/_polynomials.py:83-87 (rc4_to_rc3), :103-119 (normalise_rc),
    :122-143 (sqnorm_from_rc), :146-164 (shift/scale, window transform),
    :227-240 (stochastic Hermite), :255-276 (Legendre), :279-292 (Jacobi), :310-326 (Chebyshev)
  * reference src/spuq/polyquad/polynomials.py:44-48 (get_beta), :77-105 (family base),
    :125-150 (Legendre / StochasticHermite constructors)
  * reference src/spuq/stochastics/random_variable.py:138-182 (rv.orth_polys)
Test infrastructure only (see oracle/__init__.py).

Convention: p_{n+1} = (a_n + b_n x) p_n - c_n p_{n-1};  get_beta(n) = (a/b, 1/b, c/b) so that
x p_n = beta[1] p_{n+1} - beta[0] p_n + beta[-1] p_{n-1}.
"""
import math

import numpy as np

import scipy.special


def _factorial(n):
    # the reference calls scipy.factorial(n) (inexact, gamma-based) and converts to float
    # (_polynomials.py:238-240); scipy.special.factorial is that function's current home.
    return float(scipy.special.factorial(n))


# ---- raw (un-normalised) recurrences ------------------------------------------------------
def rc4_to_rc3(a):
    lead = float(a[0])
    return (a[1] / lead, a[2] / lead, a[3] / lead)


def rc_stoch_hermite(n):
    return (0.0, 1.0, float(n))


def sqnorm_stoch_hermite(n):
    return _factorial(n)


def rc_legendre(n):
    n = float(n)
    return rc4_to_rc3((max(0.0, n) + 1, 0.0, 2 * n + 1, n))


def sqnorm_legendre(n):
    n = float(n)
    return 1.0 / (2 * n + 1)


def rc_norm_legendre(n):
    """Analytic normalised Legendre pair (alpha, beta); _polynomials.py:261-268."""
    n = float(n)
    return (0.0, (4 - n ** -2) ** -0.5 if n > 0 else 0.0)


def rc_jacobi(n, alpha, beta):
    n = float(n)
    if n == 0:
        return ((alpha - beta) / 2.0, (alpha + beta + 2) / 2.0, 0)
    s = 2 * n + alpha + beta
    a1 = 2 * (n + 1) * (n + alpha + beta + 1) * s
    a2 = (s + 1) * (alpha ** 2 - beta ** 2)
    a3 = (s + 1) * s * (s + 2)
    a4 = 2 * (n + alpha) * (n + beta) * (s + 2)
    return rc4_to_rc3((a1, a2, a3, a4))


def rc_chebyshev_t(n):
    return (0.0, 1.0, 0.0) if n == 0 else (0.0, 2.0, 1.0)


def rc_chebyshev_u(n):
    return (0.0, 2.0, 0.0) if n == 0 else (0.0, 2.0, 1.0)


# ---- transforms ---------------------------------------------------------------------------
def sqnorm_from_rc(rc_func, n):
    """h_n = b_0/b_n * c_1 ... c_n (Law & Sledd); _polynomials.py:122-143."""
    h = rc_func(0)[1] / rc_func(n)[1]
    for i in range(1, n + 1):
        h *= rc_func(i)[2]
    return h


def rc_shift_scale(rc_func, shift, scale):
    def shifted(n):
        a, b, c = rc_func(n)
        return (a - b * shift / scale, b / scale, c)
    return shifted


def rc_window_trans(rc_func, old_domain, new_domain):
    a0, b0 = old_domain
    a1, b1 = new_domain
    shift = 0.5 * ((a1 + b1) - (a0 + b0))
    scale = float(a1 - b1) / (a0 - b0)
    return rc_shift_scale(rc_func, shift, scale)


def normalise_rc(rc_func, sqnorm_func=None):
    if sqnorm_func is None:
        def sqnorm_func(n):
            return sqnorm_from_rc(rc_func, n)

    def normalised(n):
        a, b, c = rc_func(n)
        h2 = sqnorm_func(n + 1) ** 0.5
        h1 = sqnorm_func(n) ** 0.5
        h0 = sqnorm_func(n - 1) ** 0.5 if n > 0 else 0
        return (a * h1 / h2, b * h1 / h2, c * h0 / h2)
    return normalised


def compute_poly(rc_func, n, x):
    """Values p_0..p_n at x via the recurrence; _polynomials.py:167-173."""
    f = [0 * x, 0 * x + 1]
    for i in range(n):
        a, b, c = rc_func(i)
        f.append((x * b + a) * f[i + 1] - c * f[i])
    return f[1:]


# ---- families -----------------------------------------------------------------------------
class PolynomialFamily:
    """polynomials.py:10-123 reduced to what the SG operator needs."""

    def __init__(self, rc_func, sqnorm_func=None, normalised=False):
        self._rc = rc_func
        self._sqnorm = sqnorm_func if sqnorm_func is not None else (
            lambda n: sqnorm_from_rc(rc_func, n))
        self._normalised = False
        if normalised:
            self._rc = normalise_rc(self._rc, self._sqnorm)
            self._sqnorm = None
            self._normalised = True

    def recurrence_coefficients(self, n):
        return self._rc(n)

    def get_beta(self, n):
        a, b, c = self._rc(n)
        return (a / b, 1 / b, c / b)

    def eval(self, n, x, all_degrees=False):
        vals = compute_poly(self._rc, n, x)
        return vals if all_degrees else vals[-1]

    def __getitem__(self, n):
        """The polynomial of degree n as a numpy.poly1d; polynomials.py:33-35."""
        return self.eval(n, np.poly1d([1, 0]))

    def norm(self, n, sqrt=True):
        if self._normalised:
            return 1.0
        return math.sqrt(self._sqnorm(n)) if sqrt else self._sqnorm(n)

    @property
    def normalised(self):
        return self._normalised


class LegendrePolynomials(PolynomialFamily):
    def __init__(self, a=-1.0, b=1.0, normalised=True):
        rc, sq = rc_legendre, sqnorm_legendre
        if a != -1.0 or b != 1.0:
            rc, sq = rc_window_trans(rc, (-1, 1), (a, b)), None
        super().__init__(rc, sq, normalised)


class StochasticHermitePolynomials(PolynomialFamily):
    def __init__(self, mu=0.0, sigma=1.0, normalised=True):
        rc, sq = rc_stoch_hermite, sqnorm_stoch_hermite
        if mu != 0.0 or sigma != 1.0:
            rc, sq = rc_shift_scale(rc, mu, sigma), None
        super().__init__(rc, sq, normalised)


class JacobiPolynomials(PolynomialFamily):
    def __init__(self, alpha, beta, a=-1.0, b=1.0, normalised=True):
        if alpha <= -1 or beta <= -1:
            raise TypeError("alpha and beta must be larger than -1")
        rc = lambda n: rc_jacobi(n, alpha, beta)  # noqa: E731
        if a != -1.0 or b != 1.0:
            rc = rc_window_trans(rc, (-1.0, 1.0), (a, b))
        super().__init__(rc, None, normalised)


# ---- random variables: only the orth_polys hook matters on the hot path -----------------------
class RandomVariable:
    pass


class NormalRV(RandomVariable):
    def __init__(self, mu=0, sigma=1):
        self.mu, self.sigma = float(mu), float(sigma)

    @property
    def orth_polys(self):
        return StochasticHermitePolynomials(self.mu, self.sigma, normalised=True)


class UniformRV(RandomVariable):
    def __init__(self, a=-1, b=1):
        self.a, self.b = float(min(a, b)), float(max(a, b))

    @property
    def orth_polys(self):
        return LegendrePolynomials(self.a, self.b, normalised=True)
