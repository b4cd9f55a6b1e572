"""Closed-form constants of the one-loop calculation, evaluated on the host at plan-build time.

Everything here is *precompute* (once per plan, numpy/scipy on the host, fp64 or extended
precision); nothing in this module runs per evaluation.  The formulas follow the reference
(eftpipe/pybird/pybird.py, cited per function) but are evaluated from the compact polynomial
tables in `data/pybird_tables.npz` (written by tools/extract_tables.py).
"""
from __future__ import annotations

import functools
import os

import numpy as np
from scipy.special import loggamma

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "pybird_tables.npz")


@functools.lru_cache(maxsize=1)
def raw_tables():
    return dict(np.load(_DATA))


def _poly2d_ld(coef, x, y):
    """sum_ij coef[i,j] x^i y^j in extended precision (Horner in both variables)."""
    x = x.astype(np.clongdouble)
    y = y.astype(np.clongdouble)
    out = np.zeros(np.broadcast(x, y).shape, dtype=np.clongdouble)
    for i in range(coef.shape[0] - 1, -1, -1):
        inner = np.zeros(y.shape, dtype=np.clongdouble)
        for j in range(coef.shape[1] - 1, -1, -1):
            inner = inner * y + np.longdouble(coef[i, j])
        out = out * x + inner
    return out


def _poly1d_ld(coef, x):
    x = x.astype(np.clongdouble)
    out = np.zeros(x.shape, dtype=np.clongdouble)
    for i in range(coef.shape[0] - 1, -1, -1):
        out = out * x + np.longdouble(coef[i])
    return out


def loop22_rational(nu):
    """The 28 rational kernels M22b[b](nu_n, nu_m) (pybird.py:119-148) -> (28, N, N) complex128.
    Numerator and denominator polynomials are summed in 80-bit arithmetic so that the rounded
    result is the correctly rounded value of the exact rational function."""
    T = raw_tables()
    a, b = nu[:, None], nu[None, :]
    out = np.empty((28, nu.size, nu.size), dtype=complex)
    for i in range(28):
        out[i] = (_poly2d_ld(T["m22_num"][i], a, b) / _poly2d_ld(T["m22_den"][i], a, b)).astype(complex)
    return out


def loop13_rational(nu):
    """M13b[b](nu_n) (pybird.py:98-109) -> (10, N)."""
    T = raw_tables()
    return np.array([(_poly1d_ld(T["m13_num"][i], nu) / _poly1d_ld(T["m13_den"][i], nu)).astype(complex)
                     for i in range(10)])


def loop22_gamma(nu):
    """Gamma-function prefactor M22a(nu_n, nu_m) (pybird.py:152-156)."""
    a, b = nu[:, None], nu[None, :]
    lg = loggamma(1.5 - a) + loggamma(1.5 - b) + loggamma(-1.5 + a + b)
    lg = lg - (loggamma(a) + loggamma(3.0 - a - b) + loggamma(b))
    return np.exp(lg) / (8.0 * np.pi**1.5)


def loop13_prefactor(nu):
    """M13a(nu) (pybird.py:112-114)."""
    return np.tan(nu * np.pi) / (14.0 * (-3 + nu) * (-2 + nu) * (-1 + nu) * nu * np.pi)


def bessel_power(l, pn):
    """MPC(l, pn): (1/2pi^2) int t^2 t^(-2pn) j_l(t) dt (pybird.py:159-173)."""
    return np.pi**-1.5 * 2.0 ** (-2.0 * pn) * np.exp(loggamma(1.5 + l / 2.0 - pn) - loggamma(l / 2.0 + pn))


def kbird(kmax=0.3):
    """Internal k nodes (pybird.py:472-479)."""
    if kmax > 0.30:
        low = np.array([0.001, 0.005, 0.0075, 0.01, 0.0125, 0.015, 0.0175, 0.02])
        ext = np.arange(low[-1], kmax + 1e-3, 0.005)
        return np.concatenate([low, ext[1:]])
    return raw_tables()["kbird"].copy()


def sbird():
    return raw_tables()["sbird"].copy()


# mu-power carried by every term (pybird.py:570-582) and the f-power grouping of
# `Bird.reducePsCfl` (pybird.py:762-846) as (row, f-power, term) triples
MU11 = (0, 2, 4)
MUCT = (0, 2, 4, 2, 4, 6)
MUNNLO = (4, 6, 8)
MU22 = (0,) * 6 + (2,) * 7 + (4, 2, 4, 2, 4, 2) + (4,) * 3 + (6, 4, 6, 4, 6, 8)
MU13 = (0,) * 2 + (2,) * 4 + (4,) * 3 + (6,)
GROUP22 = ((0, 2, 20), (0, 3, 23), (0, 3, 24), (0, 4, 25), (0, 4, 26), (0, 4, 27),
           (1, 1, 9), (1, 2, 14), (1, 2, 15), (1, 3, 21), (1, 3, 22),
           (2, 1, 10), (2, 2, 16), (2, 2, 17),
           (4, 1, 11), (4, 2, 18), (4, 2, 19),
           (5, 0, 0), (5, 1, 6), (5, 2, 12), (5, 2, 13),
           (6, 0, 1), (6, 1, 7), (8, 0, 2), (8, 1, 8), (9, 0, 3), (10, 0, 4), (11, 0, 5))
GROUP13 = ((0, 2, 7), (0, 3, 8), (0, 3, 9), (1, 1, 3), (1, 2, 5), (1, 2, 6), (3, 1, 4),
           (5, 0, 0), (5, 1, 2), (7, 0, 1))


def legendre_weights(Nl, powers):
    """Projection of mu^p onto the first Nl even Legendre multipoles, table `mu` of
    pybird.py:89-95 (kept verbatim including its 48/148 entry)."""
    mu = raw_tables()["mu_to_legendre"]
    return np.array([[mu[p // 2][l] for p in powers] for l in range(Nl)])


def resum_coefficients(Nl):
    """q[a, l, lp, u, d]: Q^{ll'}(f) polynomial coefficients in the ORDER USED by Resum.makeQ
    (pybird.py:1367-1380): a=0 -> linear part (table index 1), a=1 -> loop/counterterm."""
    T = raw_tables()
    if Nl == 3:
        q = T["q_nl3"]
    elif Nl == 2:
        q = T["q_nl2"]
    else:
        raise NotImplementedError("IR resummation tables exist for Nl = 2 or 3 only")
    return np.ascontiguousarray(q[::-1])
