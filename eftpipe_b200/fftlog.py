"""Biased log-sampled FFTLog: grids, exponents and the LINEAR OPERATORS the CUDA path applies.

Mirrors the reference class `eftpipe.pybird.fftlog.FFTLog` (fftlog.py:43-173): same
constructor, same attributes (`x`, `dx`, `Pow`, `_CoefFactor`), same error for odd `Nmax`.
Per-evaluation transforms do not run here: `FFTLog.operator()` returns the complex matrix
that maps input samples to the power-law coefficients c_m (spline resampling, damping,
DFT, Hermitian unfolding, coefficient factor, taper - all linear for fixed abscissae), and
the plan builder (`plan.py`) fuses it with whatever follows.  The only non-linear piece, the
power-law tail beyond the last input sample (fftlog.py:146-151), is generated on the device
from two scalars per row and enters through `tail_operator()`.

`coef_host()` is a host-side (numpy) evaluation used ONLY for one-off precomputation of
survey-window matrices (window.py:262-346), never on the per-evaluation path.
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import CubicSpline


def coef_window(N, window=1):
    """Taper sending the outer FFTLog coefficients smoothly to zero (fftlog.py:17-40)."""
    n = np.arange(-N // 2, N // 2 + 1)
    ncut = N // 2 if window == 1 else int(window * N // 2.0)
    right = n[-1] - ncut
    left = n[0] + ncut
    W = np.ones(n.size)
    hi = n > right
    th = (n[-1] - n[hi]) / float(n[-1] - right - 1)
    W[hi] = th - np.sin(2 * np.pi * th) / (2 * np.pi)
    lo = n < left
    th = (n[lo] - n[0]) / float(left - n[0] - 1)
    W[lo] = th - np.sin(2 * np.pi * th) / (2 * np.pi)
    return W


class FFTLog:
    def __init__(self, Nmax, xmin, xmax, bias):
        self.Nmax = int(Nmax)
        if self.Nmax % 2 != 0:
            raise ValueError(f"expected even Nmax, instead of Nmax={self.Nmax}")
        self.xmin, self.xmax, self.bias = xmin, xmax, bias
        self.dx = np.log(self.xmax / self.xmin) / (self.Nmax - 1.0)
        self.x = np.array([self.xmin * np.exp(i * self.dx) for i in range(self.Nmax)])
        m = np.arange(self.Nmax + 1)
        self.Pow = self.bias + 1j * 2.0 * np.pi / (self.Nmax * self.dx) * (m - self.Nmax / 2.0)
        self._CoefFactor = self.xmin ** (-self.Pow) / float(self.Nmax)

    # ------------------------------------------------------------------ linear pieces
    def _post(self, window):
        """Per-coefficient multiplier applied after the DFT (fftlog.py:158-164)."""
        post = self._CoefFactor.copy()
        if window is not None:
            post = post * coef_window(self.Nmax, window)
        else:
            post[0] /= 2.0
            post[-1] /= 2.0
        return post

    def _dft(self, idx):
        """Rows m = 0..Nmax of the unfolded DFT restricted to grid columns `idx`:
        c_m = sum_i fx_i exp(-2 pi i (m - Nmax/2) i / Nmax)  (fftlog.py:153-157)."""
        m = np.arange(self.Nmax + 1) - self.Nmax // 2
        prod = np.mod(np.outer(m, idx), self.Nmax)
        return np.exp(-2j * np.pi * prod / self.Nmax)

    def support(self, xin):
        il = int(np.searchsorted(self.x, xin[0]))
        ir = int(np.searchsorted(self.x, xin[-1], side="right"))
        return il, ir

    def operator(self, xin, window=1):
        """Complex matrix L (Nmax+1, len(xin)) with  Coef = L @ f  for the part of the
        transform supported inside [xin[0], xin[-1]] (both tails zero, i.e. "padding")."""
        xin = np.asarray(xin, dtype=float)
        il, ir = self.support(xin)
        idx = np.arange(il, ir)
        # not-a-knot cubic resampling as a matrix (scipy CubicSpline is linear in y)
        S = CubicSpline(xin, np.eye(xin.size), axis=0, extrapolate=False)(self.x[il:ir])
        damp = np.exp(-self.bias * idx * self.dx)
        return (self._post(window)[:, None] * self._dft(idx)) @ (damp[:, None] * S)

    def tail_operator(self, xin, window=1):
        """High-x power-law tail (fftlog.py:146-151):  fx_i = f_last (x_i/x_last)^n damp_i for
        x_i beyond xin[-1].  Returns (Lt, logratio): Coef += Lt @ (f_last * exp(n*logratio))."""
        xin = np.asarray(xin, dtype=float)
        _, ir = self.support(xin)
        idx = np.arange(ir, self.Nmax)
        damp = np.exp(-self.bias * idx * self.dx)
        Lt = (self._post(window)[:, None] * self._dft(idx)) * damp[None, :]
        return Lt, np.log(self.x[ir:] / xin[-1])

    def has_low_tail(self, xin):
        return bool(xin[0] > self.x[0])

    # ------------------------------------------------------------------ host evaluation
    def coef_host(self, xin, f, extrap="padding", window=1, kernel=None):
        """numpy evaluation of the coefficients, "padding" tails only; precompute use."""
        if extrap != "padding":
            raise ValueError("coef_host supports extrap='padding' only (plan-build use)")
        f = np.asarray(f, dtype=float)
        il, ir = self.support(xin)
        damp = np.exp(-self.bias * np.arange(il, ir) * self.dx)
        if kernel is not None:
            damp = damp * kernel(self.x[il:ir])
        fx = np.zeros(np.broadcast_shapes(f.shape[:-1], damp.shape[:-1]) + (self.Nmax,))
        fx[..., il:ir] = CubicSpline(xin, f, axis=-1, extrapolate=False)(self.x[il:ir]) * damp
        half = np.fft.rfft(fx, axis=-1)
        c = np.concatenate([np.conj(half[..., :0:-1]), half], axis=-1)
        return c * self._post(window)
