"""CPU oracle: a numpy/scipy restatement of the reference's one-loop EFTofLSS hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (`eftpipe_b200/`) imports this file; it
is imported by `tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s CPU-baseline legs
as the *checker* / the *thing timed on the CPU*, never as a fallback of the CUDA path.

Parity pinning: the reference ships no golden vectors for this path,
so this oracle is pinned against the *live* reference: `tests/golden/make_golden.py` imports
the unmodified reference from /root/reference (through `oracle/refload.py`), runs both on the
same synthetic inputs, asserts agreement and commits the reference's outputs as
`tests/golden/*.npz`.  `tests/test_oracle_golden.py` re-checks the oracle against those
fixtures on every run (no /root/reference needed).

The functions follow the reference formulation one evaluation at a time (no batch axis; the
reference's FFTLog "extrap" branch is not batch-safe, fftlog.py:142-151), with the same
third-party primitives the reference calls (numpy rfft, scipy CubicSpline / interp1d /
loggamma).  File:line citations are relative to /root/reference/eftpipe/.
"""
from __future__ import annotations

import inspect
import os
from dataclasses import dataclass, field

import numpy as np
from numpy.fft import rfft
from scipy.interpolate import CubicSpline, interp1d
from scipy.special import eval_legendre, loggamma

_TABLES = None


def tables():
    global _TABLES
    if _TABLES is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "..", "eftpipe_b200", "data", "pybird_tables.npz")
        _TABLES = dict(np.load(path))
    return _TABLES


# --------------------------------------------------------------------------------------
# FFTLog (pybird/fftlog.py)
# --------------------------------------------------------------------------------------
def taper(N, window):
    """Edge taper of the FFTLog coefficients (fftlog.py:17-40, `CoefWindow`)."""
    n = np.arange(-N // 2, N // 2 + 1)
    ncut = N // 2 if window == 1 else int(window * N // 2.0)
    hi = n[-1] - ncut
    lo = n[0] + ncut
    W = np.ones(n.size)
    sel = n > hi
    th = (n[-1] - n[sel]) / float(n[-1] - hi - 1)
    W[sel] = th - np.sin(2 * np.pi * th) / (2 * np.pi)
    sel = n < lo
    th = (n[sel] - n[0]) / float(lo - n[0] - 1)
    W[sel] = th - np.sin(2 * np.pi * th) / (2 * np.pi)
    return W


@dataclass
class LogGrid:
    """FFTLog grid and exponents (fftlog.py:59-82)."""

    Nmax: int
    xmin: float
    xmax: float
    bias: float
    dx: float = field(init=False)
    x: np.ndarray = field(init=False)
    Pow: np.ndarray = field(init=False)
    factor: np.ndarray = field(init=False)

    def __post_init__(self):
        if self.Nmax % 2:
            raise ValueError(f"expected even Nmax, instead of Nmax={self.Nmax}")
        self.dx = np.log(self.xmax / self.xmin) / (self.Nmax - 1.0)
        self.x = np.array([self.xmin * np.exp(i * self.dx) for i in range(self.Nmax)])
        m = np.arange(self.Nmax + 1)
        self.Pow = self.bias + 1j * 2.0 * np.pi / (self.Nmax * self.dx) * (m - self.Nmax / 2.0)
        self.factor = self.xmin ** (-self.Pow) / float(self.Nmax)


def fftlog_coef(g: LogGrid, xin, f, extrap="extrap", window=1, kernel=None):
    """Power-law coefficients c_m of f(x) (fftlog.py:84-166).  `f` may carry leading axes
    only with extrap="padding" (the reference's "extrap" tails index the first axis)."""
    f = np.asarray(f, dtype=float)
    if not isinstance(extrap, tuple):
        extrap = (extrap, extrap)
    if any(e not in ("padding", "extrap") for e in extrap):
        raise ValueError(f"unexpected extrap = {extrap}")
    spline = CubicSpline(xin, f, axis=-1, extrapolate=False)
    lead = f.shape[:-1]
    fx = np.zeros(lead + (g.Nmax,))
    il = np.searchsorted(g.x, xin[0])
    ir = np.searchsorted(g.x, xin[-1], side="right")
    damp = np.exp(-g.bias * np.arange(il, ir) * g.dx)
    if kernel is not None:
        damp = damp * kernel(g.x[il:ir])
    fx[..., il:ir] = spline(g.x[il:ir]) * damp
    if extrap[0] == "extrap" and xin[0] > g.x[0]:
        assert f.ndim == 1
        slope = (np.log(f[1]) - np.log(f[0])) / (np.log(xin[1]) - np.log(xin[0]))
        amp = f[0] / xin[0] ** slope
        fx[:il] = amp * g.x[:il] ** slope * np.exp(-g.bias * np.arange(0, il) * g.dx)
    if extrap[1] == "extrap" and xin[-1] < g.x[-1]:
        assert f.ndim == 1
        slope = (np.log(f[-1]) - np.log(f[-2])) / (np.log(xin[-1]) - np.log(xin[-2]))
        amp = f[-1] / xin[-1] ** slope
        fx[ir:] = amp * g.x[ir:] ** slope * np.exp(-g.bias * np.arange(ir, g.Nmax) * g.dx)
    half = rfft(fx, axis=-1)
    c = np.empty(lead + (g.Nmax + 1,), dtype=complex)
    c[..., : g.Nmax // 2] = np.conj(half[..., 1:][..., ::-1])
    c[..., g.Nmax // 2 :] = half
    c *= g.factor
    if window is not None:
        c *= taper(g.Nmax, window)
    else:
        c[..., 0] /= 2.0
        c[..., g.Nmax] /= 2.0
    return c


# --------------------------------------------------------------------------------------
# constants (pybird.py:89-173, :472-582)
# --------------------------------------------------------------------------------------
def _polyval2d(c, x, y):
    out = np.zeros(np.broadcast(x, y).shape, dtype=complex)
    for i in range(c.shape[0]):
        for j in range(c.shape[1]):
            if c[i, j] != 0.0:
                out = out + c[i, j] * x**i * y**j
    return out


def m22_rational(b, n1, n2):
    """M22b[b](n1, n2) (pybird.py:119-148) from the extracted polynomial tables."""
    T = tables()
    return _polyval2d(T["m22_num"][b], n1, n2) / _polyval2d(T["m22_den"][b], n1, n2)


def m13_rational(b, n1):
    T = tables()
    num = sum(T["m13_num"][b, i] * n1**i for i in range(T["m13_num"].shape[1]))
    den = sum(T["m13_den"][b, i] * n1**i for i in range(T["m13_den"].shape[1]))
    return num / den


def m22_common(n1, n2):
    """pybird.py:152-156"""
    up = loggamma(1.5 - n1) + loggamma(1.5 - n2) + loggamma(-1.5 + n1 + n2)
    dn = loggamma(n1) + loggamma(3 - n1 - n2) + loggamma(n2)
    return np.exp(up) / (8.0 * np.pi**1.5 * np.exp(dn))


def m13_common(n1):
    """pybird.py:112-114"""
    return np.tan(n1 * np.pi) / (14.0 * (-3 + n1) * (-2 + n1) * (-1 + n1) * n1 * np.pi)


def mpc(l, pn):
    """Spherical-Bessel power-law transform coefficient (pybird.py:159-173)."""
    return np.pi**-1.5 * 2.0 ** (-2.0 * pn) * np.exp(
        loggamma(1.5 + l / 2.0 - pn) - loggamma(l / 2.0 + pn)
    )


def kgrid(kmax=0.3):
    """pybird.py:472-479"""
    if kmax > 0.30:
        low = np.array([0.001, 0.005, 0.0075, 0.01, 0.0125, 0.015, 0.0175, 0.02])
        ext = np.arange(low[-1], kmax + 1e-3, 0.005)
        return np.concatenate([low, ext[1:]])
    return tables()["kbird"].copy()


# mu-power of every loop term (pybird.py:570-582)
MU11 = (0, 2, 4)
MUCT = (0, 2, 4, 2, 4, 6)
MUNNLO = (4, 6, 8)
MU22 = (0,) * 6 + (2,) * 7 + (4, 2, 4, 2, 4, 2) + (4,) * 3 + (6, 4, 6, 4, 6, 8)
MU13 = (0,) * 2 + (2,) * 4 + (4,) * 3 + (6,)
# (row, f-power, term) of `reducePsCfl` (pybird.py:762-846)
GROUP22 = [(0, 2, 20), (0, 3, 23), (0, 3, 24), (0, 4, 25), (0, 4, 26), (0, 4, 27),
           (1, 1, 9), (1, 2, 14), (1, 2, 15), (1, 3, 21), (1, 3, 22),
           (2, 1, 10), (2, 2, 16), (2, 2, 17),
           (4, 1, 11), (4, 2, 18), (4, 2, 19),
           (5, 0, 0), (5, 1, 6), (5, 2, 12), (5, 2, 13),
           (6, 0, 1), (6, 1, 7), (8, 0, 2), (8, 1, 8), (9, 0, 3), (10, 0, 4), (11, 0, 5)]
GROUP13 = [(0, 2, 7), (0, 3, 8), (0, 3, 9), (1, 1, 3), (1, 2, 5), (1, 2, 6), (3, 1, 4),
           (5, 0, 0), (5, 1, 2), (7, 0, 1)]


@dataclass
class Common:
    """pybird.py:486-582"""

    Nl: int = 2
    No: int | None = None
    kmax: float = 0.3
    optiresum: bool = False
    kmA: float = 0.7
    krA: float = 0.25
    ndA: float = 3e-4
    kmB: float | None = None
    krB: float | None = None
    ndB: float | None = None
    counterform: str = "westcoast"
    with_NNLO: bool = False
    kIR: float | None = None
    IRcutoff: bool | str = False

    def __post_init__(self):
        if self.IRcutoff and self.kIR is None:  # :528-529
            raise ValueError("kIR must be specified when doing IRcutoff")
        if self.IRcutoff is True:  # :530-531
            self.IRcutoff = "all"
        self.No = self.Nl if self.No is None else self.No
        if self.No > self.Nl:
            raise ValueError("No should always be smaller than Nl")
        self.kmB = self.kmA if self.kmB is None else self.kmB
        self.krB = self.krA if self.krB is None else self.krB
        self.ndB = self.ndA if self.ndB is None else self.ndB
        self.k = kgrid(self.kmax)
        self.Nk = self.k.size
        self.s = np.arange(70.0, 200.0, 2.5) if self.optiresum else tables()["sbird"].copy()
        self.Ns = self.s.size
        self.kr = self.k[0.02 <= self.k]
        self.Nkr = self.kr.size
        self.Nklow = self.Nk - self.Nkr
        mu = tables()["mu_to_legendre"]  # [power/2, l/2]
        pick = lambda powers: np.array([[mu[p // 2][l] for p in powers] for l in range(self.Nl)])
        self.l11, self.lct, self.lctNNLO = pick(MU11), pick(MUCT), pick(MUNNLO)
        self.l22, self.l13 = pick(MU22), pick(MU13)


@dataclass
class Bird:
    """The per-evaluation container (pybird.py:635-724); only what the path touches."""

    co: Common
    kin: np.ndarray
    Pin: np.ndarray
    f: float
    DA: float | None = None
    H: float | None = None
    z: float | None = None
    rdrag: float | None = None
    h: float | None = None

    def __post_init__(self):
        self.P11 = interp1d(self.kin, self.Pin, kind="cubic")(self.co.k)  # pybird.py:694-695
        self.Picc = np.zeros((self.co.Nl, self.co.Nk))


# --------------------------------------------------------------------------------------
# NonLinear (pybird.py:870-1171)
# --------------------------------------------------------------------------------------
class NonLinear:
    def __init__(self, co: Common, NFFT=256):
        self.co = co
        self.grid = LogGrid(Nmax=NFFT, xmin=1.5e-5, xmax=1000.0, bias=-1.6)  # :919
        nu = -0.5 * self.grid.Pow
        a, b = nu[:, None], nu[None, :]
        common = m22_common(a, b)
        self.M22 = np.array([common * m22_rational(i, a, b) for i in range(28)])  # :1005-1016
        c13 = m13_common(nu)
        self.M13 = np.array([c13 * m13_rational(i, nu) for i in range(10)])  # :1018-1023
        ell = 2 * np.arange(co.Nl)
        self.Mcf11 = mpc(ell[:, None], nu[None, :])  # :1029
        self.Ml = mpc(ell[:, None, None], a[None] + b[None] - 1.5)  # :1035-1038
        self.Mcf22 = self.Ml[:, None] * self.M22[None]  # :1042 (axis order l,b,n,m)
        self.Mcfct = mpc(ell[:, None], nu - 1.0)  # :1052
        self.McfctNNLO = mpc(ell[:, None], nu - 2.0)  # :1056
        self.kPow = np.exp(np.outer(self.grid.Pow, np.log(co.k)))  # :1060
        self.sPow = np.exp(np.outer(-self.grid.Pow - 3.0, np.log(co.s)))  # :1064

    def coef(self, bird: Bird, window=0.2, IRcut=False):
        """pybird.py:1127-1141: with IRcut the samples below kIR are dropped and the low side is zero padded"""
        k, Pin, extrap = bird.kin, bird.Pin, ("extrap", "extrap")
        if IRcut:
            idx = np.searchsorted(k, self.co.kIR)
            k, Pin, extrap = k[idx:], Pin[idx:], ("padding", "extrap")
        return fftlog_coef(self.grid, k, Pin, extrap=extrap, window=window)

    def PsCf(self, bird: Bird, window=0.2):
        """pybird.py:1143-1171"""
        co = self.co
        mode = co.IRcutoff
        if mode == "all" or mode is False:  # :1151-1160
            c_cf = c_pk = self.coef(bird, window, IRcut=bool(mode))
        elif mode == "loop":
            c_pk, c_cf = self.coef(bird, window, IRcut=True), self.coef(bird, window, IRcut=False)
        elif mode == "resum":
            c_pk, c_cf = self.coef(bird, window, IRcut=False), self.coef(bird, window, IRcut=True)
        else:
            raise ValueError(f"unexpected IRcutoff option: {mode}")
        bird.coef, bird.coef_cf = c_pk, c_cf
        v = c_pk[:, None] * self.kPow  # (N, Nk)
        u = c_cf[:, None] * self.sPow  # (N, Ns)
        c = c_pk
        # P22[b,k] = k^3 Re sum_nm v_nk v_mk M22[b,n,m]   (:1074-1078); contraction order of the reference's
        # einsum path: matrix times panel first (one zgemm), then the dot with the second panel
        N = c.size
        tmp = (self.M22.reshape(28 * N, N) @ v).reshape(28, N, co.Nk)
        bird.P22 = co.k**3 * np.real(np.einsum("nk,bnk->bk", v, tmp))
        bird.P13 = co.k**3 * bird.P11 * np.real(self.M13 @ v)  # :1080-1086
        bird.C11 = np.real(self.Mcf11 @ u)  # :1088-1090
        bird.Cct = co.s**-2 * np.real(self.Mcfct @ u)  # :1092-1096
        if co.with_NNLO:
            bird.CctNNLO = co.s**-4 * np.real(self.McfctNNLO @ u)  # :1098-1101
        # C22[l,b,s] = Re sum_nm u_ns u_ms Ml[l,n,m] M22[b,n,m]   (:1042, :1103-1113)
        t = (self.Mcf22.reshape(co.Nl * 28 * N, N) @ u).reshape(co.Nl, 28, N, co.Ns)
        C22 = np.real(np.einsum("ns,lbns->lbs", u, t))
        # C13[l,b,s] = Re sum_nm u_ns u_ms Ml[l,n,m] M13[b,n]   (:1046, :1115-1125)
        w = (self.Ml.reshape(co.Nl * N, N) @ u).reshape(co.Nl, N, co.Ns)
        C13 = np.real(np.einsum("bn,ns,lns->lbs", self.M13, u, w))
        bird.C22, bird.C13 = C22, C13


def set_PsCfl(bird: Bird):
    """Legendre weighting, f-power grouping, shot-noise subtraction (pybird.py:737-866)."""
    co, f = bird.co, bird.f
    k2P = co.k**2 * bird.P11
    bird.P11l = co.l11[:, :, None] * bird.P11[None, None, :]
    bird.Pctl = co.lct[:, :, None] * k2P[None, None, :]
    bird.PctNNLOl = co.lctNNLO[:, :, None] * (co.k**4 * bird.P11)[None, None, :]
    P22l = co.l22[:, :, None] * bird.P22[None]
    P13l = co.l13[:, :, None] * bird.P13[None]
    C22l = co.l22[:, :, None] * bird.C22
    C13l = co.l13[:, :, None] * bird.C13
    Ploopl = np.zeros((co.Nl, 12, co.Nk))
    Cloopl = np.zeros((co.Nl, 12, co.Ns))
    for row, p, b in GROUP22:
        Ploopl[:, row] += f**p * P22l[:, b]
        Cloopl[:, row] += f**p * C22l[:, b]
    for row, p, b in GROUP13:
        Ploopl[:, row] += f**p * P13l[:, b]
        Cloopl[:, row] += f**p * C13l[:, b]
    Ploopl -= Ploopl[:, :, :1]  # :861-866
    bird.Ploopl, bird.Cloopl = Ploopl, Cloopl
    Pstl = np.zeros((co.Nl, 3, co.Nk))  # :850-859
    Pstl[0, 0] = 1.0
    Pstl[0, 1] = co.k**2
    if co.Nl >= 2:
        Pstl[1, 2] = co.k**2
    bird.Pstl = Pstl


# --------------------------------------------------------------------------------------
# IR resummation (pybird.py:1174-1464): full resummation and the "optiresum" BAO-peak variant
# --------------------------------------------------------------------------------------
class Resum:
    def __init__(self, co: Common, LambdaIR=0.2, NFFT=192):
        self.co = co
        self.LambdaIR = LambdaIR
        if co.optiresum:  # :1235-1244
            self.idlow = np.where(co.s > 70.0)[0][0]
            self.idhigh = np.where(co.s > 190.0)[0][0]
            self.sbao = co.s[self.idlow : self.idhigh]
            self.snobao = np.concatenate([co.s[: self.idlow], co.s[self.idhigh :]])
            self.sr = self.sbao
        else:
            self.sr = co.s
        self.NIR = 16 if co.Nl == 3 else 8  # :1247-1250
        self.Na = 3 if self.NIR == 16 else 2
        self.Nn = 2 * self.NIR * self.Na
        k2 = np.array([co.kr ** (2 * (p + 1)) for p in range(self.NIR)])
        self.k2p = np.concatenate([k2, k2])  # :1261-1262
        self.grid = LogGrid(Nmax=NFFT, xmin=0.1, xmax=10000.0, bias=-0.6)  # :1288
        self.M = np.array([8.0 * np.pi**3 * mpc(2 * l, -0.5 * self.grid.Pow) for l in range(co.Nl)])
        self.kPow = np.exp(np.outer(-self.grid.Pow - 3.0, np.log(co.kr)))  # :1308
        self.xgrid = LogGrid(Nmax=32, xmin=1.5e-5, xmax=10.0, bias=-2.6)  # :1293
        self.XM = np.array([mpc(2 * l, -0.5 * self.xgrid.Pow) for l in range(2)])  # :1310-1314
        self.XsPow = np.exp(np.outer(-self.xgrid.Pow - 3.0, np.log(self.sr)))  # :1304
        qt = tables()["q_nl3" if self.NIR == 16 else "q_nl2"]
        self.qcoef = qt  # [N-j, l, lp, u, degree]

    def filters(self, bird: Bird):
        """IR filters X(s), Y(s) (pybird.py:1316-1353)."""
        kin, Pin, extrap = bird.kin, bird.Pin, "extrap"
        if self.co.IRcutoff in ("all", "resum"):  # :1320-1334
            idx = np.searchsorted(kin, self.co.kIR)
            kin, Pin, extrap = kin[idx:], Pin[idx:], ("padding", "extrap")
        c = fftlog_coef(self.xgrid, kin, Pin * np.exp(-(kin**2) / self.LambdaIR**2) / kin**2,
                        extrap=extrap, window=None)
        X02 = np.real(self.XM @ (c[:, None] * self.XsPow))
        off = np.real(np.sum(c * 1.0 ** (-self.xgrid.Pow - 3.0) * self.XM[0]))
        X02[0] = off - X02[0]
        return 2.0 / 3.0 * (X02[0] - X02[1]), 2.0 * X02[1]

    def Q(self, f):
        """Bulk coefficients Q[a, l, lp, u] (pybird.py:1367-1380): a=0 uses table index 1."""
        fp = f ** np.arange(self.qcoef.shape[-1])
        return (self.qcoef @ fp)[::-1]

    def extract_bao(self, cf):
        """pybird.py:1382-1400: with optiresum, the BAO peak = cf minus a broadband that interpolates s^2 cf
        linearly between the points outside (70, 190]"""
        if not self.co.optiresum:
            return cf
        nobao_in = np.concatenate([cf[..., : self.idlow], cf[..., self.idhigh :]], axis=-1)
        nobao = interp1d(self.snobao, self.snobao**2 * nobao_in, kind="linear", axis=-1)(self.sbao) * self.sbao**-2
        return cf[..., self.idlow : self.idhigh] - nobao

    def _ir(self, XpYp, C):
        """IR[..., u, k] for correlation-function rows C[..., s] (pybird.py:1409-1441)."""
        co = self.co
        C = self.extract_bao(C)
        lead = C.shape[:-1]
        out = np.zeros(lead + (self.Nn, co.Nk))
        prod = XpYp[(None,) * len(lead)] * C[..., None, :]  # (..., j, s)
        coef = fftlog_coef(self.grid, self.sr, prod, extrap="padding", window=None)
        for j in range(2 * self.NIR):
            ir = np.real(np.einsum("vn,...n,nk->...vk", self.M[: self.Na], coef[..., j, :], self.kPow))
            out[..., j * self.Na : (j + 1) * self.Na, co.Nklow :] = self.k2p[j] * ir
        return out

    def Ps(self, bird: Bird):
        """pybird.py:1413-1464"""
        co = self.co
        Q = self.Q(bird.f)
        X, Y = self.filters(bird)
        bird.X, bird.Y = X, Y
        XpYp = np.concatenate([[X ** (p + 1) for p in range(self.NIR)],
                               [Y * X**p for p in range(self.NIR)]])
        IR11 = self._ir(XpYp, bird.C11)
        IRct = self._ir(XpYp, bird.Cct)
        IRloop = self._ir(XpYp, bird.Cloopl)
        bird.P11l = bird.P11l + np.einsum("lpn,pnk,pi->lik", Q[0], IR11, co.l11)
        bird.Pctl = bird.Pctl + np.einsum("lpn,pnk,pi->lik", Q[1], IRct, co.lct)
        bird.Ploopl = bird.Ploopl + np.einsum("lpn,pink->lik", Q[1], IRloop)
        if co.with_NNLO:
            IRn = self._ir(XpYp, bird.CctNNLO)
            bird.PctNNLOl = bird.PctNNLOl + np.einsum("lpn,pnk,pi->lik", Q[1], IRn, co.lctNNLO)


# --------------------------------------------------------------------------------------
# Alcock-Paczynski (pybird.py:1467-1628)
# --------------------------------------------------------------------------------------
def hubble(Om, z):
    return (Om * (1 + z) ** 3.0 + (1 - Om)) ** 0.5  # pybird.py:34-36


def dafunc(Om, z):
    from scipy.integrate import quad

    return quad(lambda x: 1.0 / hubble(Om, x), 0, z)[0] / (1 + z)  # pybird.py:39-42


class APeffect:
    def __init__(self, co: Common, Om_AP=None, z_AP=None, DA=None, H=None, nbinsmu=200,
                 accboost=1, APst=False):
        self.co, self.APst = co, APst
        if DA is not None and H is not None:
            self.DA, self.H = DA, H
        else:
            self.DA, self.H = dafunc(Om_AP, z_AP), hubble(Om_AP, z_AP)
        self.mu = np.linspace(0, 1, accboost * nbinsmu)
        self.kk, self.mm = np.meshgrid(co.k, self.mu, indexing="ij")
        ells = 2 * np.arange(co.Nl)
        self.Lmu = np.array([(2 * l + 1) / 2.0 * eval_legendre(l, self.mm) for l in ells])

    def _integrate(self, P, kp, Lmup):
        spl = interp1d(self.co.k, P, axis=-1, kind="cubic", bounds_error=False, fill_value="extrapolate")
        Pkmu = np.einsum("lpkm,lkm->pkm", spl(kp), Lmup)
        return 2 * np.trapz(np.einsum("pkm,lkm->lpkm", Pkmu, self.Lmu), x=self.mm, axis=-1)

    def AP(self, bird: Bird):
        """pybird.py:1598-1621"""
        qperp, qpar = bird.DA / self.DA, self.H / bird.H
        F = qpar / qperp
        root = 1 + self.mm**2 * (F**-2 - 1)
        kp = self.kk / qperp * root**0.5
        mup = self.mm / F * root**-0.5
        Lmup = np.array([eval_legendre(2 * i, mup) for i in range(self.co.Nl)])
        norm = 1.0 / (qperp**2 * qpar)
        for name in ("P11l", "Pctl", "Ploopl") + (("PctNNLOl",) if self.co.with_NNLO else ()) + (
            ("Pstl",) if self.APst else ()
        ):
            setattr(bird, name, norm * self._integrate(getattr(bird, name), kp, Lmup))


# --------------------------------------------------------------------------------------
# fibre collisions (pybird.py:44-85, :1631-1809), effective-window method
# --------------------------------------------------------------------------------------
def _w2d(x):
    from scipy.special import j1

    return 2.0 * j1(x) / x  # pybird.py:44-46


def _hllp(l, lp, x):
    """pybird.py:49-65"""
    if (l, lp) == (2, 0):
        return x**2 - 1.0
    if (l, lp) == (4, 0):
        return 1.75 * x**4 - 2.5 * x**2 + 0.75
    if (l, lp) == (4, 2):
        return x**4 - x**2
    return x * 0.0


def fiber_dPcorr(co: Common, PS, fs, Dfc, ktrust=0.25):
    """FiberCollision.dPcorr(co.k, co.k, PS) (pybird.py:1703-1756); PS: (Nl, nrow, Nk)."""
    k = co.k
    q = np.geomspace(k.min(), ktrust, num=1024)
    dq = np.concatenate([[0.0], q[1:] - q[:-1]])
    PSq = interp1d(k, PS, axis=-1, bounds_error=False, fill_value="extrapolate")(q)
    out = np.zeros(PS.shape)
    for l in range(co.Nl):
        for lp in range(co.Nl):
            L, Lp = 2 * l, 2 * lp
            for i, kv in enumerate(k):
                if lp <= l:  # IR: q < k  (pybird.py:68-75)
                    m = q < kv
                    x = q[m] / kv
                    f = x * _w2d(q[m] * Dfc) * (x**L if L == Lp else (2.0 * L + 1.0) / 2.0 * _hllp(max(L, Lp), min(L, Lp), x))
                    out[l, :, i] += -0.5 * fs * Dfc**2 * (PSq[lp][:, m] @ (q[m] * dq[m] * f))
                if lp >= l:  # UV: k < q < ktrust  (pybird.py:78-85)
                    m = (q > kv) & (q < ktrust)
                    x = kv / q[m]
                    f = _w2d(q[m] * Dfc) * (x**L if L == Lp else (2.0 * L + 1.0) / 2.0 * _hllp(max(L, Lp), min(L, Lp), x))
                    out[l, :, i] += -0.5 * fs * Dfc**2 * (PSq[lp][:, m] @ (q[m] * dq[m] * f))
    return out


def fibcol_window(bird: Bird, fs, Dfc, ktrust=0.25, fiberst=False):
    """FiberCollision.fibcolWindow (pybird.py:1760-1806)"""
    names = ["P11l", "Pctl", "Ploopl"] + (["PctNNLOl"] if bird.co.with_NNLO else []) + (["Pstl"] if fiberst else [])
    for name in names:
        P = getattr(bird, name)
        setattr(bird, name, P + fiber_dPcorr(bird.co, P, fs, Dfc, ktrust))


# --------------------------------------------------------------------------------------
# window / integral constraint apply step, binning, chained
# --------------------------------------------------------------------------------------
def window_pgrid(kmax=0.3, accboost=1):
    """window.py:27-33"""
    return np.concatenate([np.geomspace(1e-5, 0.015, 100 * accboost, endpoint=False),
                           np.arange(0.015, kmax, 1e-3 / accboost)])


def mask_and_measure(Wal, p, k, windowk, withmask=True):
    """`_compute_Waldk` (window.py:348-359, icc.py:448-459)."""
    W = Wal
    if withmask:
        keep = (p[None, :] < k[:, None] + windowk) & (p[None, :] > k[:, None] - windowk)
        W = Wal * keep[None, None]
    dp = np.concatenate([[0.0], p[1:] - p[:-1]])
    return W * dp


def convolve(Waldk, p, k, P):
    """`integrWindow` (window.py:371-387, icc.py:471-484)."""
    Pp = interp1d(k, P, axis=-1, kind="cubic", bounds_error=False, fill_value="extrapolate")(p)
    return np.einsum("alkp,lsp->ask", Waldk, Pp)


def apply_window(bird: Bird, Waldk, p, window_st=True, icc=None):
    """`Window.Window` (window.py:389-415); icc = (Waldk_ic, PSN*Pshot) or None."""
    k = bird.co.k
    names = ["P11l", "Pctl", "Ploopl"] + (["PctNNLOl"] if bird.co.with_NNLO else []) + (
        ["Pstl"] if window_st else [])
    for name in names:
        P = getattr(bird, name)
        out = convolve(Waldk, p, k, P)
        if icc is not None:
            out = out - convolve(icc[0], p, k, P)
        setattr(bird, name, out)
    if icc is not None:
        bird.Picc = bird.Picc - icc[1]


def compute_Wal(s_Q, co: Common, Na, Nl, Nq=3, pmax=None, accboost=1, Nmax=4096,
                xmin_factor=1.0, xmax_factor=100.0, bias=-1.6, window_param=1):
    """Fourier-space window matrix from the configuration-space Q_l(s)
    (`Window._compute_Wal`, window.py:262-346).  s_Q: array (ns, 1+nq) = s, Q0, Q2, ..."""
    from scipy.special import spherical_jn

    pmax = float(co.k.max()) if pmax is None else pmax
    p = window_pgrid(pmax, accboost)
    tab = np.asarray(s_Q, dtype=float)
    while tab[0, 0] == 0.0:
        tab = tab[1:]
    tab = tab[:, : 1 + Nq]
    # C_{a l q} = (2a+1) (a l q; 0 0 0)^2 -like coupling table (window.py:286-303)
    Calq = np.array([
        [[1, 0, 0, 0], [0, 1 / 5, 0, 0], [0, 0, 1 / 9, 0], [0, 0, 0, 1 / 13]],
        [[0, 1, 0, 0], [1, 2 / 7, 2 / 7, 0], [0, 2 / 7, 100 / 693, 25 / 143], [0, 0, 25 / 143, 14 / 143]],
        [[0, 0, 1, 0], [0, 18 / 35, 20 / 77, 45 / 143], [1, 20 / 77, 162 / 1001, 20 / 143],
         [0, 45 / 143, 20 / 143, 252 / 2431]],
        [[0, 0, 0, 1], [0, 0, 5 / 11, 14 / 55], [0, 5 / 11, 20 / 99, 28 / 187],
         [1, 14 / 55, 28 / 187, 400 / 3553]],
    ])[..., :Nq]
    sw, Qq = tab[:, 0], tab[:, 1:].T
    Qal = np.einsum("alq,qs->als", Calq, Qq)[:Na, :Nl]
    g = LogGrid(Nmax=Nmax, xmin=sw[0] * xmin_factor, xmax=sw[-1] * xmax_factor, bias=bias)
    pPow = np.exp(np.outer(-g.Pow - 3.0, np.log(p)))
    M = np.array([4 * np.pi * mpc(2 * l, -0.5 * g.Pow) for l in range(Nl)])
    a_idx = np.arange(Na)
    l_idx = np.arange(Nl)
    kern = lambda x: spherical_jn(2 * a_idx[:, None, None, None],
                                  x[None, None, None, :] * co.k[None, None, :, None])
    coef = fftlog_coef(g, sw, Qal[:, :, None, :] * np.ones(co.Nk)[None, None, :, None],
                       extrap="padding", window=window_param, kernel=kern)
    coef = ((-1j) ** (2 * a_idx))[:, None, None, None] * ((1j) ** (2 * l_idx))[None, :, None, None] * coef
    Wal = p**2 * np.real(np.einsum("alkn,np,ln->alkp", coef, pPow, M))
    return Wal, p


class Binning:
    """binning.py:17-162 (bins inferred from the last data spacing)."""

    def __init__(self, kout, co: Common, accboost=1, decimals=2):
        from scipy.integrate import quad

        kout = np.asarray(kout, dtype=float)
        self.co = co
        dk = np.round(kout[-1] - kout[-2], decimals)
        centre = (kout[-1] - dk * np.arange(len(kout)))[::-1]
        self.binmin, self.binmax = centre - dk / 2, centre + dk / 2
        self.binvol = np.array([quad(lambda k: k**2, a, b)[0] for a, b in zip(self.binmin, self.binmax)])
        self.keff = np.array([quad(lambda k: k**3, a, b)[0] for a, b in zip(self.binmin, self.binmax)]) / self.binvol
        self.points = np.array([np.linspace(a, b, 100 * accboost) for a, b in zip(self.binmin, self.binmax)])

    def integrate(self, P):
        spl = interp1d(self.co.k, P, axis=-1, kind="cubic", bounds_error=False, fill_value="extrapolate")
        return np.trapz(spl(self.points) * self.points**2, x=self.points, axis=-1) / self.binvol

    def transform(self, terms: dict):
        return {k: self.integrate(v) for k, v in terms.items()}


def chained_matrix(Nl):
    """chained.py:13-54"""
    A = lambda l: ((2 * l + 1) * eval_legendre(l, 0.0)) / ((2 * l + 5) * eval_legendre(l + 2, 0.0))
    m = np.zeros((Nl - 1, Nl))
    for i in range(Nl - 1):
        m[i, i] = 1.0
        m[i, i + 1] = -A(2 * i)
    return m


def chained_transform(terms: dict, Nl):
    m = chained_matrix(Nl)
    return {k: np.einsum("al,l...->a...", m, v) for k, v in terms.items()}


def bird_terms(bird: Bird):
    return dict(P11l=bird.P11l, Ploopl=bird.Ploopl, Pctl=bird.Pctl, Pstl=bird.Pstl, Picc=bird.Picc,
                PctNNLOl=bird.PctNNLOl)


# --------------------------------------------------------------------------------------
# bias reduction (parambasis.py:42-136, :249-316) - west-coast basis
# --------------------------------------------------------------------------------------
def nnlo_vector(co: Common, f, b1A, cnnlo):
    """bctNNLOAB (parambasis.py:96-107): west [b1^2 cr4, b1 cr6, 0] / (4 kr^4); east ctilde [-b1^2 f^4, -2 b1 f^5, -f^6]"""
    if co.counterform == "westcoast":
        cr4, cr6 = cnnlo
        return np.array([0.25 * b1A**2 / co.krA**4 * cr4, 0.25 * b1A / co.krA**4 * cr6, 0.0])
    return cnnlo[0] * np.array([-(b1A**2) * f**4, -2 * b1A * f**5, -(f**6)])


def bias_vectors(co: Common, f, bsA, bsB=None, es=(0.0, 0.0, 0.0)):
    b1A, b2A, b3A, b4A, cctA, cr1A, cr2A = bsA
    b1B, b2B, b3B, b4B, cctB, cr1B, cr2B = bsB if bsB is not None else bsA
    kmA, krA, ndA, kmB, krB, ndB = co.kmA, co.krA, co.ndA, co.kmB, co.krB, co.ndB
    b11 = np.array([b1A * b1B, (b1A + b1B) * f, f**2])
    if co.counterform == "westcoast":
        bct = np.array([
            b1A * cctB / kmB**2 + b1B * cctA / kmA**2,
            b1B * cr1A / krA**2 + b1A * cr1B / krB**2,
            b1B * cr2A / krA**2 + b1A * cr2B / krB**2,
            (cctA / kmA**2 + cctB / kmB**2) * f,
            (cr1A / krA**2 + cr1B / krB**2) * f,
            (cr2A / krA**2 + cr2B / krB**2) * f,
        ])
    else:
        bct = np.array([-cctA - cctB, -(cr1A + cr1B) * f, -(cr2A + cr2B) * f**2, 0.0, 0.0, 0.0])
    bloop = np.array([
        1.0, 0.5 * (b1A + b1B), 0.5 * (b2A + b2B), 0.5 * (b3A + b3B), 0.5 * (b4A + b4B), b1A * b1B,
        0.5 * (b1A * b2B + b1B * b2A), 0.5 * (b1A * b3B + b1B * b3A), 0.5 * (b1A * b4B + b1B * b4A),
        b2A * b2B, 0.5 * (b2A * b4B + b2B * b4A), b4A * b4B,
    ])
    x1 = 0.5 * (1.0 / ndA + 1.0 / ndB)
    x2 = 0.5 * (1.0 / ndA / kmA**2 + 1.0 / ndB / kmB**2)
    ce0, cemono, cequad = es
    bst = np.array([ce0 * x1, cemono * x2, cequad * x2])
    return b11, bct, bloop, bst


def reduce_Plk(co: Common, f, terms: dict, bsA, bsB=None, es=(0.0, 0.0, 0.0), cnnlo=None):
    """Full multipoles P_l(k) = sum_b bias_b * term_b + Picc (parambasis.py:129-136 + `.sum()`).
    cnnlo: (cr4, cr6) west / (ctilde,) east when co.with_NNLO (parambasis.py:96-107, :132-134)."""
    b11, bct, bloop, bst = bias_vectors(co, f, bsA, bsB, es)
    No = min(co.No, terms["P11l"].shape[0])
    out = np.einsum("b,lbx->lx", b11, terms["P11l"][:No])
    out = out + np.einsum("b,lbx->lx", bloop, terms["Ploopl"][:No])
    pct = np.einsum("b,lbx->lx", bct, terms["Pctl"][:No])
    if co.with_NNLO and cnnlo is not None:
        pct = pct + np.einsum("b,lbx->lx", nnlo_vector(co, f, bsA[0], cnnlo), terms["PctNNLOl"][:No])
    out = out + pct
    out = out + np.einsum("b,lbx->lx", bst, terms["Pstl"][:No])
    return out + terms["Picc"][:No]


def east_to_west(f, b1, b2, bG2, bGamma3, c0, c2, c4, Pshot, a0, a2):
    """EastCoastBasis.reduce_Plk parameter map (parambasis.py:379-400): returns (bsA, es)"""
    bsA = [b1, b1 + 7 / 2 * bG2, b1 + 15 * bG2 + 6 * bGamma3, 1 / 2 * b2 - 7 / 2 * bG2,
           c0 - f / 3 * c2 + 3 / 35 * f**2 * c4, c2 - 6 / 7 * f * c4, c4]
    return bsA, [Pshot, a0 + 1 / 3 * a2, 2 / 3 * a2]


def gaussian_table_east(co: Common, f, terms: dict, b1):
    """EastCoastBasis.reduce_Plk_gaussian_table (parambasis.py:403-454)"""
    No = min(co.No, terms["P11l"].shape[0])
    L, C, S = terms["Ploopl"][:No], terms["Pctl"][:No], terms["Pstl"][:No]
    x1 = 0.5 * (1.0 / co.ndA + 1.0 / co.ndB)
    x2 = 0.5 * (1.0 / co.ndA / co.kmA**2 + 1.0 / co.ndB / co.kmB**2)
    out = {"bGamma3": 6.0 * (L[:, 3] + b1 * L[:, 7]), "c0": -2.0 * C[:, 0], "c2": 2 / 3 * f * C[:, 0] - 2.0 * f * C[:, 1],
           "c4": -6 / 35 * f**2 * C[:, 0] + 12 / 7 * f**2 * C[:, 1] - 2.0 * f**2 * C[:, 2]}
    if co.with_NNLO:
        N = terms["PctNNLOl"][:No]
        out["ctilde"] = -(b1**2) * f**4 * N[:, 0] - 2.0 * b1 * f**5 * N[:, 1] - f**6 * N[:, 2]
    out["Pshot"], out["a0"], out["a2"] = x1 * S[:, 0], x2 * S[:, 1], x2 / 3 * (S[:, 1] + 2.0 * S[:, 2])
    return out


def gaussian_table_west(co: Common, f, terms: dict, b1A, b1B=None, cross=False):
    """dP/d(gaussian parameter) rows (parambasis.py:249-316).  Returns dict name -> (No, nk);
    auto: b3,cct,cr1,cr2,ce0,cemono,cequad; cross: A_b3..A_cr2, B_b3..B_cr2, ce0,cemono,cequad."""
    No = min(co.No, terms["P11l"].shape[0])
    L, C, S = terms["Ploopl"][:No], terms["Pctl"][:No], terms["Pstl"][:No]
    kmA, krA, ndA, kmB, krB, ndB = co.kmA, co.krA, co.ndA, co.kmB, co.krB, co.ndB
    out = {}
    if cross:
        out["A_b3"] = 0.5 * L[:, 3] + 0.5 * b1B * L[:, 7]
        out["A_cct"] = b1B / kmA**2 * C[:, 0] + f / kmA**2 * C[:, 3]
        out["A_cr1"] = b1B / krA**2 * C[:, 1] + f / krA**2 * C[:, 4]
        out["A_cr2"] = b1B / krA**2 * C[:, 2] + f / krA**2 * C[:, 5]
        out["B_b3"] = 0.5 * L[:, 3] + 0.5 * b1A * L[:, 7]
        out["B_cct"] = b1A / kmB**2 * C[:, 0] + f / kmB**2 * C[:, 3]
        out["B_cr1"] = b1A / krB**2 * C[:, 1] + f / krB**2 * C[:, 4]
        out["B_cr2"] = b1A / krB**2 * C[:, 2] + f / krB**2 * C[:, 5]
    else:
        out["b3"] = L[:, 3] + b1A * L[:, 7]
        out["cct"] = 2.0 * b1A / kmA**2 * C[:, 0] + 2.0 * f / kmA**2 * C[:, 3]
        out["cr1"] = 2.0 * b1A / krA**2 * C[:, 1] + 2.0 * f / krA**2 * C[:, 4]
        out["cr2"] = 2.0 * b1A / krA**2 * C[:, 2] + 2.0 * f / krA**2 * C[:, 5]
    if co.with_NNLO and not cross:  # parambasis.py:303-307
        N = terms["PctNNLOl"][:No]
        out["cr4"], out["cr6"] = 0.25 * b1A**2 / krA**4 * N[:, 0], 0.25 * b1A / krA**4 * N[:, 1]
    x1 = 0.5 * (1.0 / ndA + 1.0 / ndB)
    x2 = 0.5 * (1.0 / ndA / kmA**2 + 1.0 / ndB / kmB**2)
    out["ce0"], out["cemono"], out["cequad"] = S[:, 0] * x1, S[:, 1] * x2, S[:, 2] * x2
    return out


# --------------------------------------------------------------------------------------
# analytic marginalisation (marginal.py:79-196)
# --------------------------------------------------------------------------------------
def marginalized_logp(PNG, PG, data, invcov, mu_G=None, sigma_inv=None, jeffreys=False,
                      return_bestfit=False):
    nG = PG.shape[0]
    mu_G = np.zeros(nG) if mu_G is None else mu_G
    sigma_inv = np.zeros((nG, nG)) if sigma_inv is None else sigma_inv
    res = PNG - data
    F2 = np.einsum("ia,ab,jb->ij", PG, invcov, PG) + sigma_inv  # :167-175
    F1 = -np.einsum("ia,ab,b->i", PG, invcov, res) + sigma_inv @ mu_G  # :177-185
    F0 = res @ invcov @ res + mu_G @ sigma_inv @ mu_G  # :187-196
    sign, logdet = np.linalg.slogdet(F2 / (2 * np.pi))
    if sign <= 0:
        raise RuntimeError("det of F2ij <= 0")  # :113-116
    best = np.linalg.solve(F2, F1)
    chi2 = -F1 @ best + F0 + (0.0 if jeffreys else logdet)  # :118-122
    if not return_bestfit:
        return -0.5 * chi2
    r = best @ PG + PNG - data
    return -0.5 * chi2, r @ invcov @ r, best


def eval_callable(s, env):
    """marginal.py:13-20"""
    fn = eval(s, env)
    return fn(*(env[p] for p in inspect.getfullargspec(fn).args))


def prior_mu_sigma_inv(valid_prior, env):
    """marginal.py:60-77 `mu_G`, `sigma_inv` for one point: loc / scale may be strings eval'ed against `env`"""
    loc = [eval_callable(d["loc"], env) if isinstance(d["loc"], str) else d["loc"] for d in valid_prior.values()]
    std = [eval_callable(d["scale"], env) if isinstance(d["scale"], str) else d["scale"] for d in valid_prior.values()]
    n = len(std)
    if np.inf in std:
        return np.array(loc, dtype=np.float64), np.zeros((n, n))  # :74-75
    return np.array(loc, dtype=np.float64), np.diag(1 / np.array(std, dtype=np.float64) ** 2)


def hartlap(Nreal, ndata):
    return (Nreal - ndata - 2) / (Nreal - 1)  # likelihood.py:163-164


# --------------------------------------------------------------------------------------
# un-binned likelihood products (theory.py:75-106, likelihood.py:503-547)
# --------------------------------------------------------------------------------------
def plk_interpolator(kgrid, Plk):
    """theory.py:75-106 `PlkInterpolator`: cubic interpolation of k P_l(k) through the nodes plus an inserted
    (k, kP) = (0, 0) point, extrapolating; returns fn(k) -> P_l(k) of shape Plk.shape[:-1] + k.shape."""
    kg = np.hstack(([0.0], kgrid))
    P = np.insert(np.asarray(Plk, float), 0, 0.0, axis=-1)
    tmp = interp1d(kg, kg * P, axis=-1, kind="cubic", bounds_error=False, fill_value="extrapolate")
    return lambda k: tmp(k) / k


def gaussian_row_interp(kgrid, plk, kout):
    """likelihood.py:510-513: the marginalised-parameter rows are interpolated WITHOUT the inserted origin"""
    return interp1d(kgrid, kgrid * np.asarray(plk, float), kind="cubic", axis=-1)(kout) / kout
