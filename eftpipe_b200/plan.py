"""Plan construction: every cosmology-independent operator of the hot path, built once on the
host (numpy/scipy, fp64 with extended precision where phases are involved) and uploaded to
the GPU.  Nothing here runs per evaluation.

Design: for fixed grids almost every stage of the reference pipeline is a
fixed linear map, and the one-loop contraction  sum_{nm} c_n c_m M[n,m] x^{eta_n+eta_m}
depends on x only through n+m.  The plan therefore holds

  front      real matrix  [P_lin samples | power-law tails] -> FFTLog coefficients c_n and every
             quantity linear in c (P11, the 13-loop in k-space, C11, Cct, IR filters X, Y)
  antidiag   pair table  M[t][p][ch] = sym. kernel values along the anti-diagonal n+m=t, so that
             D_ch[t] = sum_{n+m=t} c_n c_m M_ch[n,m]  (28 22-type + 10 13-type channels)
  spectral   real matrices taking D[t] to P22(k), C22_l(s), C13_l(s)
  resum      real matrix R[v,k,s] (spline + FFTLog-192 + Bessel back-transform, all linear)
             and the Q^{ll'}(f) polynomial table
  ap         not-a-knot B-spline collocation inverse, per-interval basis polynomials, the
             mu quadrature with Legendre weights
  project    window (+ integral constraint), binning and chained-multipole mixing composed into
             one real matrix acting on the Nl*Nk internal nodes

All per-evaluation arrays on the device are "batch-minor": the cosmology index is the fastest
axis, so every kernel is coalesced with one lane per cosmology and every fixed operator is a
GEMM  C[M, B] = A[M, K] X[K, B].
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy.interpolate import CubicSpline, PPoly, make_interp_spline

from . import tables
from .fftlog import FFTLog

N22, N13, NCH = 28, 10, 38
N11, NCT, NLOOP, NST, NNNLO = 3, 6, 12, 3, 3
LAMBDA_IR = 0.2


def cubic_matrix(xk, xq):
    """Matrix of `interp1d(xk, ., kind="cubic", fill_value="extrapolate")(xq)`: not-a-knot cubic
    spline, end polynomials continued outside (pybird.py:1586-1593, window.py:376-383,
    binning.py:135-142).  Shape (len(xq), len(xk))."""
    return CubicSpline(np.asarray(xk, float), np.eye(len(xk)), axis=0, extrapolate=True)(np.asarray(xq, float))


# --------------------------------------------------------------------------------------------
@dataclass
class GridConfig:
    """Static sizes and grids, the numeric content of `pybird.Common` (pybird.py:486-582)."""

    Nl: int = 2
    kmax: float = 0.3
    NFFT: int = 256
    with_NNLO: bool = False
    optiresum: bool = False
    k: np.ndarray = field(init=False)
    s: np.ndarray = field(init=False)

    def __post_init__(self):
        self.k = tables.kbird(self.kmax)
        # `s`: the grid the reference evaluates correlation functions on (co.s); `sr`: the grid the resummation
        # integrates over; `E` (len(sr), len(s)): the linear map C(s) -> resummation input.  Full resummation: sr = s,
        # E = identity (None).  optiresum (pybird.py:553-554, :1235-1244, :1382-1400): s = arange(70, 200, 2.5),
        # sr = the BAO range (70, 190], E = restriction minus the broadband that interpolates s^2 C linearly between
        # the points outside that range.  The device only ever holds E.C, so its `Ns` is len(sr).
        if self.optiresum:
            self.s = np.arange(70.0, 200.0, 2.5)
            idlow, idhigh = np.where(self.s > 70.0)[0][0], np.where(self.s > 190.0)[0][0]
            self.sr = self.s[idlow:idhigh]
            out = np.concatenate([np.arange(idlow), np.arange(idhigh, self.s.size)])
            snobao = self.s[out]
            from scipy.interpolate import interp1d

            broadband = interp1d(snobao, np.eye(snobao.size), kind="linear", axis=0)(self.sr)  # (Nsr, Nout)
            E = np.zeros((self.sr.size, self.s.size))
            E[np.arange(self.sr.size), idlow + np.arange(self.sr.size)] = 1.0
            E[:, out] -= broadband * snobao[None, :] ** 2 * self.sr[:, None] ** -2
            self.E = E
        else:
            self.s = tables.sbird()
            self.sr, self.E = self.s, None
        self.Nk, self.Ns, self.Ns_full = self.k.size, self.sr.size, self.s.size
        self.kr = self.k[0.02 <= self.k]
        self.Nkr = self.kr.size
        self.Nklow = self.Nk - self.Nkr
        self.l11 = tables.legendre_weights(self.Nl, tables.MU11)
        self.lct = tables.legendre_weights(self.Nl, tables.MUCT)
        self.lctNNLO = tables.legendre_weights(self.Nl, tables.MUNNLO)
        self.l22 = tables.legendre_weights(self.Nl, tables.MU22)
        self.l13 = tables.legendre_weights(self.Nl, tables.MU13)
        self.nterm = N11 + NCT + NLOOP + NST + (NNNLO if self.with_NNLO else 0)

    def to_sr(self, mat):
        """apply E along axis -2 of an operator whose rows are indexed by s: (..., Ns_full, K) -> (..., Ns, K)"""
        return mat if self.E is None else np.einsum("rs,...sk->...rk", self.E, mat)


# --------------------------------------------------------------------------------------------
# loop kernels
# --------------------------------------------------------------------------------------------
def loop_matrices(fft: FFTLog):
    """M22 (28,N,N) and M13 (10,N) exactly as the reference defines them (pybird.py:1005-1023)."""
    nu = -0.5 * fft.Pow
    M22 = tables.loop22_gamma(nu)[None] * tables.loop22_rational(nu)
    M13 = tables.loop13_prefactor(nu)[None] * tables.loop13_rational(nu)
    return M22, M13


def antidiagonal_table(M22, M13):
    """Pair table for t = n+m <= Nmax, n <= m.  For a symmetric kernel
         D[t] = sum_{n+m=t} c_n c_m M[n,m] = sum_{n<m} c_n c_m (M[n,m]+M[m,n]) + c_n^2 M[n,n];
    the 13-type channels carry M13[b,n] on the first index only (pybird.py:1046), which the
    same symmetrisation handles.  Returns (table[npair, NCH] complex, offsets[Nmax+2] int32);
    pair p of anti-diagonal t is (n, m) = (p, t-p), p = 0..t//2."""
    N = M22.shape[-1]
    Nmax = N - 1
    offsets = np.zeros(Nmax + 2, dtype=np.int32)
    rows = []
    for t in range(Nmax + 1):
        n = np.arange(t // 2 + 1)
        m = t - n
        sym = np.where(n == m, 1.0, 2.0)
        # use the exact average of the (numerically 5e-15-symmetric) reference matrix
        v22 = 0.5 * sym[None, :] * (M22[:, n, m] + M22[:, m, n])
        v13 = np.where(n == m, M13[:, n], M13[:, n] + M13[:, m])
        rows.append(np.concatenate([v22, v13], axis=0).T)
        offsets[t + 1] = offsets[t] + n.size
    return np.ascontiguousarray(np.concatenate(rows, axis=0)), offsets


def _phase_ld(eta, logx):
    """exp(eta * log x) with the (large) phase accumulated in extended precision."""
    e = np.asarray(eta).astype(np.clongdouble)
    lx = np.asarray(logx).astype(np.longdouble)
    return np.exp(np.multiply.outer(e, lx))


def spectral_matrices(fft: FFTLog, g: GridConfig):
    """Real matrices taking the Hermitian half of D[t] (t = 0..Nmax; K index = 2t + {0: Re, 1: Im})
    to  P22[b,k] = k^3 Re sum_t D_b[t] k^{eta2_t}  (pybird.py:1074-1078) and
        C22[l,b,s] = Re sum_t Ml[l,t] D_b[t] s^{-eta2_t-6}  (pybird.py:1042, :1103-1113),
    eta2_t = eta_n + eta_m for n+m=t, Ml[l,t] = MPC(2l, nu_n+nu_m-3/2) (pybird.py:1035-1038)."""
    Nmax = fft.Nmax
    t = np.arange(Nmax + 1)
    delta = 2.0 * np.pi / (Nmax * fft.dx)
    eta2 = 2.0 * fft.bias + 1j * delta * (t - Nmax)
    herm = np.where(t == Nmax, 1.0, 2.0)
    Ek = _phase_ld(eta2, np.log(g.k.astype(np.longdouble))) * (g.k.astype(np.longdouble) ** 3)[None, :]
    Ak = np.empty((g.Nk, 2 * (Nmax + 1)))
    Ak[:, 0::2] = (herm[:, None] * Ek.real).T.astype(float)
    Ak[:, 1::2] = (-herm[:, None] * Ek.imag).T.astype(float)
    Ml = tables.bessel_power(2 * np.arange(g.Nl)[:, None], (-0.5 * eta2 - 1.5)[None, :])
    Es = _phase_ld(-eta2 - 6.0, np.log(g.s.astype(np.longdouble)))
    As = np.empty((g.Nl, g.Ns_full, 2 * (Nmax + 1)))
    for l in range(g.Nl):
        G = Ml[l].astype(np.clongdouble)[:, None] * Es
        As[l][:, 0::2] = (herm[:, None] * G.real).T.astype(float)
        As[l][:, 1::2] = (-herm[:, None] * G.imag).T.astype(float)
    return Ak, np.ascontiguousarray(g.to_sr(As))


# --------------------------------------------------------------------------------------------
# front end: everything linear in [P_lin | tails]
# --------------------------------------------------------------------------------------------
@dataclass
class FrontLayout:
    nin: int
    ntail: int
    ntailx: int
    rows: dict  # name -> (start, count)

    @property
    def K(self):
        return self.nin + self.ntail + self.ntailx

    @property
    def M(self):
        return max(a + b for a, b in self.rows.values())


def _coef_operator(fft: FFTLog, kin, window, cut_index=0):
    """(Nmax+1, nin) complex matrix of `FFTLog.Coef` on the samples kin[cut_index:] with a zero-padded low side
    (NonLinear.Coef with IRcut, pybird.py:1137-1141); cut_index = 0 is the default two-sided "extrap" call, whose
    low tail never triggers because kin[0] lies below the FFTLog grid (checked)."""
    if cut_index == 0:
        if fft.has_low_tail(kin):
            raise NotImplementedError("input k-grid must start below the FFTLog xmin (reference default)")
        return fft.operator(kin, window=window)
    L = np.zeros((fft.Nmax + 1, kin.size), dtype=complex)
    L[:, cut_index:] = fft.operator(kin[cut_index:], window=window)
    return L


def front_operator(kin, g: GridConfig, fft: FFTLog, M13, window=0.2, lambda_ir=LAMBDA_IR, ircutoff=False, kIR=None):
    """Real matrix Wf (M, K) and layout.  Input vector u = [P_lin(kin) | tail | tailX] with
      tail_i  = P_last exp(n lr_i),  n  from the last two samples of P_lin           (fftlog.py:146-151)
      tailX_i = fX_last exp(nX lrx_i), fX = P_lin exp(-k^2/Lambda^2)/k^2            (pybird.py:1321-1325)
    Output rows: Re c_n, Im c_n (n = 0..Nmax/2), P11(k), Re sum_n c_n k^{eta_n} M13[b,n], C11, Cct
    [, CctNNLO], X(s), Y(s).

    ircutoff in {False, "all", "loop", "resum"} (pybird.py:1151-1160, :1320-1334): the k-space loops, the
    configuration-space terms and the IR filters each use either the full P_lin or the samples at k >= kIR with a
    zero-padded low side.  When the two coefficient sets differ, the configuration-space one is emitted as extra
    rows "cre_cf"/"cim_cf" and the anti-diagonal stage runs once per set."""
    kin = np.asarray(kin, float)
    if ircutoff is True:
        ircutoff = "all"
    if ircutoff not in (False, "all", "loop", "resum"):
        raise ValueError(f"unexpected IRcutoff option: {ircutoff}")
    if ircutoff and kIR is None:
        raise ValueError("kIR must be specified when doing IRcutoff")
    icut = int(np.searchsorted(kin, kIR)) if ircutoff else 0
    cut_pk = icut if ircutoff in ("all", "loop") else 0
    cut_cf = icut if ircutoff in ("all", "resum") else 0
    cut_x = icut if ircutoff in ("all", "resum") else 0
    Nmax, Nh = fft.Nmax, fft.Nmax // 2
    Lt, lr = fft.tail_operator(kin, window=window)
    Lc = np.concatenate([_coef_operator(fft, kin, window, cut_pk), Lt], axis=1)  # c = Lc @ [P | tail]
    Lc_cf = Lc if cut_cf == cut_pk else np.concatenate([_coef_operator(fft, kin, window, cut_cf), Lt], axis=1)
    nin, ntail = kin.size, lr.size

    xf = FFTLog(Nmax=32, xmin=1.5e-5, xmax=10.0, bias=-2.6)  # pybird.py:1293
    wX = np.exp(-(kin**2) / lambda_ir**2) / kin**2
    LX = _coef_operator(xf, kin, None, cut_x) * wX[None, :]
    LXt, lrx = xf.tail_operator(kin, window=None)
    ntailx = lrx.size
    K = nin + ntail + ntailx

    def embed(mat_nl=None, mat_x=None):
        nrow = (mat_nl if mat_nl is not None else mat_x).shape[0]
        out = np.zeros((nrow, K), dtype=complex)
        if mat_nl is not None:
            out[:, : nin + ntail] = mat_nl
        if mat_x is not None:
            out[:, :nin] += mat_x[:, :nin]
            out[:, nin + ntail :] = mat_x[:, nin:]
        return out

    blocks, rows, cursor = [], {}, 0

    def add(name, mat):
        nonlocal cursor
        mat = np.ascontiguousarray(mat.reshape(-1, K))
        blocks.append(mat)
        rows[name] = (cursor, mat.shape[0])
        cursor += mat.shape[0]

    Cfull = embed(Lc)
    Cfull_cf = Cfull if Lc_cf is Lc else embed(Lc_cf)
    add("cre", Cfull[: Nh + 1].real)
    add("cim", Cfull[: Nh + 1].imag)
    add("P11", embed(np.concatenate([cubic_matrix(kin, g.k), np.zeros((g.Nk, ntail))], axis=1)).real)
    kPow = np.exp(np.outer(fft.Pow, np.log(g.k)))  # pybird.py:1060
    sPow = np.exp(np.outer(-fft.Pow - 3.0, np.log(g.s)))  # pybird.py:1064
    nu = -0.5 * fft.Pow
    ell = 2 * np.arange(g.Nl)
    # P13raw[b,k] = Re sum_n M13[b,n] kPow[n,k] c_n      (pybird.py:1080-1086 without k^3 P11)
    G13 = (M13[:, None, :] * kPow.T[None, :, :]).reshape(N13 * g.Nk, -1)
    add("P13raw", (G13 @ Cfull).real)
    Mcf11 = tables.bessel_power(ell[:, None], nu[None, :])  # pybird.py:1029
    # configuration-space rows live on the resummation grid: E applied to the reference's C(s) (identity unless optiresum)
    on_sr = lambda G: g.to_sr((G.reshape(g.Nl * g.Ns_full, -1) @ Cfull_cf).real.reshape(g.Nl, g.Ns_full, -1))
    add("C11", on_sr(Mcf11[:, None, :] * sPow.T[None]))
    Mcfct = tables.bessel_power(ell[:, None], nu[None, :] - 1.0)  # pybird.py:1052
    add("Cct", on_sr((Mcfct[:, None, :] * sPow.T[None]) * (g.s**-2)[None, :, None]))  # pybird.py:1092-1096
    if g.with_NNLO:
        Mn = tables.bessel_power(ell[:, None], nu[None, :] - 2.0)  # pybird.py:1056
        add("CctNNLO", on_sr((Mn[:, None, :] * sPow.T[None]) * (g.s**-4)[None, :, None]))
    # IR filters (pybird.py:1316-1353) on the resummation grid: X = 2/3 (X0off - X0 - X2), Y = 2 X2
    CX = embed(mat_x=np.concatenate([LX, LXt], axis=1))
    XM = np.array([tables.bessel_power(2 * l, -0.5 * xf.Pow) for l in range(2)])
    XsPow = np.exp(np.outer(-xf.Pow - 3.0, np.log(g.sr)))
    X02 = (XM[:, None, :] * XsPow.T[None]) @ CX  # (2, Ns, K) complex
    off = (XM[0] @ CX)[None, :]
    X0 = off - X02[0]
    add("X", (2.0 / 3.0 * (X0 - X02[1])).real)
    add("Y", (2.0 * X02[1]).real)
    if Cfull_cf is not Cfull:
        add("cre_cf", Cfull_cf[: Nh + 1].real)
        add("cim_cf", Cfull_cf[: Nh + 1].imag)
    Wf = np.concatenate(blocks, axis=0)
    layout = FrontLayout(nin=nin, ntail=ntail, ntailx=ntailx, rows=rows)
    aux = dict(lr=lr, lrx=lrx, wX_last=wX[-1], wX_prev=wX[-2],
               inv_dlog=1.0 / (np.log(kin[-1]) - np.log(kin[-2])))
    return Wf, layout, aux


# --------------------------------------------------------------------------------------------
# IR resummation
# --------------------------------------------------------------------------------------------
def resum_operator(g: GridConfig, NFFT=192):
    """R[v, k, s]: the linear chain  C(s) -> cubic spline on the FFTLog-192 grid (zero padded
    outside [s_0, s_-1]) -> coefficients -> Bessel transform to k_r (pybird.py:1288-1308,
    :1355-1365, :1409-1411).  IR[v,k] = sum_s R[v,k,s] (XpYp_j * C)(s)."""
    Nl = g.Nl
    NIR = 16 if Nl == 3 else 8  # pybird.py:1247-1258
    Na = 3 if NIR == 16 else 2
    fft = FFTLog(Nmax=NFFT, xmin=0.1, xmax=10000.0, bias=-0.6)
    L = fft.operator(g.sr, window=None)  # (N, Ns)
    M = np.array([8.0 * np.pi**3 * tables.bessel_power(2 * l, -0.5 * fft.Pow) for l in range(Na)])
    kPow = np.exp(np.outer(-fft.Pow - 3.0, np.log(g.kr)))
    R = np.real(np.einsum("vn,nk,ns->vks", M, kPow, L))
    return dict(R=np.ascontiguousarray(R), NIR=NIR, Na=Na, q=tables.resum_coefficients(Nl),
                kr2=g.kr**2)


def resum_slot_count(rs):
    """Number of (l', kind, v) polynomials of Q^{ll'} that do not vanish identically - the ones the resum
    kernel evaluates (csrc/resum.cu resum_pack applies the same rule)."""
    q, NIR, Na = rs["q"], rs["NIR"], rs["Na"]
    Nl = q.shape[1]
    nz = np.any(q.reshape(2, Nl, Nl, 2, NIR, Na, -1) != 0, axis=(0, 1, 4, 6))  # (l', kind, v)
    return int(nz.sum())


# --------------------------------------------------------------------------------------------
# Alcock-Paczynski
# --------------------------------------------------------------------------------------------
def ap_operator(g: GridConfig, nbinsmu=200, accboost=1):
    """Constants of the AP resampling (pybird.py:1538-1548, :1581-1596).

    The reference spline `interp1d(kind="cubic")` is the not-a-knot B-spline interpolant.  For
    fixed nodes its coefficient vector is `Cinv @ values`; `basis[j, r, d]` are the monomial
    coefficients, in (x - knot_lo[j]), of the 4 B-splines alive on interval j, so that
        spline(x) = sum_r coef[j + r] * sum_d basis[j, r, d] (x - knot_lo[j])^d ,
    with the first/last interval continued for extrapolation (fill_value="extrapolate")."""
    k = g.k
    n = k.size
    spl = make_interp_spline(k, np.eye(n), k=3, axis=0)
    tk = spl.t  # knots: x0*4, x2..x_{n-3}, x_{n-1}*4
    Cinv = np.ascontiguousarray(spl.c)  # (n coef, n values)
    nint = n - 3
    lo = tk[3 : 3 + nint]
    basis = np.zeros((nint, 4, 4))
    from scipy.interpolate import BSpline

    for r_abs in range(n):
        e = np.zeros(n)
        e[r_abs] = 1.0
        pp = PPoly.from_spline(BSpline(tk, e, 3, extrapolate=True))
        # pp.c[d', i] on breakpoints pp.x (with repeated end knots); map to our intervals
        for j in range(nint):
            r = r_abs - j
            if 0 <= r <= 3:
                i = j + 3  # knot interval index in pp
                basis[j, r, :] = pp.c[::-1, i]
    mu = np.linspace(0, 1, nbinsmu * accboost)
    # np.trapz weights on the actual linspace values
    d = np.diff(mu)
    wt = np.zeros_like(mu)
    wt[:-1] += 0.5 * d
    wt[1:] += 0.5 * d
    from scipy.special import eval_legendre

    # 2 * trapz( (2l+1)/2 L_l(mu) * ... )   (pybird.py:1546-1548, :1595-1596)
    wl = np.array([2.0 * wt * (2 * l + 1) / 2.0 * eval_legendre(l, mu) for l in 2 * np.arange(g.Nl)])
    return dict(Cinv=Cinv, knot_lo=np.ascontiguousarray(lo), basis=np.ascontiguousarray(basis), mu=mu,
                wl=np.ascontiguousarray(wl), nint=nint)


# --------------------------------------------------------------------------------------------
# window / ICC / binning / chained -> one projection matrix
# --------------------------------------------------------------------------------------------
def window_pgrid(kmax=0.3, accboost=1):
    """window.py:27-33"""
    return np.concatenate([np.geomspace(1e-5, 0.015, 100 * accboost, endpoint=False),
                           np.arange(0.015, kmax, 1e-3 / accboost)])


def window_effective_matrix(Wal, p, k, windowk=0.05, withmask=True):
    """(Na, Nk, Nl, Nk) operator equal to mask + dp weights + cubic resampling + einsum of
    `Window.integrWindow` (window.py:348-359, :371-387)."""
    W = Wal
    if withmask:
        keep = (p[None, :] < k[:, None] + windowk) & (p[None, :] > k[:, None] - windowk)
        W = Wal * keep[None, None]
    dp = np.concatenate([[0.0], p[1:] - p[:-1]])
    Waldk = W * dp
    S = cubic_matrix(k, p)  # (Np, Nk)
    return np.einsum("alkp,pn->akln", Waldk, S)


def binning_matrix(k, kout, accboost=1, decimals=2, kedges=None):
    """(nbin, Nk) operator of `Binning.integrBinning` and the effective k of each bin
    (binning.py:100-144)."""
    from scipy.integrate import quad

    kout = np.asarray(kout, float)
    if kedges is None:
        dk = np.round(kout[-1] - kout[-2], decimals)
        centre = (kout[-1] - dk * np.arange(len(kout)))[::-1]
        bmin, bmax = centre - dk / 2, centre + dk / 2
    else:
        bmin, bmax = np.asarray(kedges[:-1], float), np.asarray(kedges[1:], float)
    vol = np.array([quad(lambda x: x**2, a, b)[0] for a, b in zip(bmin, bmax)])
    keff = np.array([quad(lambda x: x**3, a, b)[0] for a, b in zip(bmin, bmax)]) / vol
    mat = np.zeros((len(bmin), len(k)))
    for i, (a, b) in enumerate(zip(bmin, bmax)):
        pts = np.linspace(a, b, 100 * accboost)
        d = np.diff(pts)
        wt = np.zeros_like(pts)
        wt[:-1] += 0.5 * d
        wt[1:] += 0.5 * d
        mat[i] = (wt * pts**2) @ cubic_matrix(k, pts) / vol[i]
    return mat, keff, bmin, bmax


def interp_matrices(k, kout):
    """Operators of the un-binned ("with_interp") likelihood products on the internal nodes `k`:
      S_png (nkout, Nk): `PlkInterpolator` (theory.py:75-106) - cubic interpolation of k P(k) through the nodes plus an
                         inserted (0, 0) point, divided by kout;
      S_pg  (nkout, Nk): the marginalised rows (likelihood.py:510-513) - the same WITHOUT the inserted origin."""
    k, kout = np.asarray(k, float), np.asarray(kout, float)
    S_png = cubic_matrix(np.hstack(([0.0], k)), kout)[:, 1:] * k[None, :] / kout[:, None]
    S_pg = cubic_matrix(k, kout) * k[None, :] / kout[:, None]
    return S_png, S_pg


def fiber_matrix(k, Nl, fs, Dfc, ktrust=0.25):
    """(Nl, Nk, Nl, Nk) operator F of `FiberCollision.dPcorr` (pybird.py:1703-1756), so that the corrected spectrum
    is P + F.P: linear interpolation of P_l'(k) onto 1024 log-spaced q in [k_0, ktrust] (pybird.py:1714-1721),
    then -fs Dfc^2/2 * sum_q q dq P_l'(q) f_ll'(k, q) with the IR kernel for q < k, l' <= l and the UV kernel for
    k < q < ktrust, l' >= l (pybird.py:44-85, :1728-1755)."""
    from scipy.special import j1

    k = np.asarray(k, float)
    q = np.geomspace(k.min(), ktrust, num=1024)
    dq = np.concatenate([[0.0], q[1:] - q[:-1]])
    # linear interpolation matrix S[q, n] of scipy interp1d(kind="linear", fill_value="extrapolate")
    idx = np.clip(np.searchsorted(k, q, side="right") - 1, 0, k.size - 2)
    w = (q - k[idx]) / (k[idx + 1] - k[idx])
    S = np.zeros((q.size, k.size))
    S[np.arange(q.size), idx] = 1.0 - w
    S[np.arange(q.size), idx + 1] += w
    W2D = 2.0 * j1(q * Dfc) / (q * Dfc)

    def H(l, lp, x):  # pybird.py:49-65
        table = {(2, 0): x**2 - 1.0, (4, 0): 1.75 * x**4 - 2.5 * x**2 + 0.75, (4, 2): x**4 - x**2}
        return table.get((l, lp), 0.0 * x)

    F = np.zeros((Nl, k.size, Nl, k.size))
    for i, kv in enumerate(k):
        ir, uv = q < kv, (q > kv) & (q < ktrust)
        for l in range(Nl):
            for lp in range(Nl):
                L, Lp = 2 * l, 2 * lp
                wq = np.zeros_like(q)
                if lp <= l:
                    x = q[ir] / kv
                    f = x * W2D[ir] * (x**L if L == Lp else (2.0 * L + 1.0) / 2.0 * H(max(L, Lp), min(L, Lp), x))
                    wq[ir] += q[ir] * dq[ir] * f
                if lp >= l:
                    x = kv / q[uv]
                    f = W2D[uv] * (x**L if L == Lp else (2.0 * L + 1.0) / 2.0 * H(max(L, Lp), min(L, Lp), x))
                    wq[uv] += q[uv] * dq[uv] * f
                F[l, i, lp, :] = -0.5 * fs * Dfc**2 * (wq @ S)
    return F


def chained_matrix(Nl):
    """Q_l = P_l - A_l P_{l+2} (chained.py:13-54)."""
    from scipy.special import eval_legendre

    A = lambda l: ((2 * l + 1) * eval_legendre(l, 0.0)) / ((2 * l + 5) * eval_legendre(l + 2, 0.0))
    m = np.zeros((Nl - 1, Nl))
    for i in range(Nl - 1):
        m[i, i] = 1.0
        m[i, i + 1] = -A(2 * i)
    return m


# --------------------------------------------------------------------------------------------
# assembled per-tracer plan (host arrays)
# --------------------------------------------------------------------------------------------
@dataclass
class TracerPlan:
    """Host-side constants of one tracer's pipeline.  `upload()`-ed by `engine.Engine`."""

    grid: GridConfig
    kin: np.ndarray
    Wf: np.ndarray
    front: FrontLayout
    front_aux: dict
    pair_table: np.ndarray  # (npair, NCH) complex128
    pair_offsets: np.ndarray  # (Nmax+2,) int32
    Ak: np.ndarray  # (Nk, 2(Nmax+1))
    As: np.ndarray  # (Nl, Ns, 2(Nmax+1))
    resum: dict | None = None
    ap: dict | None = None
    ap_fid: tuple | None = None  # (DA_fid, H_fid)
    ap_st: bool = False
    project: np.ndarray | None = None  # (Nout, Nl*Nk)
    project_stoch: np.ndarray | None = None  # (Nout, Nl*Nk): operator of the stochastic terms when it differs
    project_st: bool = True  # apply the projection to the stochastic terms
    picc_out: np.ndarray | None = None  # (Nout,)
    out_shape: tuple | None = None  # (Nl_out, nk_out)
    kout: np.ndarray | None = None

    @property
    def Nmax(self):
        return self.grid.NFFT


def build_tracer_plan(Nl=3, kmax=0.3, NFFT=256, with_NNLO=False, kin=None, window=0.2,
                      with_resum=True, resum_NFFT=192, ap=None, projection=None, loop_cache=None,
                      optiresum=False, lambda_ir=LAMBDA_IR, ircutoff=False, kIR=None):
    """ap: None or dict(DA=, H=, nbinsmu=200, accboost=1, APst=False);
    projection: None or dict(matrix=(Nout, Nl*Nk), picc=(Nout,), shape=(Nl_out, nk_out), kout=, st=True);
    optiresum / lambda_ir / ircutoff / kIR: `Common(optiresum=, IRcutoff=, kIR=)`, `Resum(LambdaIR=)`."""
    g = GridConfig(Nl=Nl, kmax=kmax, NFFT=NFFT, with_NNLO=with_NNLO, optiresum=optiresum)
    kin = np.logspace(-5, 0, 200) if kin is None else np.asarray(kin, float)
    fft = FFTLog(Nmax=NFFT, xmin=1.5e-5, xmax=1000.0, bias=-1.6)  # pybird.py:919
    M22, M13 = loop_cache if loop_cache is not None else loop_matrices(fft)
    Wf, layout, aux = front_operator(kin, g, fft, M13, window=window, lambda_ir=lambda_ir, ircutoff=ircutoff, kIR=kIR)
    table, offsets = antidiagonal_table(M22, M13)
    Ak, As = spectral_matrices(fft, g)
    plan = TracerPlan(grid=g, kin=kin, Wf=Wf, front=layout, front_aux=aux, pair_table=table,
                      pair_offsets=offsets, Ak=Ak, As=As)
    if with_resum:
        plan.resum = resum_operator(g, NFFT=resum_NFFT)
    if ap is not None:
        plan.ap = ap_operator(g, nbinsmu=ap.get("nbinsmu", 200), accboost=ap.get("accboost", 1))
        plan.ap_fid = (float(ap["DA"]), float(ap["H"]))
        plan.ap_st = bool(ap.get("APst", False))
    if projection is not None:
        plan.project = np.ascontiguousarray(projection["matrix"], dtype=float)
        plan.picc_out = np.ascontiguousarray(projection.get("picc", np.zeros(plan.project.shape[0])), dtype=float)
        plan.out_shape = tuple(projection["shape"])
        plan.kout = projection.get("kout")
        plan.project_st = bool(projection.get("st", True))
        if projection.get("matrix_st") is not None:
            plan.project_stoch = np.ascontiguousarray(projection["matrix_st"], dtype=float)
    return plan


def stack_projections(blocks):
    """Several projections of one tracer (its (chained, binned) products, theory.py:590-604, un-binned interpolation
    points, window / fibre snapshots) as ONE operator: the blocks' rows one after the other, so that a single projection
    GEMM of the fused pipeline yields every product; `offsets[i] : offsets[i + 1]` are the rows of block i."""
    mats = [np.asarray(b["matrix"], float) for b in blocks]
    out = dict(matrix=np.vstack(mats), picc=np.concatenate([np.asarray(b["picc"], float).reshape(-1) for b in blocks]),
               shape=(1, int(sum(m.shape[0] for m in mats))), st=True, matrix_st=None,
               offsets=np.concatenate([[0], np.cumsum([m.shape[0] for m in mats])]).astype(int))
    if any(b.get("matrix_st") is not None for b in blocks):
        out["matrix_st"] = np.vstack([np.asarray(b["matrix_st"] if b.get("matrix_st") is not None else b["matrix"], float) for b in blocks])
    return out


def compose_projection(g: GridConfig, window=None, icc=None, binning=None, chained=False, window_st=True, fiber=None,
                       fiber_st=False, window_stoch=None, window_picc=None):
    """Compose window (+ICC), fibre collisions, binning and chained mixing into one matrix on the Nl*Nk nodes, in the
    reference's order (theory.py:582-604).

    window : None or (Na, Nk, Nl, Nk) effective matrix (`window_effective_matrix`)
    icc    : None or dict(matrix=(Na,Nk,Nl,Nk), PSN_times_Pshot=(Na,Nk))    (window.py:393-405)
    fiber  : None or (Nl, Nk, Nl, Nk) matrix F of `fiber_matrix`: P <- P + F.P  (pybird.py:1760-1806; not Picc)
    binning: None or (nbin, Nk) matrix (`binning_matrix`)
    window_stoch / window_picc: a probed custom window stage (plugins.probe_linear_stage) brings its own operator for the
             stochastic terms and its own constant on Picc
    Returns dict(matrix=(Nout, Nl*Nk), picc=(Nout,), shape=(Nl_out, nk_out), st=True, matrix_st=None or (Nout, Nl*Nk)).
    `matrix_st` is the operator of the stochastic terms when it differs from `matrix`: the window leaves them alone
    with window_st=False (window.py:401-403), the fibre correction unless fiberst (pybird.py:1798-1806)."""
    Nl, Nk = g.Nl, g.Nk
    eye = np.eye(Nl * Nk).reshape(Nl, Nk, Nl, Nk)
    op, op_st = eye, eye
    picc = np.zeros((Nl, Nk))
    if window is not None:
        op = np.array(window, dtype=float)
        if icc is not None:
            op = op - icc["matrix"]
            picc = picc - icc["PSN_times_Pshot"]
        if window_stoch is not None:
            op_st = np.array(window_stoch, dtype=float)
        elif window_st:
            op_st = op
        elif op.shape != eye.shape:
            raise ValueError("window_st=False needs a window that preserves the node grid")
        if window_picc is not None:
            picc = picc + np.asarray(window_picc, float)
    if fiber is not None:
        if op.shape[0] != Nl:
            raise ValueError("fibre-collision correction needs Na == Nl window output")
        add = eye + np.asarray(fiber, float)
        op = np.einsum("akbm,bmln->akln", add, op)
        if fiber_st:
            op_st = np.einsum("akbm,bmln->akln", add, op_st)
    differs = op_st is not op and not np.array_equal(op_st, op)
    Na = op.shape[0]

    def finish(o):
        if binning is not None:
            o = np.einsum("bk,akln->abln", binning, o)
        if chained:
            o = np.einsum("ca,akln->ckln", chained_matrix(Na), o)
        return o

    op = finish(op)
    if binning is not None:
        picc = picc @ binning.T
    if chained:
        picc = chained_matrix(Na) @ picc
    shape = op.shape[:2]
    out = dict(matrix=op.reshape(shape[0] * shape[1], Nl * Nk), picc=picc.reshape(-1), shape=shape, st=True, matrix_st=None)
    if differs:
        out["matrix_st"] = finish(op_st).reshape(shape[0] * shape[1], Nl * Nk)
    return out
