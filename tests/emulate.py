"""numpy emulation of the CUDA kernels' arithmetic in the device ("batch-minor") layout.

TEST INFRASTRUCTURE.  It lets the CPU test-suite (`-m "not gpu"`) validate every plan operator
built by `eftpipe_b200/plan.py` against the oracle/golden vectors without a GPU, and documents
kernel-by-kernel what the CUDA code computes.  The product never imports this module.
"""
from __future__ import annotations

import numpy as np

from eftpipe_b200 import plan as P
from eftpipe_b200 import tables


def front_inputs(pl, plin):
    """kernel `front_prepare`: u[K, B] from plin[B, nin]."""
    aux, lay = pl.front_aux, pl.front
    B = plin.shape[0]
    u = np.empty((lay.K, B))
    u[: lay.nin] = plin.T
    last, prev = plin[:, -1], plin[:, -2]
    n = (np.log(last) - np.log(prev)) * aux["inv_dlog"]
    u[lay.nin : lay.nin + lay.ntail] = last[None, :] * np.exp(n[None, :] * aux["lr"][:, None])
    fl, fp = last * aux["wX_last"], prev * aux["wX_prev"]
    nx = (np.log(fl) - np.log(fp)) * aux["inv_dlog"]
    u[lay.nin + lay.ntail :] = fl[None, :] * np.exp(nx[None, :] * aux["lrx"][:, None])
    return u


def rows(pl, F, name):
    a, n = pl.front.rows[name]
    return F[a : a + n]


def antidiag(pl, cre, cim):
    """kernel `antidiag`: D[ch, t, 2, B] for t = 0..Nmax from the Hermitian half of c."""
    Nmax = pl.Nmax
    Nh = Nmax // 2
    c_half = cre + 1j * cim  # (Nh+1, B)
    full = np.concatenate([c_half, np.conj(c_half[:Nh][::-1])], axis=0)  # (Nmax+1, B)
    B = cre.shape[1]
    D = np.zeros((P.NCH, Nmax + 1, 2, B))
    off = pl.pair_offsets
    for t in range(Nmax + 1):
        n = np.arange(t // 2 + 1)
        prod = full[n] * full[t - n]  # (np, B)
        tab = pl.pair_table[off[t] : off[t + 1]]  # (np, NCH)
        d = tab.T @ prod
        D[:, t, 0], D[:, t, 1] = d.real, d.imag
    return D


def spectral(pl, D, Dcf=None):
    """GEMMs `Ak @ D_b` and `As[l] @ D_ch`; Dcf: anti-diagonal sums of the configuration-space coefficient set
    when it differs from the k-space one (IRcutoff "loop" / "resum")."""
    B = D.shape[-1]
    Dm = D.reshape(P.NCH, -1, B)
    P22 = np.einsum("kt,btB->bkB", pl.Ak, Dm[: P.N22])
    Cs = np.einsum("lst,ctB->lcsB", pl.As, (D if Dcf is None else Dcf).reshape(P.NCH, -1, B))
    return P22, Cs


def group(pl, F, P22, Cs, f):
    """kernel `group_terms`: Legendre weighting, f-power grouping, shot-noise subtraction
    (pybird.py:737-866).  Returns T[l, k, i, B] and Crows[l, r, s, B] (r: C11, Cct, Cloopl x12
    [, CctNNLO])."""
    g = pl.grid
    B = f.size
    P11 = rows(pl, F, "P11")
    P13 = (g.k**3)[None, :, None] * P11[None] * rows(pl, F, "P13raw").reshape(P.N13, g.Nk, B)
    T = np.zeros((g.Nl, g.Nk, g.nterm, B))
    Cr = np.zeros((g.Nl, 14 + (1 if g.with_NNLO else 0), g.Ns, B))
    Cr[:, 0] = rows(pl, F, "C11").reshape(g.Nl, g.Ns, B)
    Cr[:, 1] = rows(pl, F, "Cct").reshape(g.Nl, g.Ns, B)
    if g.with_NNLO:
        Cr[:, 14] = rows(pl, F, "CctNNLO").reshape(g.Nl, g.Ns, B)
    fp = f[None, :] ** np.arange(5)[:, None]
    for l in range(g.Nl):
        for i in range(3):
            T[l, :, i] = g.l11[l, i] * P11
        for i in range(6):
            T[l, :, 3 + i] = g.lct[l, i] * (g.k**2)[:, None] * P11
        for row, p, b in tables.GROUP22:
            T[l, :, 9 + row] += fp[p] * g.l22[l, b] * P22[b]
            Cr[l, 2 + row] += fp[p] * g.l22[l, b] * Cs[l, b]
        for row, p, b in tables.GROUP13:
            T[l, :, 9 + row] += fp[p] * g.l13[l, b] * P13[b]
            Cr[l, 2 + row] += fp[p] * g.l13[l, b] * Cs[l, P.N22 + b]
        T[l, :, 9:21] -= T[l, :1, 9:21]
        if g.with_NNLO:
            for i in range(3):
                T[l, :, 24 + i] = g.lctNNLO[l, i] * (g.k**4)[:, None] * P11
    T[0, :, 21] = 1.0
    T[0, :, 22] = (g.k**2)[:, None]
    if g.Nl >= 2:
        T[1, :, 23] = (g.k**2)[:, None]
    return T, Cr


def resum(pl, T, Cr, X, Y, f):
    """kernel `resum`: out[l,i,k] += sum_{l',s} T_a[l,l',k,s] C[l',i,s] with
    T_a = sum_v R[v,k,s] sum_p z^p (z Q_a[..,p*Na+v] + Y k^2 Q_a[..,(NIR+p)*Na+v]), z = k^2 X(s)."""
    g, rs = pl.grid, pl.resum
    NIR, Na, R, q = rs["NIR"], rs["Na"], rs["R"], rs["q"]
    B = f.size
    fpow = f[None, :] ** np.arange(q.shape[-1])[:, None]
    Q = np.einsum("alpud,dB->alpuB", q, fpow)  # (2, Nl, Nl, Nn, B)
    out = T.copy()
    k2 = rs["kr2"]
    for ik in range(g.Nkr):
        z = k2[ik] * X  # (Ns, B)
        yk = k2[ik] * Y
        Ta = np.zeros((2, g.Nl, g.Nl, g.Ns, B))
        for v in range(Na):
            A = np.zeros((2, g.Nl, g.Nl, g.Ns, B))
            Bq = np.zeros_like(A)
            for p in range(NIR - 1, -1, -1):
                A = A * z + Q[:, :, :, p * Na + v, None, :]
                Bq = Bq * z + Q[:, :, :, (NIR + p) * Na + v, None, :]
            S = z * A + yk * Bq
            Ta += R[v, ik][None, None, None, :, None] * S
        kk = g.Nklow + ik
        # linear: rows i<3 with l11 weights, a=0
        a11 = np.einsum("lpsB,psB->lpB", Ta[0], Cr[:, 0])
        out[:, kk, 0:3] += np.einsum("lpB,pi->liB", a11, g.l11)
        act = np.einsum("lpsB,psB->lpB", Ta[1], Cr[:, 1])
        out[:, kk, 3:9] += np.einsum("lpB,pi->liB", act, g.lct)
        out[:, kk, 9:21] += np.einsum("lpsB,pisB->liB", Ta[1], Cr[:, 2:14])
        if g.with_NNLO:
            an = np.einsum("lpsB,psB->lpB", Ta[1], Cr[:, 14])
            out[:, kk, 24:27] += np.einsum("lpB,pi->liB", an, g.lctNNLO)
    return out


def ap(pl, T, DA, H):
    """GEMM (B-spline coefficients) + kernel `ap`."""
    g, a = pl.grid, pl.ap
    B = DA.size
    qperp, qpar = DA / pl.ap_fid[0], pl.ap_fid[1] / H
    F = qpar / qperp
    nt = g.nterm if pl.ap_st else g.nterm - 3 if not g.with_NNLO else g.nterm
    coef = np.einsum("jk,lkiB->ljiB", a["Cinv"], T)  # (Nl, Nk, nterm, B)
    out = T.copy()
    mu, wl, lo, basis = a["mu"], a["wl"], a["knot_lo"], a["basis"]
    nint = a["nint"]
    norm = 1.0 / (qperp**2 * qpar)
    ap_rows = [i for i in range(g.nterm) if pl.ap_st or not (21 <= i < 24)]
    for ik, kv in enumerate(g.k):
        acc = np.zeros((g.Nl, g.nterm, B))
        for im, m in enumerate(mu):
            root = 1.0 + m * m * (F**-2 - 1.0)
            kp = kv / qperp * np.sqrt(root)
            mp = m / F / np.sqrt(root)
            j = np.clip(np.searchsorted(lo, kp, side="right") - 1, 0, nint - 1)  # (B,)
            x = kp - lo[j]
            bas = ((basis[j, :, 3] * x[:, None] + basis[j, :, 2]) * x[:, None] + basis[j, :, 1]) * x[:, None] + basis[j, :, 0]  # (B,4)
            L = [np.ones(B), 0.5 * (3 * mp**2 - 1), (35 * mp**4 - 30 * mp**2 + 3) / 8.0]
            val = np.zeros((g.nterm, B))
            for lp in range(g.Nl):
                for r in range(4):
                    c = coef[lp][j + r, :, np.arange(B)].T  # (nterm, B)
                    val += (L[lp] * bas[:, r])[None, :] * c
            acc += wl[:, im][:, None, None] * val[None]
        out[:, ik, ap_rows] = (norm[None, None, :] * acc)[:, ap_rows]
    return out


def project(pl, T):
    """GEMM `project @ T` over the (l,k) nodes + constant integral-constraint vector."""
    g = pl.grid
    B = T.shape[-1]
    out = np.einsum("on,niB->oiB", pl.project, T.reshape(g.Nl * g.Nk, g.nterm, B))
    return out


def run_chain(pl, plin, f, DA=None, H=None, upto="project"):
    u = front_inputs(pl, plin)
    F = pl.Wf @ u
    D = antidiag(pl, rows(pl, F, "cre"), rows(pl, F, "cim"))
    Dcf = antidiag(pl, rows(pl, F, "cre_cf"), rows(pl, F, "cim_cf")) if "cre_cf" in pl.front.rows else None
    P22, Cs = spectral(pl, D, Dcf)
    T, Cr = group(pl, F, P22, Cs, f)
    out = dict(F=F, D=D, P22=P22, Cs=Cs, T_pre=T, Cr=Cr)
    if pl.resum is not None:
        T = resum(pl, T, Cr, rows(pl, F, "X"), rows(pl, F, "Y"), f)
        out["T_res"] = T
    if pl.ap is not None:
        T = ap(pl, T, DA, H)
        out["T_ap"] = T
    if pl.project is not None:
        out["out"] = project(pl, T)
    return out
