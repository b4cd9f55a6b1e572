"""Generate golden vectors by running the UNMODIFIED reference (zhaoruiyang98/eftpipe,
mounted at /root/reference) on the synthetic inputs of `eftpipe_b200.synthetic`, check the
oracle (`oracle/pybird_oracle.py`) against it, and write `tests/golden/*.npz`.

Build-container only (needs /root/reference).  Versions used for the committed fixtures are
recorded inside each file (`meta`).  Usage:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import refload  # noqa: E402
import pybird_oracle as orc  # noqa: E402
from eftpipe_b200 import synthetic  # noqa: E402

REFDATA = os.path.join(refload.REFERENCE_ROOT, "data", "DR16_noric")


def relerr(a, b):
    """max |a-b| / max|b| over the trailing axis (row-max relative)."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max(axis=-1, keepdims=True)
    scale = np.where(scale == 0, 1.0, scale)
    return float(np.max(np.abs(a - b) / scale))


def meta():
    return json.dumps(dict(numpy=np.__version__, scipy=scipy.__version__,
                           reference="zhaoruiyang98/eftpipe 0.1.0 (unpinned sha)",
                           generated_by="tests/golden/make_golden.py"))


def ref_chain(ref, co, nl, rs, kin, plin, f, DA, H, z, ap=None, win=None, binning=None):
    pb = ref.pybird
    out = {}
    bird = pb.Bird(kin, plin, f, DA, H, z, co=co, rdrag=synthetic.RDRAG, h=0.6777)
    out["P11"] = bird.P11.copy()
    out["coef"] = nl.Coef(bird, window=0.2)
    nl.PsCf(bird)
    for n in ("P22", "P13", "C11", "Cct", "C22", "C13"):
        out[n] = getattr(bird, n).copy()
    bird.setPsCfl()
    for n in ("P11l", "Pctl", "Ploopl", "Cloopl", "Pstl"):
        out["pre_" + n] = getattr(bird, n).copy()
    X, Y = rs.IRFilters(bird)
    out["X"], out["Y"] = X, Y
    rs.Ps(bird)
    for n in ("P11l", "Pctl", "Ploopl"):
        out["res_" + n] = getattr(bird, n).copy()
    if ap is not None:
        ap.AP(bird)
        for n in ("P11l", "Pctl", "Ploopl", "Pstl"):
            out["ap_" + n] = getattr(bird, n).copy()
    if win is not None:
        win.Window(bird)
        for n in ("P11l", "Pctl", "Ploopl", "Pstl", "Picc"):
            out["win_" + n] = getattr(bird, n).copy()
    if binning is not None:
        bl = binning.transform(bird)
        for n in ("P11l", "Pctl", "Ploopl", "Pstl", "Picc"):
            out["bin_" + n] = getattr(bl, n).copy()
        ch = ref.chained.Chained().transform(bl)
        for n in ("P11l", "Pctl", "Ploopl", "Pstl", "Picc"):
            out["chn_" + n] = getattr(ch, n).copy()
        out["_binned"], out["_chained"] = bl, ch
    out["_bird"] = bird
    return out


def orc_chain(co, nl, rs, kin, plin, f, DA, H, z, ap=None, win=None, binning=None):
    out = {}
    bird = orc.Bird(co, kin, plin, f, DA, H, z)
    out["P11"] = bird.P11.copy()
    nl.PsCf(bird)
    out["coef"] = bird.coef
    for n in ("P22", "P13", "C11", "Cct", "C22", "C13"):
        out[n] = getattr(bird, n).copy()
    orc.set_PsCfl(bird)
    for n in ("P11l", "Pctl", "Ploopl", "Cloopl", "Pstl"):
        out["pre_" + n] = getattr(bird, n).copy()
    rs.Ps(bird)
    out["X"], out["Y"] = bird.X, bird.Y
    for n in ("P11l", "Pctl", "Ploopl"):
        out["res_" + n] = getattr(bird, n).copy()
    if ap is not None:
        ap.AP(bird)
        for n in ("P11l", "Pctl", "Ploopl", "Pstl"):
            out["ap_" + n] = getattr(bird, n).copy()
    if win is not None:
        orc.apply_window(bird, **win)
        for n in ("P11l", "Pctl", "Ploopl", "Pstl", "Picc"):
            out["win_" + n] = getattr(bird, n).copy()
    if binning is not None:
        bl = binning.transform(orc.bird_terms(bird))
        for n in ("P11l", "Pctl", "Ploopl", "Pstl", "Picc"):
            out["bin_" + n] = bl[n].copy()
        ch = orc.chained_transform(bl, co.Nl)
        for n in ("P11l", "Pctl", "Ploopl", "Pstl", "Picc"):
            out["chn_" + n] = ch[n].copy()
        out["_binned"], out["_chained"] = bl, ch
    out["_bird"] = bird
    return out


def compare(tag, r, o, tol=2e-11):
    worst = 0.0
    for key in r:
        if key.startswith("_"):
            continue
        e = relerr(o[key], r[key])
        worst = max(worst, e)
        flag = "" if e < tol else "   <-- ABOVE TOL"
        print(f"  [{tag}] {key:12s} oracle vs reference rowmax-rel err = {e:.2e}{flag}")
    return worst


def main():
    t0 = time.time()
    ref = refload.load()
    pb = ref.pybird
    worst = 0.0

    # ------------------------------------------------------------------ config 1 (+2 stages)
    z = 0.7
    batch = synthetic.make_batch(3, z, seed=20261018 + 1)
    kw = dict(Nl=3, No=3, kmax=0.3, kmA=0.7, krA=0.25, ndA=4.5e-5)
    co_r = pb.Common(**kw)
    nl_r = pb.NonLinear(load=False, save=False, co=co_r)
    rs_r = pb.Resum(co=co_r)
    co_o = orc.Common(**kw)
    nl_o = orc.NonLinear(co_o)
    rs_o = orc.Resum(co_o)
    print("M22 oracle vs ref", relerr(nl_o.M22.reshape(28, -1), nl_r.M22.reshape(28, -1)),
          "M13", relerr(nl_o.M13, nl_r.M13))

    apkw = dict(Om_AP=0.307115, z_AP=0.696, APst=True)
    ap_r = pb.APeffect(rdrag_AP=147.66, h_AP=0.6777, co=co_r, **apkw)
    ap_o = orc.APeffect(co_o, **apkw)

    # window: DR16 NGC LRG, production settings (accboost 4, windowk 0.1)
    winfile = os.path.join(REFDATA, "win_NGC_LRG.txt")
    tw = time.time()
    win_r = ref.window.Window(window_configspace_file=winfile, co=co_r, accboost=4, windowk=0.1,
                              load=False, save=False)
    print("reference Window build %.1fs" % (time.time() - tw))
    sQ = np.loadtxt(winfile)
    tw = time.time()
    Wal_o, p_o = orc.compute_Wal(sQ, co_o, Na=3, Nl=3, accboost=4)
    print("oracle Window build %.1fs; Wal err %.2e" % (time.time() - tw, relerr(Wal_o, win_r.Wal)))
    worst = max(worst, relerr(Wal_o, win_r.Wal))
    Waldk_o = orc.mask_and_measure(Wal_o, p_o, co_o.k, windowk=0.1)

    # synthetic integral-constraint matrices: Wal_ic = 0.05 Wal, PSN ~ 1e-3/k
    Pshot = 1.0 / 4.5e-5
    PSN = 1e-3 / co_r.k[None, :] * np.array([1.0, 0.3, 0.1])[:, None]
    icc_npz = "/tmp/_golden_icc.npz"
    np.savez(icc_npz, PSN=PSN, Wal=0.05 * win_r.Wal)
    icc_r = ref.icc.IntegralConstraint(Pshot=Pshot, icc_fourier_file=icc_npz, co=co_r, accboost=4,
                                       windowk=0.1, save=False, check_meta=False)
    win_r.icc = icc_r
    Waldk_ic_o = orc.mask_and_measure(0.05 * Wal_o, p_o, co_o.k, windowk=0.1)
    win_o = dict(Waldk=Waldk_o, p=p_o, window_st=True, icc=(Waldk_ic_o, PSN * Pshot))

    kdata = np.loadtxt(os.path.join(REFDATA, "NGC_LRG_P.txt"))[:, 0]
    kout = kdata[(kdata >= 0.02) & (kdata <= 0.20)]
    bin_r = ref.binning.Binning(kout, co=co_r)
    bin_o = orc.Binning(kout, co_o)
    print("binning keff err", relerr(bin_o.keff, bin_r.keff), "nbin", kout.size)

    gold = dict(meta=meta(), kin=batch.kin, plin=batch.plin, f=batch.f, DA=batch.DA, H=batch.H,
                z=z, kout=kout, common=json.dumps(kw), ap=json.dumps(apkw), PSN=PSN, Pshot=Pshot)
    stages = {}
    nuis = synthetic.draw_nuisance(len(batch), seed=20261018 + 1)
    reduced, gtables = [], []
    for i in range(len(batch)):
        args = (batch.kin, batch.plin[i], batch.f[i], batch.DA[i], batch.H[i], z)
        r = ref_chain(ref, co_r, nl_r, rs_r, *args, ap=ap_r, win=win_r, binning=bin_r)
        o = orc_chain(co_o, nl_o, rs_o, *args, ap=ap_o, win=win_o, binning=bin_o)
        worst = max(worst, compare(f"cfg2 cosmology {i}", r, o))
        for key, val in r.items():
            if not key.startswith("_"):
                stages.setdefault(key, []).append(val)
        # bias reduction on the binned (unchained) and chained products
        b1, c2, b3, c4, cct, cr1, cr2, ce0, cemono, cequad = nuis[i]
        b2, b4 = synthetic.c2c4_to_b2b4(c2, c4)
        basis = ref.parambasis.WestCoastBasis(prefix="")
        params = dict(b1=b1, b2=b2, b3=b3, b4=b4, cct=cct, cr1=cr1, cr2=cr2, ce0=ce0,
                      cemono=cemono, cequad=cequad)
        for kind in ("_binned", "_chained"):
            pr = basis.reduce_Plk(r[kind], params).sum()
            po = orc.reduce_Plk(co_o, batch.f[i], o[kind], (b1, b2, b3, b4, cct, cr1, cr2),
                                es=(ce0, cemono, cequad))
            e = relerr(po, pr)
            worst = max(worst, e)
            print(f"  [cfg2 cosmology {i}] reduce{kind} err = {e:.2e}")
            tr = basis.reduce_Plk_gaussian_table(r[kind], params)
            to = orc.gaussian_table_west(co_o, batch.f[i], o[kind], b1)
            for name in to:
                worst = max(worst, relerr(to[name], tr[name]))
            if kind == "_binned":
                reduced.append(pr)
                gtables.append(np.array([tr[n] for n in ("b3", "cct", "cr1", "cr2", "ce0", "cemono", "cequad")]))
            else:
                gold.setdefault("reduced_chained", []).append(pr)
    for key, vals in stages.items():
        gold[key] = np.array(vals)
    gold["nuisance"] = nuis
    gold["reduced_binned"] = np.array(reduced)
    gold["gaussian_table_binned"] = np.array(gtables)
    gold["reduced_chained"] = np.array(gold["reduced_chained"])

    # single-tracer marginalised likelihood on the LRG data/cov (likelihood.py:337-363)
    lk = ref.likelihood
    cov = np.loadtxt(os.path.join(REFDATA, "cov_NGC_L024_P.txt"))
    dat = np.loadtxt(os.path.join(REFDATA, "NGC_LRG_P.txt"))
    kall = dat[:, 0]
    covm = lk.mask_covariance(cov / lk.hartlap(1000, 3 * kout.size), [0, 2, 4], [0, 2, 4], kall, 0.02, 0.20)
    invcov = np.linalg.inv(covm)
    sel = (kall >= 0.02) & (kall <= 0.20)
    dvec = np.hstack([dat[sel, 1 + i] for i in range(3)])
    gold["lrg_invcov"], gold["lrg_data"] = invcov, dvec

    class _Marg(ref.marginal.Marginalizable):
        def __init__(self, PNG, PG, scales):
            self._png, self._pg = PNG, PG
            self.valid_prior = {f"p{i}": {"loc": 0.0, "scale": s} for i, s in enumerate(scales)}
            self._sigma_inv = np.zeros((len(scales),) * 2)
        PNG = lambda self: self._png
        PG = lambda self: self._pg
        get_data_vector = lambda self: dvec
        get_invcov = lambda self: invcov
        mpi_debug = lambda self, *a: None
        mpi_warning = lambda self, *a: None

    logps = []
    use = [0, 1, 2, 3, 4, 6]  # b3 cct cr1 cr2 ce0 cequad (production excludes cemono)
    for i in range(len(batch)):
        PG = gold["gaussian_table_binned"][i][use].reshape(len(use), -1)
        # PNG: gaussian parameters set to zero (likelihood.py: marginalised params are not sampled)
        b1, c2, b3, c4, *_ = nuis[i]
        b2, b4 = synthetic.c2c4_to_b2b4(c2, c4)
        basis = ref.parambasis.WestCoastBasis(prefix="")
        binned_i = ref_chain(ref, co_r, nl_r, rs_r, batch.kin, batch.plin[i], batch.f[i], batch.DA[i],
                             batch.H[i], z, ap=ap_r, win=win_r, binning=bin_r)["_binned"]
        PNG = basis.reduce_Plk(binned_i, dict(b1=b1, b2=b2, b4=b4)).sum().reshape(-1)
        row = []
        for scales, jeff in (([np.inf] * 6, True), ([4, 2, 4, 4, 2, 2], False)):
            m = _Marg(PNG, PG, scales)
            lp_r, full_r, best_r = m.marginalized_logp(return_bGbest=True, jeffreys=jeff)
            sig = np.zeros((6, 6)) if np.inf in scales else np.diag(1.0 / np.array(scales, float) ** 2)
            lp_o, full_o, best_o = orc.marginalized_logp(PNG, PG, dvec, invcov, sigma_inv=sig,
                                                         jeffreys=jeff, return_bestfit=True)
            e = abs(lp_o - lp_r) / abs(lp_r)
            worst = max(worst, e)
            print(f"  [marg {i} jeffreys={jeff}] logp ref={lp_r:.10g} oracle rel err={e:.2e}")
            row += [lp_r, full_r] + list(best_r.values())
        logps.append(row)
        gold.setdefault("marg_PNG", []).append(PNG)
    gold["marg_PNG"] = np.array(gold["marg_PNG"])
    gold["marg_out"] = np.array(logps)  # [logp_jeff, fullchi2, best(6), logp_gauss, fullchi2, best(6)]

    # effective window operator (Weff[a,k,l,k'] applied on the co.k nodes) - compact fixture
    eye = np.eye(co_r.Nk)
    Weff = np.zeros((3, co_r.Nk, 3, co_r.Nk))
    for l in range(3):
        P = np.zeros((3, co_r.Nk, co_r.Nk))
        P[l] = eye
        Weff[:, :, l, :] = win_r.integrWindow(P).transpose(0, 2, 1)
    gold["Weff_LRG"] = Weff
    np.savez_compressed(os.path.join(HERE, "config2_chain.npz"), **gold)

    # ------------------------------------------------------------------ FFTLog unit vectors
    # the reference's own property test setup (tests/compare/test_fftlog.py:5-23)
    fl_r = ref.fftlog.FFTLog(Nmax=256, xmin=1e-5, xmax=10, bias=-0.3)
    x = np.logspace(-4, 0.5, 300)
    rng = np.random.default_rng(5)
    rows = np.array([np.exp(-0.5 * (np.log(x) - m) ** 2 / s**2)
                     for m, s in zip(rng.uniform(-6, -1, 8), rng.uniform(0.3, 1.5, 8))])
    cr = fl_r.Coef(x, rows, extrap="padding", window=0.3)
    g = orc.LogGrid(Nmax=256, xmin=1e-5, xmax=10, bias=-0.3)
    co_ = orc.fftlog_coef(g, x, rows, extrap="padding", window=0.3)
    e = relerr(np.abs(co_ - cr), np.abs(cr)) if False else float(np.abs(co_ - cr).max() / np.abs(cr).max())
    print("fftlog padding/window=0.3 err", e)
    worst = max(worst, e)
    np.savez_compressed(os.path.join(HERE, "fftlog_kat.npz"), meta=meta(), x=x, rows=rows, coef=cr,
                        hubble=pb.Hubble(0.2, 1.0), dafunc=pb.DAfunc(0.2, 1.0))

    # ------------------------------------------------------------------ Nl=2 variant (NIR=8)
    kw2 = dict(Nl=2, No=2, kmax=0.3, kmA=0.7, krA=0.25, ndA=3e-4)
    co_r2, co_o2 = pb.Common(**kw2), orc.Common(**kw2)
    nl_r2, rs_r2 = pb.NonLinear(load=False, save=False, co=co_r2), pb.Resum(co=co_r2)
    nl_o2, rs_o2 = orc.NonLinear(co_o2), orc.Resum(co_o2)
    g2 = dict(meta=meta(), common=json.dumps(kw2), kin=batch.kin, plin=batch.plin[:2], f=batch.f[:2], z=z)
    st2 = {}
    for i in range(2):
        args = (batch.kin, batch.plin[i], batch.f[i], batch.DA[i], batch.H[i], z)
        r = ref_chain(ref, co_r2, nl_r2, rs_r2, *args)
        o = orc_chain(co_o2, nl_o2, rs_o2, *args)
        worst = max(worst, compare(f"Nl=2 cosmology {i}", r, o))
        for key in ("P11l", "Pctl", "Ploopl"):
            st2.setdefault("res_" + key, []).append(r["res_" + key])
    for key, vals in st2.items():
        g2[key] = np.array(vals)
    np.savez_compressed(os.path.join(HERE, "nl2_resum.npz"), **g2)

    # ---- fibre collisions: reference FiberCollision.dPcorr / fibcolWindow on random term arrays (eastcoast mock yaml
    # values tests/yamls/mock_eBOSS_LRG_ELG_NGC_all_like.yaml:30-39)
    fkw = dict(fs=0.6, Dfc=0.43 / 0.6777, ktrust=0.25)
    co_r3, co_o3 = pb.Common(Nl=3), orc.Common(Nl=3)
    fc = pb.FiberCollision(co=co_r3, **fkw)
    rng = np.random.default_rng(20261018)
    PS = rng.normal(size=(3, 4, co_r3.Nk)) * np.array([1e4, 1e3, 1e2])[:, None, None]
    d_ref = fc.dPcorr(co_r3.k, co_r3.k, PS, **fkw)
    d_orc = orc.fiber_dPcorr(co_o3, PS, **fkw)
    worst = max(worst, relerr(d_orc, d_ref))
    print("fibre dPcorr: oracle vs reference %.2e" % relerr(d_orc, d_ref))
    np.savez_compressed(os.path.join(HERE, "fiber_kat.npz"), meta=meta(), fiber=json.dumps(fkw), PS=PS, dPcorr=d_ref)

    print("worst oracle-vs-reference error: %.3e   (%.0fs)" % (worst, time.time() - t0))
    assert worst < 1e-9, "oracle does not reproduce the reference"


if __name__ == "__main__":
    main()
