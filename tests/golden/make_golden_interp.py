"""Golden vectors for the un-binned likelihood products: `PlkInterpolator` (theory.py:75-106)
for the non-marginalised part, `interp1d(k, k P)(kout)/kout` for the marginalised rows (likelihood.py:503-547), and
the raw-grid mode (`with_binning: false, with_interp: false`).  Runs the UNMODIFIED reference classes
(`eftpipe.theory.PlkInterpolator`, `WestCoastBasis`, `Marginalizable`) on a single-tracer LRG chain (one-loop + IR
resummation + AP, no window) and writes tests/golden/interp_kat.npz.  Build-container only.
Usage:  python tests/golden/make_golden_interp.py
"""
from __future__ import annotations

import importlib
import json
import os
import sys

import numpy as np
from scipy.interpolate import interp1d

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REFDATA, meta, orc, refload, relerr, synthetic  # noqa: E402

GAUSS = ("b3", "cct", "cr1", "cr2", "ce0", "cequad")


def main():
    ref = refload.load()
    pb = ref.pybird
    th = importlib.import_module("eftpipe.theory")
    lk = ref.likelihood
    z = 0.7
    n = 3
    batch = synthetic.make_batch(n, z, seed=20261018 + 11)
    nuis = synthetic.draw_nuisance(n, seed=20261018 + 11)
    kw = dict(Nl=3, No=3, kmax=0.3, kmA=0.7, krA=0.25, ndA=4.5e-5)
    apkw = dict(Om_AP=0.307115, z_AP=0.696, APst=True)
    co_r, co_o = pb.Common(**kw), orc.Common(**kw)
    nl_r, rs_r = pb.NonLinear(load=False, save=False, co=co_r), pb.Resum(co=co_r)
    ap_r = pb.APeffect(rdrag_AP=147.66, h_AP=0.6777, co=co_r, **apkw)
    nl_o, rs_o, ap_o = orc.NonLinear(co_o), orc.Resum(co_o), orc.APeffect(co_o, **apkw)

    dat = np.loadtxt(os.path.join(REFDATA, "NGC_LRG_P.txt"))
    kall = dat[:, 0]
    sel = (kall >= 0.02) & (kall <= 0.20)
    kout = kall[sel]
    cov = np.loadtxt(os.path.join(REFDATA, "cov_NGC_L024_P.txt"))
    covm = lk.mask_covariance(cov / lk.hartlap(1000, 3 * kout.size), [0, 2, 4], [0, 2, 4], kall, 0.02, 0.20)
    invcov = np.linalg.inv(covm)
    dvec = np.hstack([dat[sel, 1 + i] for i in range(3)])
    # raw-grid mode: the "data" live on the internal nodes (likelihood.py:546-547 flattens without a mask)
    rng = np.random.default_rng(20261018)
    nraw = 3 * co_r.Nk
    A = rng.normal(size=(nraw, nraw))
    invcov_raw = A @ A.T / nraw + np.eye(nraw)

    class _Marg(ref.marginal.Marginalizable):
        def __init__(self, PNG, PG, data, icov):
            self._png, self._pg, self._d, self._ic = PNG, PG, data, icov
            self.valid_prior = {f"p{i}": {"loc": 0.0, "scale": np.inf} for i in range(PG.shape[0])}
            self._sigma_inv = np.zeros((PG.shape[0],) * 2)
        PNG = lambda self: self._png
        PG = lambda self: self._pg
        get_data_vector = lambda self: self._d
        get_invcov = lambda self: self._ic
        mpi_debug = lambda self, *a: None
        mpi_warning = lambda self, *a: None

    gold = dict(meta=meta(), kin=batch.kin, plin=batch.plin, f=batch.f, DA=batch.DA, H=batch.H, z=z, kout=kout,
                common=json.dumps(kw), ap=json.dumps(apkw), nuisance=nuis, invcov=invcov, data=dvec,
                invcov_raw=invcov_raw, gauss=json.dumps(GAUSS))
    worst = 0.0
    out = {k: [] for k in ("Plk", "PNG_interp", "PG_interp", "logp_interp", "PNG_raw", "PG_raw", "logp_raw", "data_raw")}
    basis = ref.parambasis.WestCoastBasis(prefix="")
    for i in range(n):
        bird = pb.Bird(batch.kin, batch.plin[i], batch.f[i], batch.DA[i], batch.H[i], z, co=co_r, rdrag=synthetic.RDRAG, h=0.6777)
        nl_r.PsCf(bird); bird.setPsCfl(); rs_r.Ps(bird); ap_r.AP(bird)
        ob = orc.Bird(co_o, batch.kin, batch.plin[i], batch.f[i], batch.DA[i], batch.H[i], z)
        nl_o.PsCf(ob); orc.set_PsCfl(ob); rs_o.Ps(ob); ap_o.AP(ob)
        b1, c2, b3, c4, *_ = nuis[i]
        b2, b4 = synthetic.c2c4_to_b2b4(c2, c4)
        params = dict(b1=b1, b2=b2, b4=b4)  # marginalised parameters are not sampled
        Plk = basis.reduce_Plk(bird, params).sum()
        table = basis.reduce_Plk_gaussian_table(bird, params)
        fn = th.PlkInterpolator([0, 2, 4], co_r.k, Plk)
        PNG = fn([0, 2, 4], kout).reshape(-1)  # likelihood.py:541-544 (kout_mask selects everything here)
        PG = np.array([(interp1d(co_r.k, co_r.k * table[g], kind="cubic", axis=-1)(kout) / kout).reshape(-1) for g in GAUSS])
        lp = _Marg(PNG, PG, dvec, invcov).marginalized_logp(jeffreys=True)
        # oracle check
        Po = orc.reduce_Plk(co_o, batch.f[i], orc.bird_terms(ob), (b1, b2, 0.0, b4, 0.0, 0.0, 0.0))
        to = orc.gaussian_table_west(co_o, batch.f[i], orc.bird_terms(ob), b1)
        PNGo = orc.plk_interpolator(co_o.k, Po)(kout).reshape(-1)
        PGo = np.array([orc.gaussian_row_interp(co_o.k, to[g], kout).reshape(-1) for g in GAUSS])
        lpo = orc.marginalized_logp(PNGo, PGo, dvec, invcov, jeffreys=True)
        e = max(relerr(PNGo, PNG), relerr(PGo, PG), abs(lpo - lp) / abs(lp))
        worst = max(worst, e)
        print(f"[interp {i}] oracle vs reference: PNG {relerr(PNGo, PNG):.2e} PG {relerr(PGo, PG):.2e} logp {abs(lpo - lp) / abs(lp):.2e}")
        # raw grid: every node of every multipole; synthetic data = the first cosmology's spectrum plus noise
        PNGr = Plk.reshape(-1)
        PGr = np.array([table[g].reshape(-1) for g in GAUSS])
        if i == 0:
            data_raw = PNGr * (1.0 + 0.01 * rng.normal(size=nraw))
        lpr = _Marg(PNGr, PGr, data_raw, invcov_raw).marginalized_logp(jeffreys=True)
        lpro = orc.marginalized_logp(Po.reshape(-1), np.array([to[g].reshape(-1) for g in GAUSS]), data_raw, invcov_raw, jeffreys=True)
        worst = max(worst, abs(lpro - lpr) / abs(lpr))
        for k_, v in (("Plk", Plk), ("PNG_interp", PNG), ("PG_interp", PG), ("logp_interp", lp), ("PNG_raw", PNGr),
                      ("PG_raw", PGr), ("logp_raw", lpr)):
            out[k_].append(v)
    gold["data_raw"] = data_raw
    for k_, v in out.items():
        if k_ != "data_raw":
            gold[k_] = np.array(v)
    np.savez_compressed(os.path.join(HERE, "interp_kat.npz"), **gold)
    print("worst oracle-vs-reference error: %.3e" % worst)
    assert worst < 1e-9


if __name__ == "__main__":
    main()


def linear_power_file_golden():
    """reference `LinearPowerFile` (boltzmann.py:246-309) on a synthetic template that starts above k = 1e-5"""
    import tempfile

    ref = refload.load()
    bz = importlib.import_module("eftpipe.boltzmann")
    b = synthetic.make_batch(1, 0.7, seed=3)
    k = np.logspace(-4, 0.3, 300)
    pk = np.exp(np.interp(np.log(k), np.log(b.kin), np.log(b.plin[0])))
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as fh:
        np.savetxt(fh, np.column_stack([k, pk]))
        path = fh.name
    ext = bz.LinearPowerFile(path, gz=0.8, prefix="t_")
    kh = np.logspace(-5, 0, 200)
    np.savez_compressed(os.path.join(HERE, "linear_power_file.npz"), meta=meta(), k=k, pk=pk, gz=0.8, kh=kh, pkh=ext.Pkh(kh),
                        requirements=json.dumps(sorted(ext.get_requirements())))
    os.unlink(path)


if __name__ == "__main__" and os.environ.get("EFTB_GOLDEN_LPF", "1") == "1":
    linear_power_file_golden()
