"""Golden vectors from the UNMODIFIED reference for what round 1 only checked against the oracle:

  config3_like.npz   the DR16 NGC LRG x ELG x cross likelihood (BASELINE config 3: 142 data points, 14 marginalised
                     parameters) at B = 32 points, driven through the reference's own Cobaya components
                     (theory.py EFTLSS / EFTLeafKernel / EFTLeaf, likelihood.py EFTLike) by oracle/refshim/cobaya:
                     per-tracer reduced multipoles, PNG, PG, logp (Jeffreys and Gaussian priors), fullchi2, best fit
  reduce_kat.npz     parambasis.py on fixed term arrays: West-coast auto with NNLO (cr4, cr6), West-coast cross,
                     East-coast with and without NNLO (ctilde): `reduce_Plk(...).sum()` and the Gaussian tables
  nnlo_chain.npz     with_NNLO=True through PsCf / setPsCfl / Resum / AP / window / binning (PctNNLOl and the other
                     terms after every stage), B = 2, plus the reduced multipoles and the Gaussian table with cr4 / cr6

Build-container only (needs /root/reference).  Usage: python tests/golden/make_golden_config3.py
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

import refdriver  # noqa: E402
import refload  # noqa: E402
from eftpipe_b200 import synthetic  # noqa: E402

B = 32


def meta(script):
    return json.dumps(dict(numpy=np.__version__, scipy=scipy.__version__, reference="zhaoruiyang98/eftpipe 0.1.0 (unpinned sha)",
                           generated_by=script))


def config3():
    paths = refdriver.write_dr16("/tmp/dr16txt")
    tables = refdriver.synthetic_tables(B)
    cache = "/tmp/eftpipe_b200_bench_cache/ref"
    os.makedirs(cache, exist_ok=True)
    info = refdriver.config3_info(paths, tables, cache_dir=cache, likelihoods=("jeffreys", "gauss"))
    model = refdriver.reference_model(info)
    pts = refdriver.draw_points(B)
    out = dict(meta=meta("tests/golden/make_golden_config3.py"))
    for name, tab in tables.items():
        for k, v in tab.items():
            out[f"{name}.{k}"] = v
    for k, v in pts.items():
        out["pt." + k] = v
    likes = model.likelihood
    acc = {n: dict(logp=[], PNG=[], PG=[], fullchi2=[], best=[]) for n in likes}
    plk = {t: [] for t, _ in refdriver.TRACERS}
    t0 = time.time()
    for i in range(B):
        lp = model.logposterior({k: v[i] for k, v in pts.items()})
        for n, lk in likes.items():
            acc[n]["logp"].append(lk.current_logp)
            acc[n]["PNG"].append(lk.PNG().copy())
            acc[n]["PG"].append(lk.PG().copy())
            # marginal.py:79-140 once more for this likelihood's own best fit (the derived names are shared by the two)
            _, full, best = lk.marginalized_logp(return_bGbest=True, jeffreys=lk.jeffreys)
            acc[n]["fullchi2"].append(full)
            acc[n]["best"].append([best[g] for g in lk._bGidx_cache])
        eft = model.theory["eftpipe.eftlss"]
        for (t, _), ch in zip(refdriver.TRACERS, (False, True, False)):
            ls, k, P = eft.get_nonlinear_Plk_grid(t, chained=ch, binned=True)
            plk[t].append(P.copy())
            out[f"{t}.ls"], out[f"{t}.keff"] = np.array(ls), k
    print("reference: %d three-tracer evaluations in %.1fs" % (B, time.time() - t0))
    lk = likes["LEX_NGC"]
    out["data_vector"], out["invcov"] = lk.data_vector, lk.invcov
    out["gaussian_names"] = np.array(list(lk._bGidx_cache))
    for n in likes:
        for k, v in acc[n].items():
            out[f"{n}.{k}"] = np.array(v)
    for t in plk:
        out[f"{t}.Plk"] = np.array(plk[t])
    out["gauss_scales"] = json.dumps(refdriver.GAUSS_SCALES)
    np.savez_compressed(os.path.join(HERE, "config3_like.npz"), **out)
    print("config3_like.npz: logp", out["LEX_NGC.logp"][:3], "PG", out["LEX_NGC.PG"].shape)


class _Co:
    """the attributes parambasis.py reads from `bird.co`"""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def reduce_kat(ref):
    pb, tr = ref.parambasis, ref.transformer
    rng = np.random.default_rng(20261018 + 77)
    Bk, No, nk = 8, 3, 12
    scal = dict(kmA=0.7, krA=0.25, ndA=4.5e-5, kmB=0.8, krB=0.3, ndB=2.3e-4)
    amp = np.array([1e4, 3e3, 5e2])[None, :, None, None]
    terms = dict(P11l=amp * rng.normal(size=(Bk, No, 3, nk)), Ploopl=amp * rng.normal(size=(Bk, No, 12, nk)),
                 Pctl=amp * rng.normal(size=(Bk, No, 6, nk)), Pstl=rng.normal(size=(Bk, No, 3, nk)),
                 Picc=rng.normal(size=(Bk, No, nk)), PctNNLOl=amp * 0.1 * rng.normal(size=(Bk, No, 3, nk)))
    f = rng.uniform(0.6, 0.9, Bk)
    out = dict(meta=meta("tests/golden/make_golden_config3.py"), scales=json.dumps(scal), f=f, **{"T." + k: v for k, v in terms.items()})

    def bird(i, counterform, nnlo):
        co = _Co(No=No, counterform=counterform, with_NNLO=nnlo, **scal)
        return tr.PlainBird(f=f[i], co=co, P11l=terms["P11l"][i], Ploopl=terms["Ploopl"][i], Pctl=terms["Pctl"][i],
                            Pstl=terms["Pstl"][i], Picc=terms["Picc"][i], PctNNLOl=terms["PctNNLOl"][i])

    def run(tag, basis, counterform, nnlo, names):
        vals = {n: rng.normal(1.0, 0.7, Bk) for n in names}
        out[f"{tag}.params"] = json.dumps({n: v.tolist() for n, v in vals.items()})
        red, tabs = [], {}
        for i in range(Bk):
            p = {n: float(v[i]) for n, v in vals.items()}
            b = bird(i, counterform, nnlo)
            red.append(basis.reduce_Plk(b, p).sum())
            for g, arr in basis.reduce_Plk_gaussian_table(b, p).items():
                tabs.setdefault(g, []).append(arr)
        out[f"{tag}.reduced"] = np.array(red)
        out[f"{tag}.table_names"] = np.array(list(tabs))
        out[f"{tag}.table"] = np.array([tabs[g] for g in tabs])  # (ng, B, No, nk)

    west = ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2", "ce0", "cemono", "cequad")
    run("west_nnlo", pb.WestCoastBasis(prefix="w_"), "westcoast", True, ["w_" + n for n in west + ("cr4", "cr6")])
    run("west_auto", pb.WestCoastBasis(prefix="w_"), "westcoast", False, ["w_" + n for n in west])
    cross = [p + n for p in ("A_", "B_") for n in ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2")] + ["X_ce0", "X_cemono", "X_cequad"]
    run("west_cross", pb.WestCoastBasis(prefix="X_", cross_prefix=["A_", "B_"]), "westcoast", False, cross)
    run("west_cross_nnlo", pb.WestCoastBasis(prefix="X_", cross_prefix=["A_", "B_"]), "westcoast", True, cross + ["X_cr4", "X_cr6"])
    east = ("b1", "b2", "bG2", "bGamma3", "c0", "c2", "c4", "Pshot", "a0", "a2")
    run("east", pb.EastCoastBasis(prefix="e_"), "eastcoast", False, ["e_" + n for n in east])
    run("east_nnlo", pb.EastCoastBasis(prefix="e_"), "eastcoast", True, ["e_" + n for n in east + ("ctilde",)])
    np.savez_compressed(os.path.join(HERE, "reduce_kat.npz"), **out)
    print("reduce_kat.npz:", [k for k in out if k.endswith(".reduced")])


def nnlo_chain(ref):
    pb = ref.pybird
    z = 0.7
    batch = synthetic.make_batch(2, z, seed=20261018 + 5)
    kw = dict(Nl=3, No=3, kmax=0.3, kmA=0.7, krA=0.25, ndA=4.5e-5, with_NNLO=True)
    co = pb.Common(**kw)
    nl, rs = pb.NonLinear(load=False, save=False, co=co), pb.Resum(co=co)
    apkw = dict(Om_AP=0.307115, z_AP=0.696, APst=True)
    ap = pb.APeffect(rdrag_AP=147.66, h_AP=0.6777, co=co, **apkw)
    g2 = np.load(os.path.join(HERE, "config2_chain.npz"))
    Weff = g2["Weff_LRG"]  # the reference's own effective LRG window operator (make_golden.py)
    bin_r = ref.binning.Binning(g2["kout"], co=co)
    out = dict(meta=meta("tests/golden/make_golden_config3.py"), common=json.dumps(kw), ap=json.dumps(apkw), kin=batch.kin,
               plin=batch.plin, f=batch.f, DA=batch.DA, H=batch.H, z=z, kout=g2["kout"])
    names = ("P11l", "Pctl", "Ploopl", "Pstl", "PctNNLOl")
    st = {}
    basis = ref.parambasis.WestCoastBasis(prefix="")
    nuis = synthetic.draw_nuisance(2, seed=20261018 + 5)
    for i in range(2):
        b = pb.Bird(batch.kin, batch.plin[i], batch.f[i], batch.DA[i], batch.H[i], z, co=co, rdrag=synthetic.RDRAG, h=0.6777)
        nl.PsCf(b)
        b.setPsCfl()
        for n in names:
            st.setdefault("pre_" + n, []).append(getattr(b, n).copy())
        rs.Ps(b)
        for n in names:
            st.setdefault("res_" + n, []).append(getattr(b, n).copy())
        ap.AP(b)
        for n in names:
            st.setdefault("ap_" + n, []).append(getattr(b, n).copy())
        # window as the fixed operator of the reference's Window (every term array incl. PctNNLOl, window.py:389-415)
        for n in names:
            setattr(b, n, np.einsum("akln,lin->aik", Weff, getattr(b, n)))
        bl = bin_r.transform(b)
        for n in names + ("Picc",):
            st.setdefault("bin_" + n, []).append(getattr(bl, n).copy())
        b1, c2, b3, c4, cct, cr1, cr2, ce0, cemono, cequad = nuis[i]
        b2, b4 = synthetic.c2c4_to_b2b4(c2, c4)
        params = dict(b1=b1, b2=b2, b3=b3, b4=b4, cct=cct, cr1=cr1, cr2=cr2, ce0=ce0, cemono=cemono, cequad=cequad,
                      cr4=0.7 - 0.2 * i, cr6=-0.4 + 0.3 * i)
        st.setdefault("reduced", []).append(basis.reduce_Plk(bl, params).sum())
        tab = basis.reduce_Plk_gaussian_table(bl, params)
        st.setdefault("table", []).append(np.array([tab[n] for n in ("b3", "cct", "cr1", "cr2", "cr4", "cr6", "ce0", "cemono", "cequad")]))
        st.setdefault("params", []).append([params[n] for n in ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2", "ce0", "cemono", "cequad", "cr4", "cr6")])
    out.update({k: np.array(v) for k, v in st.items()})
    np.savez_compressed(os.path.join(HERE, "nnlo_chain.npz"), **out)
    print("nnlo_chain.npz: PctNNLOl after resum+AP", out["ap_PctNNLOl"].shape)


def main():
    ref = refload.load()
    reduce_kat(ref)
    nnlo_chain(ref)
    config3()


if __name__ == "__main__":
    main()
