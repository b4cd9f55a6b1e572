"""CPU: the oracle against the live-reference goldens of tests/golden/make_golden_config3.py - the bias reductions the
first round only compared GPU <-> oracle (West-coast cross, East-coast, NNLO), the with_NNLO chain, and the DR16 NGC
three-tracer likelihood (BASELINE config 3) as the reference's own Cobaya components computed it."""
import json
import os

import numpy as np
import pytest

import pybird_oracle as orc
from conftest import GOLDEN, rowmax_rel

TOL = 5e-11


@pytest.fixture(scope="module")
def kat():
    return dict(np.load(os.path.join(GOLDEN, "reduce_kat.npz")))


def _terms(kat, i):
    return {n: kat["T." + n][i] for n in ("P11l", "Ploopl", "Pctl", "Pstl", "Picc", "PctNNLOl")}


def _co(kat, counterform, nnlo):
    co = orc.Common(Nl=3, counterform=counterform, with_NNLO=nnlo, **json.loads(str(kat["scales"])))
    return co


@pytest.mark.parametrize("tag,nnlo", [("west_auto", False), ("west_nnlo", True)])
def test_west_auto_reduction(kat, tag, nnlo):
    # the fixture keeps distinct A / B scales also on the auto basis: parambasis.py:68-75 reads them from `co` as they are
    p = {k: np.array(v) for k, v in json.loads(str(kat[tag + ".params"])).items()}
    names = [str(n) for n in kat[tag + ".table_names"]]
    for i in range(kat["f"].size):
        g = lambda n: p["w_" + n][i]
        bsA = [g(n) for n in ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2")]
        es = [g(n) for n in ("ce0", "cemono", "cequad")]
        co2 = _co(kat, "westcoast", nnlo)
        red = orc.reduce_Plk(co2, kat["f"][i], _terms(kat, i), bsA, es=es, cnnlo=(g("cr4"), g("cr6")) if nnlo else None)
        assert rowmax_rel(red, kat[tag + ".reduced"][i]) <= TOL
        tab = orc.gaussian_table_west(co2, kat["f"][i], _terms(kat, i), bsA[0])
        for j, n in enumerate(names):
            assert rowmax_rel(tab[n[2:]], kat[tag + ".table"][j, i]) <= TOL, n
    assert ("w_cr4" in names) == nnlo


@pytest.mark.parametrize("tag,nnlo", [("west_cross", False), ("west_cross_nnlo", True)])
def test_west_cross_reduction(kat, tag, nnlo):
    co = _co(kat, "westcoast", nnlo)
    p = {k: np.array(v) for k, v in json.loads(str(kat[tag + ".params"])).items()}
    names = [str(n) for n in kat[tag + ".table_names"]]
    assert "X_cr4" not in names  # the reference has no NNLO rows for a cross (parambasis.py:277-296)
    for i in range(kat["f"].size):
        bsA = [p["A_" + n][i] for n in ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2")]
        bsB = [p["B_" + n][i] for n in ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2")]
        es = [p["X_" + n][i] for n in ("ce0", "cemono", "cequad")]
        cn = (p["X_cr4"][i], p["X_cr6"][i]) if nnlo else None
        red = orc.reduce_Plk(co, kat["f"][i], _terms(kat, i), bsA, bsB, es=es, cnnlo=cn)
        assert rowmax_rel(red, kat[tag + ".reduced"][i]) <= TOL
        tab = orc.gaussian_table_west(co, kat["f"][i], _terms(kat, i), bsA[0], bsB[0], cross=True)
        for j, n in enumerate(names):
            key = n[2:] if n.startswith("X_") else n
            assert rowmax_rel(tab[key], kat[tag + ".table"][j, i]) <= TOL, n


@pytest.mark.parametrize("tag,nnlo", [("east", False), ("east_nnlo", True)])
def test_east_reduction(kat, tag, nnlo):
    co = _co(kat, "eastcoast", nnlo)
    p = {k: np.array(v) for k, v in json.loads(str(kat[tag + ".params"])).items()}
    names = [str(n) for n in kat[tag + ".table_names"]]
    for i in range(kat["f"].size):
        f = kat["f"][i]
        vals = [p["e_" + n][i] for n in ("b1", "b2", "bG2", "bGamma3", "c0", "c2", "c4", "Pshot", "a0", "a2")]
        bsA, es = orc.east_to_west(f, *vals)
        red = orc.reduce_Plk(co, f, _terms(kat, i), bsA, es=es, cnnlo=(p["e_ctilde"][i],) if nnlo else None)
        assert rowmax_rel(red, kat[tag + ".reduced"][i]) <= TOL
        tab = orc.gaussian_table_east(co, f, _terms(kat, i), vals[0])
        for j, n in enumerate(names):
            assert rowmax_rel(tab[n[2:]], kat[tag + ".table"][j, i]) <= TOL, n
    assert ("e_ctilde" in names) == nnlo


def test_nnlo_chain():
    g = dict(np.load(os.path.join(GOLDEN, "nnlo_chain.npz")))
    co = orc.Common(**json.loads(str(g["common"])))
    nl, rs = orc.NonLinear(co), orc.Resum(co)
    ap = orc.APeffect(co, **json.loads(str(g["ap"])))
    names = ("P11l", "Pctl", "Ploopl", "Pstl", "PctNNLOl")
    for i in range(g["plin"].shape[0]):
        b = orc.Bird(co, g["kin"], g["plin"][i], g["f"][i], g["DA"][i], g["H"][i], float(g["z"]))
        nl.PsCf(b)
        orc.set_PsCfl(b)
        for n in names:
            assert rowmax_rel(getattr(b, n), g["pre_" + n][i]) <= TOL, n
        rs.Ps(b)
        for n in names:
            assert rowmax_rel(getattr(b, n), g["res_" + n][i]) <= TOL, n
        ap.AP(b)
        for n in names:
            assert rowmax_rel(getattr(b, n), g["ap_" + n][i]) <= TOL, n
        # reduction of the reference's binned terms with cr4 / cr6
        terms = {n: g["bin_" + n][i] for n in names + ("Picc",)}
        b1, b2, b3, b4, cct, cr1, cr2, ce0, cemono, cequad, cr4, cr6 = g["params"][i]
        red = orc.reduce_Plk(co, g["f"][i], terms, [b1, b2, b3, b4, cct, cr1, cr2], es=[ce0, cemono, cequad], cnnlo=(cr4, cr6))
        assert rowmax_rel(red, g["reduced"][i]) <= TOL
        tab = orc.gaussian_table_west(co, g["f"][i], terms, b1)
        for j, n in enumerate(("b3", "cct", "cr1", "cr2", "cr4", "cr6", "ce0", "cemono", "cequad")):
            assert rowmax_rel(tab[n], g["table"][i][j]) <= TOL, n


@pytest.fixture(scope="module")
def golden3():
    return dict(np.load(os.path.join(GOLDEN, "config3_like.npz")))


def test_config3_marginalisation_from_reference_vectors(golden3):
    """marginal.py:79-140 restated: the reference's own PNG / PG -> its logp, fullchi2 and best fit, all 32 points"""
    g = golden3
    scales = json.loads(str(g["gauss_scales"]))
    names = [str(n) for n in g["gaussian_names"]]
    sig = np.diag([1.0 / scales[n.rsplit("_", 1)[1]] ** 2 for n in names])
    for tag, kw in (("LEX_NGC", dict(jeffreys=True)), ("LEX_NGC_gauss", dict(sigma_inv=sig))):
        for i in range(g[tag + ".logp"].size):
            lp, full, best = orc.marginalized_logp(g[tag + ".PNG"][i], g[tag + ".PG"][i], g["data_vector"], g["invcov"],
                                                   return_bestfit=True, **kw)
            assert lp == pytest.approx(g[tag + ".logp"][i], rel=1e-10)
            assert full == pytest.approx(g[tag + ".fullchi2"][i], rel=1e-8)
            np.testing.assert_allclose(best, g[tag + ".best"][i], rtol=1e-6, atol=1e-9)


def test_config3_one_point_full_chain(golden3):
    """the oracle's own three-tracer chain (its own window build, binning, chained transform, reductions) against the
    reference's Cobaya products for one point"""
    g = golden3
    fx = np.load(os.path.join(os.path.dirname(GOLDEN), "..", "eftpipe_b200", "data", "dr16_ngc.npz"))
    i = 1
    cfgs = {"LRG_NGC": dict(z=0.696, nd=4.5e-5, win="win_LRG", data="NGC_LRG_P", kmin=0.02, chained=False),
            "ELG_NGC": dict(z=0.849, nd=2.3e-4, win="win_ELG", data="NGC_ELG_Q", kmin=0.03, chained=True),
            "X_NGC": dict(z=0.763, win="win_X", data="NGC_X_P", kmin=0.02, chained=False)}
    b1 = {"LRG_NGC": g["pt.LRG_NGC_b1"][i], "ELG_NGC": g["pt.ELG_NGC_b1"][i]}
    b2 = {t: g[f"pt.{t}_c2"][i] / np.sqrt(2.0) for t in b1}
    png = []
    for t, c in cfgs.items():
        if t == "X_NGC":
            co = orc.Common(Nl=3, kmA=0.7, krA=0.25, ndA=4.5e-5, kmB=0.7, krB=0.25, ndB=2.3e-4)
        else:
            co = orc.Common(Nl=3, kmA=0.7, krA=0.25, ndA=c["nd"])
        nl, rs = orc.NonLinear(co), orc.Resum(co)
        ap = orc.APeffect(co, Om_AP=0.307115, z_AP=c["z"], APst=True)
        sQ = fx[c["win"]]
        Wal, p = orc.compute_Wal(sQ, co, Na=3, Nl=3, accboost=4)
        Waldk = orc.mask_and_measure(Wal, p, co.k, windowk=0.1)
        kall = fx[c["data"]][:, 0]
        kout = kall[(kall >= c["kmin"]) & (kall <= 0.20)]
        b = orc.Bird(co, np.logspace(-5, 0, 200), g[t + ".pkh"][i], g[t + ".f"][i], g[t + ".DA"][i], g[t + ".H"][i], c["z"])
        nl.PsCf(b)
        orc.set_PsCfl(b)
        rs.Ps(b)
        ap.AP(b)
        orc.apply_window(b, Waldk, p, window_st=True)
        terms = orc.Binning(kout, co).transform(orc.bird_terms(b))
        if c["chained"]:
            terms = orc.chained_transform(terms, co.Nl)
            co.No = 2
        zero = [0.0] * 4
        if t == "X_NGC":
            red = orc.reduce_Plk(co, b.f, terms, [b1["LRG_NGC"], b2["LRG_NGC"], 0.0, b2["LRG_NGC"]] + zero[:3],
                                 [b1["ELG_NGC"], b2["ELG_NGC"], 0.0, b2["ELG_NGC"]] + zero[:3])
        else:
            red = orc.reduce_Plk(co, b.f, terms, [b1[t], b2[t], 0.0, b2[t]] + zero[:3])
        assert rowmax_rel(red, g[t + ".Plk"][i]) <= 1e-9, t
        png.append(red.reshape(-1))
    # LRG / X: all three multipoles on 18 bins; ELG: chained l = 0, 2 on 17 bins
    assert rowmax_rel(np.concatenate(png), g["LEX_NGC.PNG"][i]) <= 1e-9


def test_every_product_of_one_tracer_full_chain():
    """the oracle's own chain (window build, binning, chained transform, reduction, Gaussian tables, interpolator) against the
    products the unmodified reference's theory component served in one evaluation (tests/golden/products_elg.npz, modelled on
    the reference's tests/regression/test_eftlss.py)"""
    g = dict(np.load(os.path.join(GOLDEN, "products_elg.npz")))
    fx = np.load(os.path.join(os.path.dirname(GOLDEN), "..", "eftpipe_b200", "data", "dr16_ngc.npz"))
    i, T = 2, "ELG_NGC"
    pt = {k[len("pt." + T) + 1:]: v[i] for k, v in g.items() if k.startswith("pt." + T)}
    co = orc.Common(Nl=3, kmA=0.7, krA=0.25, ndA=2.3e-4)
    nl, rs = orc.NonLinear(co), orc.Resum(co)
    ap = orc.APeffect(co, Om_AP=0.307115, z_AP=0.849, APst=True)
    Wal, p = orc.compute_Wal(fx["win_ELG"], co, Na=3, Nl=3, accboost=4)
    Waldk = orc.mask_and_measure(Wal, p, co.k, windowk=0.1)
    b = orc.Bird(co, np.logspace(-5, 0, 200), g["tab.pkh"][i], g["tab.f"][i], g["tab.DA"][i], g["tab.H"][i], 0.849)
    nl.PsCf(b)
    orc.set_PsCfl(b)
    rs.Ps(b)
    ap.AP(b)
    orc.apply_window(b, Waldk, p, window_st=True)
    plain = orc.bird_terms(b)
    binned = orc.Binning(g["kout"], co).transform(plain)
    b2 = pt["c2"] / np.sqrt(2.0)
    bs = [pt["b1"], b2, pt["b3"], b2, pt["cct"], pt["cr1"], pt["cr2"]]
    es = [pt["ce0"], 0.0, pt["cequad"]]
    for ch in (False, True):
        for bn in (False, True):
            tag = f"c{int(ch)}b{int(bn)}"
            terms = binned if bn else plain
            cc = orc.Common(Nl=3, kmA=0.7, krA=0.25, ndA=2.3e-4)
            if ch:
                terms = orc.chained_transform(terms, co.Nl)
                cc.No = 2
            red = orc.reduce_Plk(cc, b.f, terms, bs, es=es)
            assert rowmax_rel(red, g[f"grid_{tag}.P"][i]) <= 1e-9, tag
            tab = orc.gaussian_table_west(cc, b.f, terms, pt["b1"])
            for n in ("b3", "cct", "cr1", "cr2", "ce0", "cemono", "cequad"):
                assert rowmax_rel(tab[n], g[f"gauss_{tag}.{T}_{n}"][i]) <= 1e-9, (tag, n)
    # the interpolators are built from the un-binned products (theory.py:862-871)
    red = orc.reduce_Plk(co, b.f, plain, bs, es=es)
    assert rowmax_rel(orc.plk_interpolator(co.k, red)(g["kout"]), g["plk"][i]) <= 1e-9
    cc = orc.Common(Nl=3, kmA=0.7, krA=0.25, ndA=2.3e-4)
    cc.No = 2
    redc = orc.reduce_Plk(cc, b.f, orc.chained_transform(plain, co.Nl), bs, es=es)
    assert rowmax_rel(orc.plk_interpolator(co.k, redc)(g["kout"]), g["plk_chained"][i]) <= 1e-9
