"""CPU: the oracle (numpy restatement) against the golden vectors produced by the live reference."""
import json

import numpy as np
import pytest

import pybird_oracle as orc
from conftest import rowmax_rel
from eftpipe_b200 import synthetic

TOL = 5e-11


def test_reference_known_answers(fftlog_kat):
    # the only literal known-answer values of the reference's own suite (tests/test_pybird.py:8-9)
    assert orc.hubble(0.2, 1.0) == pytest.approx(1.549193338482967, rel=1e-15)
    assert orc.dafunc(0.2, 1.0) == pytest.approx(0.4117451980802465, rel=1e-12)
    assert float(fftlog_kat["hubble"]) == pytest.approx(1.549193338482967, rel=1e-15)
    assert synthetic.hubble(0.2, 1.0) == pytest.approx(1.549193338482967, rel=1e-15)
    assert synthetic.angular_distance(0.2, 1.0) == pytest.approx(0.4117451980802465, rel=1e-12)


def test_fftlog_vectorised_equals_looped(fftlog_kat):
    # the reference's property test (tests/compare/test_fftlog.py:5-23), plus its stored output
    g = orc.LogGrid(Nmax=256, xmin=1e-5, xmax=10, bias=-0.3)
    x, rows = fftlog_kat["x"], fftlog_kat["rows"]
    vec = orc.fftlog_coef(g, x, rows, extrap="padding", window=0.3)
    loop = np.array([orc.fftlog_coef(g, x, r, extrap="padding", window=0.3) for r in rows])
    np.testing.assert_allclose(vec, loop, rtol=1e-6, atol=0)
    assert np.abs(vec - fftlog_kat["coef"]).max() <= 1e-13 * np.abs(fftlog_kat["coef"]).max()


def test_fftlog_rejects_odd_nmax():
    with pytest.raises(ValueError):
        orc.LogGrid(Nmax=255, xmin=1e-5, xmax=10, bias=-0.3)


@pytest.fixture(scope="module")
def oracle_chain(golden2):
    g = golden2
    kw = json.loads(str(g["common"]))
    co = orc.Common(**kw)
    nl, rs = orc.NonLinear(co), orc.Resum(co)
    apk = json.loads(str(g["ap"]))
    ap = orc.APeffect(co, **apk)
    birds = []
    for i in range(g["plin"].shape[0]):
        b = orc.Bird(co, g["kin"], g["plin"][i], g["f"][i], g["DA"][i], g["H"][i], float(g["z"]))
        nl.PsCf(b)
        stage = dict(coef=b.coef, P11=b.P11, P22=b.P22, P13=b.P13, C11=b.C11, Cct=b.Cct, C22=b.C22, C13=b.C13)
        orc.set_PsCfl(b)
        stage.update(pre_P11l=b.P11l, pre_Pctl=b.Pctl, pre_Ploopl=b.Ploopl, pre_Cloopl=b.Cloopl, pre_Pstl=b.Pstl)
        rs.Ps(b)
        stage.update(X=b.X, Y=b.Y, res_P11l=b.P11l, res_Pctl=b.Pctl, res_Ploopl=b.Ploopl)
        ap.AP(b)
        stage.update(ap_P11l=b.P11l, ap_Pctl=b.Pctl, ap_Ploopl=b.Ploopl, ap_Pstl=b.Pstl)
        birds.append(stage)
    return birds


@pytest.mark.parametrize("key", ["coef", "P11", "P22", "P13", "C11", "Cct", "C22", "C13", "pre_P11l", "pre_Pctl",
                                 "pre_Ploopl", "pre_Cloopl", "pre_Pstl", "X", "Y", "res_P11l", "res_Pctl",
                                 "res_Ploopl", "ap_P11l", "ap_Pctl", "ap_Ploopl", "ap_Pstl"])
def test_oracle_stage_matches_reference(oracle_chain, golden2, key):
    for i, stage in enumerate(oracle_chain):
        ref = golden2[key][i]
        got = stage[key]
        if np.iscomplexobj(ref):
            assert np.abs(got - ref).max() <= TOL * np.abs(ref).max()
        else:
            assert rowmax_rel(got, ref) <= TOL, (key, i)


def test_oracle_nl2(golden_nl2):
    g = golden_nl2
    co = orc.Common(**json.loads(str(g["common"])))
    nl, rs = orc.NonLinear(co), orc.Resum(co)
    for i in range(g["plin"].shape[0]):
        b = orc.Bird(co, g["kin"], g["plin"][i], g["f"][i])
        nl.PsCf(b)
        orc.set_PsCfl(b)
        rs.Ps(b)
        for key in ("P11l", "Pctl", "Ploopl"):
            assert rowmax_rel(getattr(b, key), g["res_" + key][i]) <= TOL


def test_oracle_marginalisation(golden2):
    g = golden2
    use = [0, 1, 2, 3, 4, 6]
    for i in range(g["plin"].shape[0]):
        PG = g["gaussian_table_binned"][i][use].reshape(len(use), -1)
        PNG = g["marg_PNG"][i]
        lp, full, best = orc.marginalized_logp(PNG, PG, g["lrg_data"], g["lrg_invcov"], jeffreys=True, return_bestfit=True)
        ref = g["marg_out"][i]
        assert lp == pytest.approx(ref[0], rel=1e-9)
        assert full == pytest.approx(ref[1], rel=1e-9)
        np.testing.assert_allclose(best, ref[2:8], rtol=1e-7)
        sig = np.diag(1.0 / np.array([4, 2, 4, 4, 2, 2], float) ** 2)
        lp2 = orc.marginalized_logp(PNG, PG, g["lrg_data"], g["lrg_invcov"], sigma_inv=sig)
        assert lp2 == pytest.approx(ref[8], rel=1e-9)


def test_oracle_rejects_non_pd():
    PG = np.zeros((2, 4))
    with pytest.raises(RuntimeError):
        orc.marginalized_logp(np.ones(4), PG, np.zeros(4), np.eye(4))


def test_unbinned_products_against_reference(interp_kat):
    """theory.py:75-106 PlkInterpolator and likelihood.py:510-513, restated in the oracle and as the fixed operators
    the CUDA path composes into its projection (plan.interp_matrices), against outputs of the live reference."""
    import pybird_oracle as orc
    from eftpipe_b200 import plan as P

    g = interp_kat
    k, kout = P.GridConfig(Nl=3).k, g["kout"]
    S_png, S_pg = P.interp_matrices(k, kout)
    assert np.abs(S_png - S_pg).max() > 1e-7  # the two interpolations really differ (inserted origin)
    for i in range(g["Plk"].shape[0]):
        png = orc.plk_interpolator(k, g["Plk"][i])(kout).reshape(-1)
        assert np.max(np.abs(png - g["PNG_interp"][i])) <= 1e-12 * np.abs(g["PNG_interp"][i]).max()
        assert np.max(np.abs((g["Plk"][i] @ S_png.T).reshape(-1) - g["PNG_interp"][i])) <= 1e-11 * np.abs(g["PNG_interp"][i]).max()
        # marginalised rows: rebuild them from the raw-grid rows with the second operator
        pg = (g["PG_raw"][i].reshape(-1, 3, k.size) @ S_pg.T).reshape(g["PG_interp"][i].shape)
        scale = np.abs(g["PG_interp"][i]).max(axis=-1, keepdims=True)
        assert np.max(np.abs(pg - g["PG_interp"][i]) / scale) <= 1e-11
        lp = orc.marginalized_logp(g["PNG_interp"][i], g["PG_interp"][i], g["data"], g["invcov"], jeffreys=True)
        assert abs(lp - g["logp_interp"][i]) <= 1e-10 * abs(g["logp_interp"][i])
