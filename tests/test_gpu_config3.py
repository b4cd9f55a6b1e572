"""GPU against LIVE-REFERENCE goldens (tests/golden/make_golden_config3.py), not only the oracle:
  * bias reduction + Gaussian tables: West-coast auto / cross, East-coast, each with and without NNLO (reduce_kat.npz);
  * with_NNLO=True through the fused pipeline, the projection, the reduction with cr4 / cr6 (nnlo_chain.npz);
  * BASELINE config 3 - the DR16 NGC LRG x ELG x cross likelihood at B = 32: per-tracer multipoles, PNG, PG, logp under
    the Jeffreys and the Gaussian-prior yaml, fullchi2, best fit - as the reference's own Cobaya components
    (theory.py, likelihood.py, marginal.py) computed them (config3_like.npz)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, rowmax_rel

pytestmark = pytest.mark.gpu
TOL = 1e-8  # north star: multipoles to 1e-8 relative (row-max scaled), chi^2 to 1e-6
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "eftpipe_b200", "data", "dr16_ngc.npz")


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def kat():
    return dict(np.load(os.path.join(GOLDEN, "reduce_kat.npz")))


def _view(kat, co, nnlo):
    """the fixture's term arrays as a device birdlike (batch-minor (No, nk, nterm, Bp))"""
    import torch

    from eftpipe_b200.transformer import PlainBird

    parts = [kat["T." + n] for n in ("P11l", "Pctl", "Ploopl", "Pstl")] + ([kat["T.PctNNLOl"]] if nnlo else [])
    T = np.concatenate(parts, axis=2)  # (B, No, nterm, nk)
    B = T.shape[0]
    Bp = (B + 31) // 32 * 32
    bm = np.zeros(T.shape[1:2] + (T.shape[3], T.shape[2], Bp))
    bm[..., :B] = T.transpose(1, 3, 2, 0)
    bm[..., B:] = bm[..., B - 1 : B]
    f = np.concatenate([kat["f"], np.full(Bp - B, kat["f"][-1])])
    return PlainBird(torch.as_tensor(kat["f"], device="cuda"), co, torch.as_tensor(bm, device="cuda"), kat["T.Picc"][0] * 0.0, B, False,
                     torch.as_tensor(f, device="cuda"))


@pytest.mark.parametrize("tag", ["west_auto", "west_nnlo", "west_cross", "west_cross_nnlo", "east", "east_nnlo"])
def test_reduction_against_reference(kat, tag):
    from eftpipe_b200 import parambasis, pybird

    nnlo = tag.endswith("nnlo")
    scal = json.loads(str(kat["scales"]))
    if tag.startswith("east"):
        basis, form = parambasis.EastCoastBasis(prefix="e_"), "eastcoast"
    elif "cross" in tag:
        basis, form = parambasis.WestCoastBasis(prefix="X_", cross_prefix=["A_", "B_"]), "westcoast"
    else:
        basis, form = parambasis.WestCoastBasis(prefix="w_"), "westcoast"
    co = pybird.Common(Nl=3, counterform=form, with_NNLO=nnlo, **scal)
    view = _view(kat, co, nnlo)
    params = {k: np.array(v) for k, v in json.loads(str(kat[tag + ".params"])).items()}
    # Picc is a per-point array in this fixture (the product carries one constant per plan): compare without it
    ref = kat[tag + ".reduced"] - kat["T.Picc"]
    got = _np(basis.reduce_Plk(view, params).sum())
    assert got.shape == ref.shape
    assert rowmax_rel(got, ref) <= TOL
    table = basis.reduce_Plk_gaussian_table(view, params)
    names = [str(n) for n in kat[tag + ".table_names"]]
    assert set(names) == set(table), (names, list(table))
    for j, n in enumerate(names):
        assert rowmax_rel(_np(table[n]), kat[tag + ".table"][j]) <= TOL, n


def test_function_form_reduce_honours_cnnlo(kat):
    """parambasis.reduce_Plk(bird, bsA, bsB, es, cnnloA) (parambasis.py:42-136): round 1 ignored cnnloA"""
    from eftpipe_b200 import parambasis, pybird

    co = pybird.Common(Nl=3, counterform="westcoast", with_NNLO=True, **json.loads(str(kat["scales"])))
    view = _view(kat, co, True)
    p = {k: np.array(v) for k, v in json.loads(str(kat["west_nnlo.params"])).items()}
    g = lambda n: p["w_" + n]
    got = _np(parambasis.reduce_Plk(view, [g(n) for n in ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2")],
                                    es=[g(n) for n in ("ce0", "cemono", "cequad")], cnnloA=[g("cr4"), g("cr6")]).sum())
    assert rowmax_rel(got, kat["west_nnlo.reduced"] - kat["T.Picc"]) <= TOL
    none = _np(parambasis.reduce_Plk(view, [g(n) for n in ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2")],
                                     es=[g(n) for n in ("ce0", "cemono", "cequad")]).sum())
    assert rowmax_rel(none, kat["west_nnlo.reduced"] - kat["T.Picc"]) > 1e-4  # the counterterm is not negligible here


def test_nnlo_chain_against_reference():
    from eftpipe_b200 import engine, parambasis, plan as P, pybird, synthetic
    from eftpipe_b200.transformer import PlainBird

    g = dict(np.load(os.path.join(GOLDEN, "nnlo_chain.npz")))
    g2 = np.load(os.path.join(GOLDEN, "config2_chain.npz"))
    apk = json.loads(str(g["ap"]))
    ap = dict(DA=synthetic.angular_distance(apk["Om_AP"], apk["z_AP"]), H=synthetic.hubble(apk["Om_AP"], apk["z_AP"]), APst=True)
    # after AP (no projection)
    dp = engine.DevicePlan(P.build_tracer_plan(Nl=3, with_NNLO=True, ap=ap))
    T, _ = dp.eval_terms(g["plin"], g["f"], g["DA"], g["H"])
    T = _np(T)  # (B, Nl, 27, Nk)
    for n, sl in (("P11l", slice(0, 3)), ("Pctl", slice(3, 9)), ("Ploopl", slice(9, 21)), ("Pstl", slice(21, 24)), ("PctNNLOl", slice(24, 27))):
        assert rowmax_rel(T[:, :, sl], g["ap_" + n]) <= TOL, n
    # window + binning, then the reduction with cr4 / cr6 and the Gaussian table
    grid = P.GridConfig(Nl=3, with_NNLO=True)
    binm, keff, _, _ = P.binning_matrix(grid.k, g["kout"])
    proj = P.compose_projection(grid, window=g2["Weff_LRG"], icc=None, binning=binm)
    dp = engine.DevicePlan(P.build_tracer_plan(Nl=3, with_NNLO=True, ap=ap, projection=proj))
    pm, bm = dp.eval_terms(g["plin"], g["f"], g["DA"], g["H"], want_bm=True)
    pm = _np(pm)
    for n, sl in (("P11l", slice(0, 3)), ("Pctl", slice(3, 9)), ("Ploopl", slice(9, 21)), ("Pstl", slice(21, 24)), ("PctNNLOl", slice(24, 27))):
        assert rowmax_rel(pm[:, :, sl], g["bin_" + n]) <= TOL, n
    co = pybird.Common(**json.loads(str(g["common"])))
    nk = g["kout"].size
    view = PlainBird(None, co, bm.reshape(3, nk, 27, -1), np.zeros((3, nk)), 2, False, dp.to_batch_minor(g["f"])[0])
    names = ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2", "ce0", "cemono", "cequad", "cr4", "cr6")
    params = {n: g["params"][:, j] for j, n in enumerate(names)}
    basis = parambasis.WestCoastBasis(prefix="")
    assert rowmax_rel(_np(basis.reduce_Plk(view, params).sum()), g["reduced"]) <= TOL
    table = basis.reduce_Plk_gaussian_table(view, params)
    for j, n in enumerate(("b3", "cct", "cr1", "cr2", "cr4", "cr6", "ce0", "cemono", "cequad")):
        assert rowmax_rel(_np(table[n]), g["table"][:, j]) <= TOL, n


# ------------------------------------------------------------------------------------------ config 3
@pytest.fixture(scope="module")
def golden3():
    return dict(np.load(os.path.join(GOLDEN, "config3_like.npz")))


def _config3(dr16, marg, jeffreys):
    from eftpipe_b200 import likelihood, theory

    ap = dict(Om_AP=0.307115, rdrag_AP=147.66, h_AP=0.6777, APst=True)
    tracers = {
        "LRG_NGC": dict(prefix="LRG_NGC_", z=0.696, nd=4.5e-5, window=dict(window_configspace_array=dr16["win_LRG"])),
        "ELG_NGC": dict(prefix="ELG_NGC_", z=0.849, nd=2.3e-4, window=dict(window_configspace_array=dr16["win_ELG"])),
        "X_NGC": dict(prefix="X_NGC_", z=0.763, cross=["LRG_NGC", "ELG_NGC"], window=dict(window_configspace_array=dr16["win_X"])),
        "default": dict(km=0.7, kr=0.25, with_IRresum=True, with_APeffect=True, with_window=True, APeffect=ap,
                        window=dict(accboost=4, windowk=0.1)),
    }
    like = likelihood.EFTLike(
        tracers=["LRG_NGC", "ELG_NGC", "X_NGC"], chained=[False, True, False],
        data={"LRG_NGC": dict(table=dr16["NGC_LRG_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20),
              "ELG_NGC": dict(table=dr16["NGC_ELG_Q"], ls=[0, 2], kmin=0.03, kmax=0.20, symbol="Q"),
              "X_NGC": dict(table=dr16["NGC_X_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20)},
        cov=dict(matrix=dr16["cov_NGC_L024E02X024_PQP"], Nreal=1000), with_binning=True, jeffreys=jeffreys, marg=marg)
    th = theory.EFTLSS(tracers).must_provide(like.get_requirements()).initialize()
    like.initialize_with_provider(th)
    return th, like


@pytest.fixture(scope="module")
def config3_jeffreys():
    import refdriver

    return _config3(dict(np.load(DATA)), refdriver.marg_block(), True)


def _inputs(g):
    cosmo = {t: dict(pkh=g[t + ".pkh"], f=g[t + ".f"], DA=g[t + ".DA"], H=g[t + ".H"], rdrag=g[t + ".rdrag"], h=g[t + ".h"])
             for t in ("LRG_NGC", "ELG_NGC", "X_NGC")}
    params = {}
    for pre in ("LRG_NGC_", "ELG_NGC_"):
        params[pre + "b1"] = g["pt." + pre + "b1"]
        params[pre + "b2"] = params[pre + "b4"] = g["pt." + pre + "c2"] / np.sqrt(2.0)
    return cosmo, params


def test_config3_host_side_matches_reference(config3_jeffreys, golden3):
    """data vector, masked + Hartlap-corrected inverse covariance (likelihood.py:337-363) and the marginalised-parameter order"""
    th, like = config3_jeffreys
    assert like.ndata == 142
    # the reference reads its text tables with pandas, whose default float parser is not correctly rounded (1 ulp)
    np.testing.assert_allclose(like.data_vector, golden3["data_vector"], rtol=1e-15, atol=0)
    # ... and the inverse amplifies those 1-ulp input differences by the condition number of the 142 x 142 covariance
    np.testing.assert_allclose(like.invcov, golden3["invcov"], rtol=1e-8, atol=1e-12 * np.abs(golden3["invcov"]).max())
    assert list(like.gaussian_names) == [str(n) for n in golden3["gaussian_names"]]
    for t in ("LRG_NGC", "ELG_NGC", "X_NGC"):
        np.testing.assert_allclose(th.info[t]["kout"], golden3[t + ".keff"], rtol=1e-13)
        assert th.info[t]["ls"][: len(golden3[t + ".ls"])] == list(golden3[t + ".ls"])


def test_config3_jeffreys_against_reference(config3_jeffreys, golden3):
    th, like = config3_jeffreys
    g = golden3
    cosmo, params = _inputs(g)
    th.calculate(cosmo)
    for t in ("LRG_NGC", "ELG_NGC", "X_NGC"):
        ls, k, plk = th.get_nonlinear_Plk_grid(t, params)
        plk = _np(plk)
        assert plk.shape == g[t + ".Plk"].shape, t
        assert rowmax_rel(plk, g[t + ".Plk"]) <= TOL, t
    png, pg = like.PNG_PG(params)
    assert rowmax_rel(_np(png), g["LEX_NGC.PNG"]) <= TOL
    assert rowmax_rel(_np(pg), g["LEX_NGC.PG"]) <= TOL
    res = like.calculate(params, want_bestfit=True)
    assert not _np(res["status"]).any()
    np.testing.assert_allclose(_np(res["logp"]), g["LEX_NGC.logp"], rtol=1e-6, atol=0)  # north star: chi^2 to 1e-6
    assert np.max(np.abs(_np(res["logp"]) / g["LEX_NGC.logp"] - 1)) <= 1e-9                # ... measured: far tighter
    np.testing.assert_allclose(_np(res["eftlike_fullchi2"]), g["LEX_NGC.fullchi2"], rtol=1e-6)
    best = np.stack([_np(res["bestfit"]["marg_" + n]) for n in like.gaussian_names], axis=1)
    np.testing.assert_allclose(best, g["LEX_NGC.best"], rtol=1e-5, atol=1e-7)
    # derived parameters of theory.py:620-648 against the reference's own APeffect.get_alperp_alpara is covered in
    # test_gpu_mirror; here: the same numbers again through one CUDA graph replay
    from eftpipe_b200.engine import capture_graph

    import torch

    dcosmo = {t: {k: torch.as_tensor(np.ascontiguousarray(v), device="cuda") for k, v in c.items() if k in ("pkh", "f", "DA", "H")}
              for t, c in cosmo.items()}
    dparams = {k: torch.as_tensor(v, device="cuda") for k, v in params.items()}

    def step():
        th.calculate(dcosmo)
        return like.calculate(dparams)["logp"]

    graph, out = capture_graph(step)
    graph.replay()
    torch.cuda.synchronize()
    np.testing.assert_allclose(_np(out), g["LEX_NGC.logp"], rtol=1e-9)


def test_config3_gaussian_priors_against_reference(golden3):
    import refdriver

    g = golden3
    th, like = _config3(dict(np.load(DATA)), refdriver.marg_block(refdriver.GAUSS_SCALES), False)
    cosmo, params = _inputs(g)
    th.calculate(cosmo)
    res = like.calculate(params, want_bestfit=True)
    np.testing.assert_allclose(_np(res["logp"]), g["LEX_NGC_gauss.logp"], rtol=1e-6, atol=0)
    assert np.max(np.abs(_np(res["logp"]) / g["LEX_NGC_gauss.logp"] - 1)) <= 1e-9
    np.testing.assert_allclose(_np(res["eftlike_fullchi2"]), g["LEX_NGC_gauss.fullchi2"], rtol=1e-6)
    best = np.stack([_np(res["bestfit"]["marg_" + n]) for n in like.gaussian_names], axis=1)
    np.testing.assert_allclose(best, g["LEX_NGC_gauss.best"], rtol=1e-5, atol=1e-7)


def test_wide_marginalisation_and_two_operand_form(golden3, monkeypatch):
    """17 marginalised parameters (the 32-column variant of the Gram / warp-Cholesky kernel) and the two-operand
    fallback (C^-1 V instead of the Cholesky factor of C^-1): both against marginal.py restated on the kernel's own
    PNG / PG vectors, which the tests above pin to the reference."""
    import pybird_oracle as orc
    import refdriver

    g = golden3
    marg = refdriver.marg_block(refdriver.GAUSS_SCALES)
    for pre in ("LRG_NGC_", "ELG_NGC_"):
        marg[pre] = dict(marg[pre], cemono={"scale": 2.0})
    marg["X_NGC_cemono"] = {"scale": 2.0}
    th, like = _config3(dict(np.load(DATA)), marg, False)
    assert len(like.gaussian_names) == 17
    cosmo, params = _inputs(g)
    th.calculate(cosmo)
    png, pg = like.PNG_PG(params)
    png, pg = _np(png), _np(pg)
    sig = like.sigma_inv

    def check():
        res = like.calculate(params, want_bestfit=True)
        lp, full = _np(res["logp"]), _np(res["eftlike_fullchi2"])
        best = np.stack([_np(res["bestfit"]["marg_" + n]) for n in like.gaussian_names], axis=1)
        for i in range(0, png.shape[0], 5):
            ref_lp, ref_full, ref_best = orc.marginalized_logp(png[i], pg[i], like.data_vector, like.invcov, sigma_inv=sig, return_bestfit=True)
            assert lp[i] == pytest.approx(ref_lp, rel=1e-9)
            assert full[i] == pytest.approx(ref_full, rel=1e-7)
            np.testing.assert_allclose(best[i], ref_best, rtol=1e-5, atol=1e-7)
        return lp

    lp_one = check()
    # the environment knob is read once per process by the library: exercise the two-operand form through a fresh process
    import subprocess
    import sys

    code = ("import os, sys, numpy as np; sys.path[:0] = [%r, %r, %r]; os.environ['EFTB_LIKE_TWO_OPERAND'] = '1'\n"
            "import test_gpu_config3 as t, refdriver\n"
            "g = dict(np.load(os.path.join(t.GOLDEN, 'config3_like.npz')))\n"
            "th, like = t._config3(dict(np.load(t.DATA)), refdriver.marg_block(), True)\n"
            "cosmo, params = t._inputs(g); th.calculate(cosmo); res = like.calculate(params)\n"
            "np.testing.assert_allclose(res['logp'].cpu().numpy(), g['LEX_NGC.logp'], rtol=1e-9); print('two-operand ok')\n"
            % (os.path.dirname(os.path.abspath(__file__)), os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"),
               os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "two-operand ok" in out.stdout, out.stderr[-2000:]
    assert np.isfinite(lp_one).all()


def test_nnlo_likelihood_through_theory_and_likelihood():
    """with_NNLO=True end to end (round 1 dropped the NNLO counterterm silently at the last step): EFTLSS tracer with
    `with_NNLO`, marginalisation over cr4 / cr6 next to the usual parameters, against marginal.py restated on vectors built
    from the REFERENCE's binned NNLO terms (tests/golden/nnlo_chain.npz) by the reference-pinned oracle reductions."""
    import pybird_oracle as orc
    from eftpipe_b200 import likelihood, theory

    g = dict(np.load(os.path.join(GOLDEN, "nnlo_chain.npz")))
    g2 = np.load(os.path.join(GOLDEN, "config2_chain.npz"))
    apk = json.loads(str(g["ap"]))
    kout = g["kout"]
    nk = kout.size
    tracers = {"LRG": dict(prefix="", z=float(g["z"]), nd=4.5e-5, km=0.7, kr=0.25, with_NNLO=True, with_IRresum=True, with_APeffect=True,
                           APeffect=dict(rdrag_AP=147.66, h_AP=0.6777, **apk), with_window="helpers.MatrixWindow",
                           window=dict(matrix=g2["Weff_LRG"]))}
    rng = np.random.default_rng(8)
    table = np.column_stack([kout] + [1e4 * rng.standard_normal(nk) for _ in range(3)])
    cov = np.diag(rng.uniform(1e5, 1e6, 3 * nk))
    names = ("b3", "cct", "cr1", "cr2", "cr4", "cr6", "ce0", "cequad")
    like = likelihood.EFTLike(tracers=["LRG"], data=dict(table=table, ls=[0, 2, 4], kmin=0.0, kmax=1.0), cov=dict(matrix=cov),
                              with_binning=True, jeffreys=False, marg={n: {"scale": 3.0} for n in names})
    th = theory.EFTLSS(tracers).must_provide(like.get_requirements()).initialize()
    like.initialize_with_provider(th)
    assert sorted(like.gaussian_names) == sorted(names)
    names = list(like.gaussian_names)  # the order of marginal.py:198-232 (the basis' parameter order)
    th.calculate({"LRG": dict(pkh=g["plin"], f=g["f"], DA=g["DA"], H=g["H"])})
    p = g["params"]  # b1, b2, b3, b4, cct, cr1, cr2, ce0, cemono, cequad, cr4, cr6
    params = {"b1": p[:, 0], "b2": p[:, 1], "b4": p[:, 3], "cemono": p[:, 8]}
    res = like.calculate(params, want_bestfit=True)
    co = orc.Common(**json.loads(str(g["common"])))
    sig = np.eye(len(names)) / 9.0
    for i in range(p.shape[0]):
        terms = {n: g["bin_" + n][i] for n in ("P11l", "Pctl", "Ploopl", "Pstl", "PctNNLOl", "Picc")}
        png = orc.reduce_Plk(co, g["f"][i], terms, [p[i, 0], p[i, 1], 0.0, p[i, 3], 0.0, 0.0, 0.0], es=[0.0, p[i, 8], 0.0],
                             cnnlo=(0.0, 0.0)).reshape(-1)
        tab = orc.gaussian_table_west(co, g["f"][i], terms, p[i, 0])
        pg = np.array([tab[n].reshape(-1) for n in names])
        assert np.abs(pg[names.index("cr4")]).max() > 0 and np.abs(pg[names.index("cr6")]).max() > 0  # the NNLO rows are really there
        ref, full, best = orc.marginalized_logp(png, pg, like.data_vector, like.invcov, sigma_inv=sig, return_bestfit=True)
        assert _np(res["logp"])[i] == pytest.approx(ref, rel=1e-8)
        got_best = np.array([_np(res["bestfit"]["marg_" + n])[i] for n in names])
        np.testing.assert_allclose(got_best, best, rtol=1e-5, atol=1e-8)


# ------------------------------------------------------------------------------------------ config 4 shard (full size)
def test_config4_shard_size_properties(config3_jeffreys, golden3):
    """BASELINE config 4 shards 65 536 points over 8 GPUs: 8192 points per GPU through the config-3 pipeline.  At that size
    the reference cannot be run as the checker, so the full shard is tied to it through size-independent properties:
      (1) the 32 reference-evaluated golden points, scattered through a batch of 8192 distinct cosmologies, still match;
      (2) evaluating a permuted batch permutes the results bit for bit (no point sees its neighbours or its position);
      (3) a point evaluated in a batch of 96 agrees with the same point in the batch of 8192 (tile shapes, grid sizes and the
          CTA interleave of the resummation kernel change with the batch size: 1e-12, measured bit-identical or 1 ulp);
      (4) the same call twice is bit-identical (no atomics on shared accumulators, no uninitialised workspace)."""
    import torch

    from eftpipe_b200 import synthetic

    th, like = config3_jeffreys
    g = golden3
    B, ng = 8192, g["LEX_NGC.logp"].size
    rng = np.random.default_rng(4)
    pos = np.sort(rng.choice(B, ng, replace=False))  # where the golden points sit
    cosmo, params = {}, {}
    for t, z in (("LRG_NGC", 0.696), ("ELG_NGC", 0.849), ("X_NGC", 0.763)):
        b = synthetic.make_batch_fast(B, z, seed=4040)
        c = dict(pkh=b.plin, f=b.f, DA=b.DA, H=b.H)
        for k in c:
            c[k][pos] = g[f"{t}.{k}"]
        cosmo[t] = c
    for pre, b1 in (("LRG_NGC_", 2.1), ("ELG_NGC_", 1.4)):
        v1 = b1 + 0.05 * rng.standard_normal(B)
        c2 = 0.7 + 0.1 * rng.standard_normal(B)
        v1[pos], c2[pos] = g["pt." + pre + "b1"], g["pt." + pre + "c2"]
        params[pre + "b1"] = v1
        params[pre + "b2"] = params[pre + "b4"] = c2 / np.sqrt(2.0)

    def run(idx):
        th.calculate({t: {k: v[idx] for k, v in c.items()} for t, c in cosmo.items()})
        res = like.calculate({k: v[idx] for k, v in params.items()})
        torch.cuda.synchronize()
        assert not _np(res["status"]).any()
        return _np(res["logp"]).copy(), _np(res["eftlike_fullchi2"]).copy()

    every = np.arange(B)
    logp, chi2 = run(every)
    assert np.isfinite(logp).all() and np.isfinite(chi2).all()
    np.testing.assert_allclose(logp[pos], g["LEX_NGC.logp"], rtol=1e-9)           # (1)
    np.testing.assert_allclose(chi2[pos], g["LEX_NGC.fullchi2"], rtol=1e-9)
    perm = rng.permutation(B)
    logp_p, chi2_p = run(perm)                                                       # (2)
    assert np.array_equal(logp_p, logp[perm]) and np.array_equal(chi2_p, chi2[perm])
    sub = np.sort(rng.choice(B, 96, replace=False))
    logp_s, chi2_s = run(sub)                                                        # (3)
    np.testing.assert_allclose(logp_s, logp[sub], rtol=1e-12)
    np.testing.assert_allclose(chi2_s, chi2[sub], rtol=1e-12)
    logp_2, chi2_2 = run(every)                                                      # (4)
    assert np.array_equal(logp_2, logp) and np.array_equal(chi2_2, chi2)
