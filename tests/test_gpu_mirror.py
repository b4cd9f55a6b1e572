"""GPU: the reference-shaped host API (Common / Bird / NonLinear / Resum / APeffect / Window /
IntegralConstraint / Binning / Chained / WestCoastBasis / EFTLSS / EFTLike) driven exactly like the
reference's own call sequence (theory.py:557-609), against goldens and the oracle."""
import json
import os

import numpy as np
import pytest

from conftest import rowmax_rel

pytestmark = pytest.mark.gpu
TOL = 1e-8
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "eftpipe_b200", "data", "dr16_ngc.npz")


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def dr16():
    return dict(np.load(DATA))


@pytest.fixture(scope="module")
def chain(golden2, dr16):
    from eftpipe_b200 import binning, chained, icc, pybird, window

    g = golden2
    co = pybird.Common(**json.loads(str(g["common"])))
    nl = pybird.NonLinear(load=False, save=False, co=co)
    rs = pybird.Resum(co=co)
    ap = pybird.APeffect(co=co, **json.loads(str(g["ap"])))
    win = window.Window(window_configspace_array=dr16["win_LRG"], co=co, accboost=4, windowk=0.1)
    ic = icc.IntegralConstraint(Pshot=float(g["Pshot"]), PSN=g["PSN"], Wal=0.05 * win.Wal, co=co, accboost=4, windowk=0.1)
    win.icc = ic
    bird = pybird.Bird(g["kin"], g["plin"], g["f"], g["DA"], g["H"], float(g["z"]), co=co)
    out = {}
    nl.PsCf(bird)
    out.update(P11=_np(bird.P11), P22=_np(bird.P22), P13=_np(bird.P13), C11=_np(bird.C11), Cct=_np(bird.Cct),
               C22=_np(bird.C22), C13=_np(bird.C13))
    bird.setPsCfl()
    out.update({"pre_" + n: _np(getattr(bird, n)) for n in ("P11l", "Pctl", "Ploopl", "Pstl", "Cloopl")})
    rs.Ps(bird)
    out.update({"res_" + n: _np(getattr(bird, n)) for n in ("P11l", "Pctl", "Ploopl")})
    ap.AP(bird)
    out.update({"ap_" + n: _np(getattr(bird, n)) for n in ("P11l", "Pctl", "Ploopl", "Pstl")})
    win.Window(bird)
    out.update({"win_" + n: _np(getattr(bird, n)) for n in ("P11l", "Pctl", "Ploopl", "Pstl", "Picc")})
    binned = binning.Binning(g["kout"], co=co).transform(bird)
    out.update({"bin_" + n: _np(getattr(binned, n)) for n in ("P11l", "Pctl", "Ploopl", "Pstl", "Picc")})
    ch = chained.Chained().transform(binned)
    out.update({"chn_" + n: _np(getattr(ch, n)) for n in ("P11l", "Pctl", "Ploopl", "Pstl", "Picc")})
    return dict(out=out, binned=binned, chained=ch, co=co, bird=bird)


def test_reference_call_sequence_matches_goldens(chain, golden2):
    for key, got in chain["out"].items():
        assert got.shape == golden2[key].shape, key
        assert rowmax_rel(got, golden2[key]) <= TOL, key


def test_unbatched_bird_keeps_reference_shapes(golden2):
    from eftpipe_b200 import pybird

    g = golden2
    co = pybird.Common(Nl=3)
    nl, rs = pybird.NonLinear(load=False, save=False, co=co), pybird.Resum(co=co)
    bird = pybird.Bird(g["kin"], g["plin"][0], float(g["f"][0]), co=co)
    nl.PsCf(bird)
    bird.setPsCfl()
    rs.Ps(bird)
    assert tuple(bird.Ploopl.shape) == (3, 12, 50) and tuple(bird.P22.shape) == (28, 50)
    assert rowmax_rel(_np(bird.Ploopl), g["res_Ploopl"][0]) <= TOL
    with pytest.raises(RuntimeError):
        pybird.Bird(g["kin"], g["plin"][0], 0.8, co=pybird.Common(Nl=3)).setPsCfl()


def _params(nuis_row):
    from eftpipe_b200 import synthetic

    b1, c2, b3, c4, cct, cr1, cr2, ce0, cemono, cequad = nuis_row.T
    b2, b4 = synthetic.c2c4_to_b2b4(c2, c4)
    return dict(b1=b1, b2=b2, b3=b3, b4=b4, cct=cct, cr1=cr1, cr2=cr2, ce0=ce0, cemono=cemono, cequad=cequad)


def test_bias_reduction_and_gaussian_table(chain, golden2):
    from eftpipe_b200 import parambasis

    basis = parambasis.WestCoastBasis(prefix="")
    params = _params(golden2["nuisance"])
    got = _np(basis.reduce_Plk(chain["binned"], params).sum())
    assert rowmax_rel(got, golden2["reduced_binned"]) <= TOL
    got = _np(basis.reduce_Plk(chain["chained"], params).sum())
    assert rowmax_rel(got, golden2["reduced_chained"]) <= TOL
    table = basis.reduce_Plk_gaussian_table(chain["binned"], params)
    for i, name in enumerate(("b3", "cct", "cr1", "cr2", "ce0", "cemono", "cequad")):
        assert rowmax_rel(_np(table[name]), golden2["gaussian_table_binned"][:, i]) <= TOL, name
    # function form with explicit lists (parambasis.py:42-136), first cosmology only
    p0 = {k: float(v[0]) for k, v in params.items()}
    one = parambasis.reduce_Plk(chain["binned"], [p0[n] for n in ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2")],
                                es=[p0[n] for n in ("ce0", "cemono", "cequad")]).sum()
    assert rowmax_rel(_np(one)[0], golden2["reduced_binned"][0]) <= TOL


def test_eastcoast_basis_against_oracle(chain):
    import pybird_oracle as orc
    from eftpipe_b200 import parambasis

    binned = chain["binned"]
    f = _np(chain["bird"]._f)
    basis = parambasis.EastCoastBasis(prefix="e_")
    vals = dict(e_b1=2.0, e_b2=0.3, e_bG2=-0.2, e_bGamma3=0.1, e_c0=5.0, e_c2=10.0, e_c4=-3.0, e_Pshot=0.4, e_a0=0.2, e_a2=-0.1)
    # counterform is a Common property (pybird.py:526): east-coast needs its own Common
    from eftpipe_b200 import pybird
    from eftpipe_b200.transformer import PlainBird

    co_e = pybird.Common(Nl=3, kmA=0.7, krA=0.25, ndA=4.5e-5, counterform="eastcoast")
    view = PlainBird(None, co_e, binned._T, binned._picc, binned.B, False, binned._f_bm)
    got = _np(basis.reduce_Plk(view, vals).sum())
    table = basis.reduce_Plk_gaussian_table(view, vals)
    oco = orc.Common(Nl=3, kmA=0.7, krA=0.25, ndA=4.5e-5, counterform="eastcoast")
    for i in range(binned.B):
        terms = {n: _np(getattr(binned, n))[i] for n in ("P11l", "Ploopl", "Pctl", "Pstl", "Picc")}
        b1, b2, bG2, bG3, c0, c2, c4 = (vals["e_" + n] for n in ("b1", "b2", "bG2", "bGamma3", "c0", "c2", "c4"))
        fi = f[i]
        bsA = [b1, b1 + 3.5 * bG2, b1 + 15 * bG2 + 6 * bG3, 0.5 * b2 - 3.5 * bG2, c0 - fi / 3 * c2 + 3 / 35 * fi**2 * c4,
               c2 - 6 / 7 * fi * c4, c4]
        es = [vals["e_Pshot"], vals["e_a0"] + vals["e_a2"] / 3, 2 / 3 * vals["e_a2"]]
        ref = orc.reduce_Plk(oco, fi, terms, bsA, es=es)
        assert rowmax_rel(got[i], ref) <= TOL
        # parambasis.py:433-437: dP/dc4
        ref_c4 = -6 / 35 * fi**2 * terms["Pctl"][:, 0] + 12 / 7 * fi**2 * terms["Pctl"][:, 1] - 2.0 * fi**2 * terms["Pctl"][:, 2]
        assert rowmax_rel(_np(table["e_c4"])[i], ref_c4) <= TOL


def test_single_tracer_marginalised_likelihood(chain, golden2, dr16):
    """EFTLike on the DR16 NGC LRG data/covariance: logp (Jeffreys and Gaussian priors), best fit."""
    from eftpipe_b200 import likelihood, parambasis

    g = golden2
    basis = parambasis.WestCoastBasis(prefix="")
    binned = chain["binned"]
    nk = g["kout"].size
    rows = np.arange(3 * nk, dtype=np.int32)
    names = ["b3", "cct", "cr1", "cr2", "ce0", "cequad"]
    tr = dict(basis=basis, co=chain["co"], nout=3 * nk, nterm=24, rows=rows, picc=binned._picc.reshape(-1))
    params = _params(g["nuisance"])
    sampled = {k: params[k] for k in ("b1", "b2", "b4")}
    import torch
    from eftpipe_b200.engine import DeviceLikelihood

    for jeff, scales, col in ((True, None, 0), (False, [4, 2, 4, 4, 2, 2], 8)):
        sig = None if scales is None else np.diag(1.0 / np.array(scales, float) ** 2)
        spec = likelihood.build_spec([tr], g["lrg_data"], g["lrg_invcov"], gaussian=names, sigma_inv=sig, jeffreys=jeff)
        dev = DeviceLikelihood(spec)
        nuis = likelihood.pack_nuisance(torch, [basis], sampled, [binned._f_bm], binned.B, binned._T.shape[-1])
        logp, status, best, full = dev.eval(binned.B, [binned._T.contiguous()], [binned._f_bm], nuis, want_bestfit=True,
                                            want_fullchi2=True)
        ref = g["marg_out"]
        np.testing.assert_allclose(_np(logp), ref[:, col], rtol=1e-6)  # north star: chi^2 to 1e-6
        np.testing.assert_allclose(_np(full), ref[:, col + 1], rtol=1e-6)  # marginal.py:129-131 fullchi2
        np.testing.assert_allclose(_np(best), ref[:, col + 2 : col + 8], rtol=1e-5, atol=1e-8)
        assert not _np(status).any()
        res = _np(dev.residuals(binned.B))  # left in the workspace by eval
        vec = _np(dev.vectors(binned.B, [binned._T.contiguous()], [binned._f_bm], nuis))
        assert rowmax_rel(vec[:, :, 0] + g["lrg_data"], g["marg_PNG"]) <= TOL
        assert np.array_equal(res, vec[:, :, 0])


def test_callable_priors_on_device(chain, golden2):
    """Callable `loc` / `scale` (marginal.py:13-20, :60-77): the prior of every point is evaluated from its own sampled
    parameters (`Marginalizable.point_priors`) and enters the kernel as per-point location / inverse variance
    (eftb_like_eval_priors); against the oracle's marginalisation point by point."""
    import pybird_oracle as orc
    import torch
    from eftpipe_b200 import likelihood, marginal, parambasis
    from eftpipe_b200.engine import DeviceLikelihood

    g = golden2
    basis = parambasis.WestCoastBasis(prefix="")
    binned = chain["binned"]
    B, nk = binned.B, g["kout"].size
    names = ["b3", "cct", "cr1", "cr2", "ce0", "cequad"]
    prior = {
        "b3": {"loc": "lambda b1: 0.5 * b1", "scale": 4.0},
        "cct": {"loc": 0.0, "scale": "lambda b1, b2: 2.0 + np.abs(b2)"},
        "cr1": {"loc": "lambda b2, b4: b2 - b4", "scale": "lambda b4: np.exp(0.1 * b4) * 4"},
        "cr2": {"loc": 1.5, "scale": 4.0},
        "ce0": {"loc": 0, "scale": 2.0},
        "cequad": {"scale": 2.0},
    }

    class M(marginal.Marginalizable):
        def marginalizable_params(self):
            return names

    m = M()
    m.setup_prior(prior)
    params = _params(g["nuisance"])
    sampled = {k: params[k] for k in ("b1", "b2", "b4")}
    env = {k: np.asarray(v, float) for k, v in sampled.items()}
    loc, sinv = m.point_priors(env, B)
    tr = dict(basis=basis, co=chain["co"], nout=3 * nk, nterm=24, rows=np.arange(3 * nk, dtype=np.int32), picc=binned._picc.reshape(-1))
    spec = likelihood.build_spec([tr], g["lrg_data"], g["lrg_invcov"], gaussian=names, jeffreys=False)
    dev = DeviceLikelihood(spec)
    nuis = likelihood.pack_nuisance(torch, [basis], sampled, [binned._f_bm], B, binned._T.shape[-1])
    args = (B, [binned._T.contiguous()], [binned._f_bm], nuis)
    logp, status, best, full = dev.eval(*args, want_bestfit=True, want_fullchi2=True,
                                        prior_loc=torch.as_tensor(loc, device="cuda"), prior_sigma_inv=torch.as_tensor(sinv, device="cuda"))
    assert not _np(status).any()
    vec = _np(dev.vectors(*args))
    for i in range(B):
        point = {"np": np, **{k: float(v[i]) for k, v in env.items()}}
        mu, sig = orc.prior_mu_sigma_inv(m.valid_prior, point)
        ref = orc.marginalized_logp(vec[i, :, 0] + g["lrg_data"], vec[i, :, 1:].T, g["lrg_data"], g["lrg_invcov"], mu_G=mu,
                                    sigma_inv=sig, jeffreys=False, return_bestfit=True)
        assert abs(_np(logp)[i] - ref[0]) <= 1e-6 * abs(ref[0])
        assert abs(_np(full)[i] - ref[1]) <= 1e-6 * abs(ref[1])
        np.testing.assert_allclose(_np(best)[i], ref[2], rtol=1e-5, atol=1e-8)
    # constant priors through the per-point entry = the plan-constant path
    scales = np.array([4, 2, 4, 4, 2, 2], float)
    mu_c = np.array([0.3, 0, -0.2, 0, 0.1, 0])
    spec_c = likelihood.build_spec([tr], g["lrg_data"], g["lrg_invcov"], gaussian=names, sigma_inv=np.diag(1 / scales**2), mu=mu_c)
    a = DeviceLikelihood(spec_c).eval(*args, want_bestfit=True, want_fullchi2=True)
    tile = lambda v: torch.as_tensor(np.ascontiguousarray(np.tile(v, (B, 1))), device="cuda")
    b = dev.eval(*args, want_bestfit=True, want_fullchi2=True, prior_loc=tile(mu_c), prior_sigma_inv=tile(1 / scales**2))
    for x, y in zip((a[0], a[2], a[3]), (b[0], b[2], b[3])):
        np.testing.assert_allclose(_np(x), _np(y), rtol=1e-12, atol=1e-12)
    with pytest.raises(ValueError):
        dev.eval(*args, prior_loc=tile(mu_c))


def test_non_positive_definite_is_flagged_not_fatal(chain, golden2):
    """marginal.py:113-116 raises; the batch path flags the point and keeps going."""
    import torch
    from eftpipe_b200 import likelihood, parambasis
    from eftpipe_b200.engine import DeviceLikelihood

    basis = parambasis.WestCoastBasis(prefix="")
    binned = chain["binned"]
    nk = golden2["kout"].size
    tr = dict(basis=basis, co=chain["co"], nout=3 * nk, nterm=24, rows=np.arange(3 * nk, dtype=np.int32),
              picc=binned._picc.reshape(-1))
    spec = likelihood.build_spec([tr], golden2["lrg_data"], -golden2["lrg_invcov"], gaussian=["b3", "cct"], jeffreys=True)
    nuis = likelihood.pack_nuisance(torch, [basis], dict(b1=2.0), [binned._f_bm], binned.B, binned._T.shape[-1])
    logp, status, _ = DeviceLikelihood(spec).eval(binned.B, [binned._T.contiguous()], [binned._f_bm], nuis)
    assert (_np(status) == 1).all() and np.isneginf(_np(logp)).all()


@pytest.fixture(scope="module")
def dr16_setup(dr16):
    """BASELINE config 3: LRG x ELG x cross, DR16 NGC production yaml
    (cobaya/yamls/DR16_noric_LEX_..._kmax0.20.yaml:7-111)."""
    from eftpipe_b200 import likelihood, synthetic, theory

    ap = dict(Om_AP=0.307115, rdrag_AP=147.66, h_AP=0.6777, APst=True)
    tracers = {
        "LRG_NGC": dict(prefix="LRG_NGC_", z=0.696, nd=4.5e-5, window=dict(window_configspace_array=dr16["win_LRG"])),
        "ELG_NGC": dict(prefix="ELG_NGC_", z=0.849, nd=2.3e-4, window=dict(window_configspace_array=dr16["win_ELG"])),
        "X_NGC": dict(prefix="X_NGC_", z=0.763, cross=["LRG_NGC", "ELG_NGC"], window=dict(window_configspace_array=dr16["win_X"])),
        "default": dict(km=0.7, kr=0.25, with_IRresum=True, with_APeffect=True, with_window=True, APeffect=ap,
                        window=dict(accboost=4, windowk=0.1)),
    }
    west = {n: {"scale": None} for n in ("b3", "cct", "cr1", "cr2", "ce0", "cequad")}
    marg = {"LRG_NGC_": west, "ELG_NGC_": dict(west), "X_NGC_ce0": {"scale": None}, "X_NGC_cequad": {"scale": None}}
    like = likelihood.EFTLike(
        tracers=["LRG_NGC", "ELG_NGC", "X_NGC"], chained=[False, True, False],
        data={"LRG_NGC": dict(table=dr16["NGC_LRG_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20),
              "ELG_NGC": dict(table=dr16["NGC_ELG_Q"], ls=[0, 2], kmin=0.03, kmax=0.20, symbol="Q"),
              "X_NGC": dict(table=dr16["NGC_X_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20)},
        cov=dict(matrix=dr16["cov_NGC_L024E02X024_PQP"], Nreal=1000), with_binning=True, jeffreys=True, marg=marg)
    th = theory.EFTLSS(tracers).must_provide(like.get_requirements()).initialize()
    like.initialize_with_provider(th)
    B = 6
    cosmo, batches = {}, {}
    for name, z in (("LRG_NGC", 0.696), ("ELG_NGC", 0.849), ("X_NGC", 0.763)):
        b = synthetic.make_batch(B, z, seed=20261018 + 3)
        batches[name] = b
        cosmo[name] = dict(pkh=b.plin, f=b.f, DA=b.DA, H=b.H)
    rng = np.random.default_rng(11)
    params = {}
    for pre, b1 in (("LRG_NGC_", 2.1), ("ELG_NGC_", 1.4)):
        params[pre + "b1"] = b1 + 0.05 * rng.standard_normal(B)
        c2 = 0.7 + 0.1 * rng.standard_normal(B)
        params[pre + "b2"], params[pre + "b4"] = c2 / np.sqrt(2), c2 / np.sqrt(2)
    return dict(th=th, like=like, cosmo=cosmo, batches=batches, params=params, B=B, tracers=tracers)


def test_multitracer_likelihood_against_oracle(dr16_setup, dr16):
    """Config 3 end to end (3 tracers, 142 data points, 14 marginalised parameters): GPU logp vs the
    oracle assembled exactly as the reference does (EFTLike.PNG/PG + marginalized_logp)."""
    import pybird_oracle as orc

    S = dr16_setup
    th, like = S["th"], S["like"]
    assert like.ndata == 142 and len(like.gaussian_names) == 14
    assert like.hartlap == pytest.approx((1000 - 142 - 2) / 999)
    th.calculate(S["cosmo"])
    # derived parameters of theory.py:620-648 and the per-tracer parameter view (theory.py:262-263)
    from eftpipe_b200 import synthetic as syn

    bl = S["batches"]["LRG_NGC"]
    np.testing.assert_allclose(th.derived["LRG_NGC_alperp"], bl.DA / syn.angular_distance(0.307115, 0.696), rtol=1e-12)
    np.testing.assert_allclose(th.derived["LRG_NGC_alpara"], syn.hubble(0.307115, 0.696) / bl.H, rtol=1e-12)
    np.testing.assert_allclose(th.derived["ELG_NGC_fz"], S["batches"]["ELG_NGC"].f)
    vals = th.get_eft_params_values_dict("LRG_NGC", S["params"])  # theory.py:839-843: every EFT parameter, absent ones 0.0
    assert {"LRG_NGC_b1", "LRG_NGC_b2", "LRG_NGC_b4", "LRG_NGC_b3", "LRG_NGC_cct", "LRG_NGC_cequad"} <= set(vals)
    assert vals["LRG_NGC_b3"] == 0.0 and vals["LRG_NGC_b1"] is S["params"]["LRG_NGC_b1"]
    ls, kk, comp = th.get_bird_component("LRG_NGC", {k: v for k, v in S["params"].items() if k.startswith("LRG")})
    assert ls == [0, 2, 4] and kk.size == 18 and tuple(comp.sum().shape) == (S["B"], 3, 18)
    res = like.calculate(S["params"], want_bestfit=True)
    logp = _np(res["logp"])
    png, pg = like.PNG_PG(S["params"])
    png, pg = _np(png), _np(pg)

    # ---- oracle: per tracer pipeline, then the reference's flatten / marginalisation ----
    names = like.gaussian_names
    scal = {"LRG_NGC": (0.7, 0.25, 4.5e-5), "ELG_NGC": (0.7, 0.25, 2.3e-4)}
    worst_png = worst_pg = 0.0
    for i in (0, S["B"] - 1):
        PNG, PG = [], {n: [] for n in names}
        for name, z in (("LRG_NGC", 0.696), ("ELG_NGC", 0.849), ("X_NGC", 0.763)):
            if name == "X_NGC":
                (kmA, krA, ndA), (kmB, krB, ndB) = scal["LRG_NGC"], scal["ELG_NGC"]
            else:
                kmA, krA, ndA = kmB, krB, ndB = scal[name]
            co = orc.Common(Nl=3, kmA=kmA, krA=krA, ndA=ndA, kmB=kmB, krB=krB, ndB=ndB)
            nl, rs = orc.NonLinear(co), orc.Resum(co)
            apo = orc.APeffect(co, Om_AP=0.307115, z_AP=z, APst=True)
            b = S["batches"][name]
            bird = orc.Bird(co, b.kin, b.plin[i], b.f[i], b.DA[i], b.H[i], z)
            nl.PsCf(bird)
            orc.set_PsCfl(bird)
            rs.Ps(bird)
            apo.AP(bird)
            wplan = th.plans[name].host  # window matrices are validated against the reference in test_host_mirror
            win = S["tracers"]
            from eftpipe_b200 import window as W, pybird as pb

            wobj = W.Window(window_configspace_array=dr16["win_" + name.split("_")[0]], co=pb.Common(Nl=3), accboost=4, windowk=0.1)
            orc.apply_window(bird, orc.mask_and_measure(wobj.Wal, wobj.p, co.k, 0.1), wobj.p, window_st=True)
            m = like.minfodict[name]
            terms = orc.Binning(m.kout, co).transform(orc.bird_terms(bird))
            if like.chained[name]:
                terms = orc.chained_transform(terms, 3)
            p = S["params"]
            pa = lambda pre: [p[pre + "b1"][i], p[pre + "b2"][i], 0.0, p[pre + "b4"][i], 0.0, 0.0, 0.0]
            if name == "X_NGC":
                bsA, bsB = pa("LRG_NGC_"), pa("ELG_NGC_")
                plk = orc.reduce_Plk(co, b.f[i], terms, bsA, bsB)
                tab = orc.gaussian_table_west(co, b.f[i], terms, bsA[0], bsB[0], cross=True)
                tab = {**{"LRG_NGC_" + k[2:]: v for k, v in tab.items() if k.startswith("A_")},
                       **{"ELG_NGC_" + k[2:]: v for k, v in tab.items() if k.startswith("B_")},
                       **{"X_NGC_" + k: v for k, v in tab.items() if k in ("ce0", "cemono", "cequad")}}
            else:
                pre = name + "_"
                plk = orc.reduce_Plk(co, b.f[i], terms, pa(pre))
                tab = {pre + k: v for k, v in orc.gaussian_table_west(co, b.f[i], terms, p[pre + "b1"][i]).items()}
            flat = lambda arr: np.hstack([arr[ell // 2, m.kout_mask[ell]] for ell in m.ls])
            PNG.append(flat(plk))
            for n in names:
                PG[n].append(flat(tab[n]) if n in tab else np.zeros(m.data_vector.size))
        PNG = np.hstack(PNG)
        PGm = np.array([np.hstack(PG[n]) for n in names])
        worst_png = max(worst_png, rowmax_rel(png[i], PNG))
        worst_pg = max(worst_pg, rowmax_rel(pg[i], PGm))
        ref_logp, _, ref_best = orc.marginalized_logp(PNG, PGm, like.data_vector, like.invcov, jeffreys=True, return_bestfit=True)
        assert logp[i] == pytest.approx(ref_logp, rel=1e-6), i
        best = np.array([_np(res["bestfit"]["marg_" + n])[i] for n in names])
        np.testing.assert_allclose(best, ref_best, rtol=1e-4, atol=1e-6)
    assert worst_png <= TOL and worst_pg <= TOL


def test_eftlike_with_callable_priors(dr16_setup, dr16):
    """EFTLike with string `loc` / `scale` entries in `marg` (likelihood.py:560-564 `env` = every tracer's EFT parameters):
    three tracers, 14 marginalised parameters, Gaussian priors whose location / width follow the sampled b1."""
    import pybird_oracle as orc
    from eftpipe_b200 import likelihood

    S = dr16_setup
    th = S["th"]
    west = lambda: {n: {"scale": 4.0} for n in ("b3", "cct", "cr1", "cr2", "ce0", "cequad")}
    lrg, elg = west(), west()
    lrg["b3"] = {"loc": "lambda LRG_NGC_b1: 0.5 * LRG_NGC_b1", "scale": 2.0}
    elg["cct"] = {"loc": 0.0, "scale": "lambda ELG_NGC_b1, LRG_NGC_b2: 1.0 + np.abs(ELG_NGC_b1 * LRG_NGC_b2)"}
    marg = {"LRG_NGC_": lrg, "ELG_NGC_": elg, "X_NGC_ce0": {"scale": 2.0}, "X_NGC_cequad": {"loc": "lambda ELG_NGC_b4: -ELG_NGC_b4", "scale": 2.0}}
    like = likelihood.EFTLike(
        tracers=["LRG_NGC", "ELG_NGC", "X_NGC"], chained=[False, True, False],
        data={"LRG_NGC": dict(table=dr16["NGC_LRG_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20),
              "ELG_NGC": dict(table=dr16["NGC_ELG_Q"], ls=[0, 2], kmin=0.03, kmax=0.20, symbol="Q"),
              "X_NGC": dict(table=dr16["NGC_X_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20)},
        cov=dict(matrix=dr16["cov_NGC_L024E02X024_PQP"], Nreal=1000), with_binning=True, jeffreys=False, marg=marg)
    like.initialize_with_provider(th)
    assert like.callable_prior and len(like.gaussian_names) == 14
    th.calculate(S["cosmo"])
    res = like.calculate(S["params"], want_bestfit=True)
    png, pg = like.PNG_PG(S["params"])
    png, pg = _np(png), _np(pg)
    assert not _np(res["status"]).any()
    for i in range(S["B"]):
        point = {"np": np, **{k: float(v[i]) for k, v in S["params"].items()}}
        mu, sig = orc.prior_mu_sigma_inv(like.valid_prior, point)
        ref_logp, ref_full, ref_best = orc.marginalized_logp(png[i], pg[i], like.data_vector, like.invcov, mu_G=mu, sigma_inv=sig,
                                                             jeffreys=False, return_bestfit=True)
        assert _np(res["logp"])[i] == pytest.approx(ref_logp, rel=1e-6), i
        assert _np(res[like.likelihood_prefix + "fullchi2"])[i] == pytest.approx(ref_full, rel=1e-6), i
        best = np.array([_np(res["bestfit"][like.marg_param_prefix + n])[i] for n in like.gaussian_names])
        np.testing.assert_allclose(best, ref_best, rtol=1e-4, atol=1e-6)


def test_theory_and_likelihood_are_graph_capturable(dr16_setup):
    """EFTLSS.calculate + EFTLike.calculate on device-resident inputs replay from one CUDA graph (no host round trips
    inside: derived parameters are lazy) and give the eager results bit for bit."""
    import torch

    from eftpipe_b200.engine import capture_graph

    S = dr16_setup
    th, like = S["th"], S["like"]
    dev = lambda x: torch.as_tensor(np.asarray(x, float), device="cuda")
    cosmo = {t: {k: dev(v) for k, v in c.items()} for t, c in S["cosmo"].items()}
    params = {k: dev(v) for k, v in S["params"].items()}

    def step():
        th.calculate(cosmo)
        res = like.calculate(params)
        return res["logp"], res["status"]

    eager = _np(step()[0]).copy()
    graph, (g_logp, g_status) = capture_graph(step)
    graph.replay()
    torch.cuda.synchronize()
    assert np.array_equal(_np(g_logp), eager) and not _np(g_status).any()
    assert "LRG_NGC_alperp" in th.derived  # evaluated on demand, after the capture


def test_custom_window_plugin_through_eftlss(golden2):
    """`with_window: "helpers.MatrixWindow"` (theory.py:62-72, :370-377): the plugin's operator - here the reference's
    window + integral-constraint operator of config 2 - is probed at plan build and composed with the binning; the
    batched terms reproduce the reference's binned goldens."""
    import helpers
    from eftpipe_b200 import theory

    g = golden2
    ap = json.loads(str(g["ap"]))
    tracers = {"LRG": dict(prefix="", z=float(g["z"]), nd=4.5e-5, km=0.7, kr=0.25, with_IRresum=True, with_APeffect=True,
                           APeffect=dict(rdrag_AP=147.66, h_AP=0.6777, **ap), with_window="helpers.MatrixWindow",
                           window=dict(matrix=0.95 * g["Weff_LRG"], picc=-g["PSN"] * float(g["Pshot"])))}
    th = theory.EFTLSS(tracers).must_provide(
        {"nonlinear_Plk_grid": {"LRG": {"ls": [0, 2, 4], "binned": True, "binning": {"kout": g["kout"]}}}}).initialize()
    th.calculate({"LRG": dict(pkh=g["plin"], f=g["f"], DA=g["DA"], H=g["H"])})
    bm, _ = th.get_nonlinear_Plk_terms("LRG")
    nk = g["kout"].size
    T = helpers.split_terms(_np(bm).reshape(3, nk, 24, -1)[..., : len(g["f"])].transpose(3, 0, 2, 1))
    for name, arr in T.items():
        assert rowmax_rel(arr, g["bin_" + name]) <= TOL, name
    assert rowmax_rel(th.info["LRG"]["picc"].reshape(3, nk), g["bin_Picc"][0]) <= TOL


YAML_NGC = """
theory:
  eftpipe.classynu:
    extra_args: {neutrino_hierarchy: degenerate}
  eftpipe.eftlss:
    tracers:
      LRG_NGC: {prefix: LRG_NGC_, z: 0.696, nd: 4.5e-5,
                window: {window_fourier_file: cache/DR16_noric_NGC_LRG_acc4.npy, window_configspace_file: data/win_NGC_LRG.txt}}
      ELG_NGC: {prefix: ELG_NGC_, z: 0.849, nd: 2.3e-4,
                window: {window_fourier_file: cache/DR16_noric_NGC_ELG_acc4.npy, window_configspace_file: data/win_NGC_ELG.txt}}
      X_NGC: {prefix: X_NGC_, z: 0.763, cross: [LRG_NGC, ELG_NGC],
              window: {window_fourier_file: cache/DR16_noric_NGC_X_acc4.npy, window_configspace_file: data/win_NGC_X.txt}}
      default:
        provider: classynu
        km: 0.7
        kr: 0.25
        use_cb: true
        with_IRresum: true
        with_APeffect: true
        with_window: true
        APeffect: {Om_AP: 0.307115, rdrag_AP: 147.66, h_AP: 0.6777, APst: true}
        window: {accboost: 4, windowk: 0.1}
likelihood:
  LEX_NGC:
    class: eftpipe.eftlike
    tracers: [LRG_NGC, ELG_NGC, X_NGC]
    chained: [false, true, false]
    data:
      LRG_NGC: {path: data/NGC_LRG_P.txt, ls: [0, 2, 4], kmin: 0.02, kmax: 0.20}
      ELG_NGC: {path: data/NGC_ELG_Q.txt, ls: [0, 2], kmin: 0.03, kmax: 0.20}
      X_NGC: {path: data/NGC_X_P.txt, ls: [0, 2, 4], kmin: 0.02, kmax: 0.20}
    cov: {path: data/cov_NGC_L024E02X024_PQP.txt, Nreal: 1000}
    with_binning: true
    jeffreys: true
    marg:
      LRG_NGC_: &westcoast_hex
        b3: {scale: }
        cct: {scale: }
        cr1: {scale: }
        cr2: {scale: }
        ce0: {scale: }
        cequad: {scale: }
      ELG_NGC_: *westcoast_hex
      X_NGC_ce0: {scale: }
      X_NGC_cequad: {scale: }
sampler:
  mcmc: {}
"""


def test_cobaya_yaml_drop_in(dr16_setup, dr16, tmp_path):
    """The reference's production input (cobaya/yamls/DR16_noric_LEX_..._kmax0.20.yaml, NGC part, same keys) loaded as
    is: files on disk, yaml anchors, `default:` block, cache files - gives the log-posterior of the hand-built setup."""
    from eftpipe_b200 import cobaya_info

    (tmp_path / "data").mkdir()
    for name, hdr in (("NGC_LRG_P", "k P0 P2 P4"), ("NGC_ELG_Q", "k Q0 Q2"), ("NGC_X_P", "k P0 P2 P4")):
        np.savetxt(tmp_path / "data" / f"{name}.txt", dr16[name], header=hdr)
    for t in ("LRG", "ELG", "X"):
        np.savetxt(tmp_path / "data" / f"win_NGC_{t}.txt", dr16[f"win_{t}"])
    np.savetxt(tmp_path / "data" / "cov_NGC_L024E02X024_PQP.txt", dr16["cov_NGC_L024E02X024_PQP"])
    (tmp_path / "run.yaml").write_text(YAML_NGC)
    th, likes = cobaya_info.load_info(str(tmp_path / "run.yaml"))
    like = likes["LEX_NGC"]
    assert like.ndata == 142 and len(like.gaussian_names) == 14
    assert (tmp_path / "cache" / "DR16_noric_NGC_LRG_acc4.npy").exists()  # window cache written where the yaml says
    S = dr16_setup
    th.calculate(S["cosmo"])
    S["th"].calculate(S["cosmo"])
    got, want = _np(like.logp(S["params"])), _np(S["like"].logp(S["params"]))
    np.testing.assert_allclose(got, want, rtol=1e-12)
    # second load: the cached Fourier-space windows are read back (meta check passes), same result
    th2, likes2 = cobaya_info.load_info(str(tmp_path / "run.yaml"))
    th2.calculate(S["cosmo"])
    np.testing.assert_allclose(_np(likes2["LEX_NGC"].logp(S["params"])), want, rtol=1e-12)


def test_per_call_taper_and_mutable_term_arrays(golden2):
    """pybird.py:1143 `PsCf(bird, window=0.2)` takes the FFTLog taper per call; the stages of the reference rebind and
    update the Bird's term arrays in place (pybird.py:1445, :1613; SURVEY.md 8b ownership)."""
    import pybird_oracle as orc
    from eftpipe_b200 import pybird

    g = golden2
    co = pybird.Common(Nl=3)
    nl, rs = pybird.NonLinear(load=False, save=False, co=co), pybird.Resum(co=co)
    oco = orc.Common(Nl=3)
    onl = orc.NonLinear(oco)
    for w in (0.3, 0.2):  # a non-default taper first, then back to the default on a fresh bird
        bird = pybird.Bird(g["kin"], g["plin"], g["f"], co=co)
        nl.PsCf(bird, window=w)
        ob = orc.Bird(oco, g["kin"], g["plin"][1], g["f"][1])
        onl.PsCf(ob, window=w)
        assert rowmax_rel(_np(bird.P22)[1], ob.P22) <= TOL and rowmax_rel(_np(bird.C13)[1], ob.C13) <= TOL
    assert rowmax_rel(_np(bird.P22), g["P22"]) <= TOL  # the last one is the golden's default taper
    bird.setPsCfl()
    rs.Ps(bird)
    before = _np(bird.Ploopl).copy()
    bird.Ploopl = bird.Ploopl * 2.0             # rebinding (a new array)
    assert np.array_equal(_np(bird.Ploopl), 2.0 * before)
    bird.Ploopl += 1.0                          # in-place update through the view
    assert np.array_equal(_np(bird.Ploopl), 2.0 * before + 1.0)
    bird.P11l = np.zeros((3, 3, 3, 50))         # numpy input, (B, Nl, 3, Nk)
    assert not _np(bird.P11l).any() and np.array_equal(_np(bird.Pctl), _np(bird.Pctl))
    bird.Picc = bird.Picc - 1.5                 # window.py:405 style
    assert np.allclose(_np(bird.Picc), -1.5)
    with pytest.raises(AttributeError):
        bird.PctNNLOl = np.zeros((3, 3, 3, 50))


def test_custom_basis_by_dotted_path(chain, golden2, dr16):
    """parambasis.py:457-465: `basis: "pkg.mod.Class"` with a plain numpy EFTBasis.  Its reductions are probed per point into
    explicit bias columns (parambasis.ProbedBasis) and run through the same kernels: reduce_Plk, the Gaussian table and a
    marginalised likelihood against the basis' own numpy code on the reference-pinned binned terms."""
    import pybird_oracle as orc
    from types import SimpleNamespace

    import custom_basis
    from eftpipe_b200 import likelihood, parambasis

    cls = parambasis.find_param_basis("custom_basis.ToyBasis")
    basis = cls(prefix="t_")
    assert isinstance(basis, parambasis.ProbedBasis) and basis.gaussian_params() == ["t_c0", "t_e0"]
    binned, co = chain["binned"], chain["co"]
    B = binned.B
    rng = np.random.default_rng(3)
    params = {"t_b1": 2.0 + 0.1 * rng.standard_normal(B), "t_s": 0.3 + 0.1 * rng.standard_normal(B), "t_c0": rng.standard_normal(B),
              "t_e0": 0.2}
    got = _np(basis.reduce_Plk(binned, params).sum())
    table = basis.reduce_Plk_gaussian_table(binned, params)
    assert set(table) == {"t_c0", "t_e0"}
    inner = custom_basis.ToyBasis(prefix="t_")
    f = _np(chain["bird"]._f)
    terms = {n: golden2["bin_" + n] for n in ("P11l", "Pctl", "Ploopl", "Pstl", "Picc")}
    nk = terms["P11l"].shape[-1]
    rows = np.arange(3 * nk, dtype=np.int32)
    refs = []
    for i in range(B):
        nb = SimpleNamespace(co=SimpleNamespace(No=3, kmA=co.kmA, ndA=co.ndA), f=f[i], PctNNLOl=None, **{n: v[i] for n, v in terms.items()})
        p = {k: (float(v[i]) if np.ndim(v) else v) for k, v in params.items()}
        ref = inner.reduce_Plk(nb, p).sum()
        assert rowmax_rel(got[i], ref) <= TOL
        tab = inner.reduce_Plk_gaussian_table(nb, p)
        for n in tab:
            assert rowmax_rel(_np(table[n])[i], tab[n]) <= TOL, n
        refs.append((inner.reduce_Plk(nb, dict(p, t_c0=0.0, t_e0=0.0)).sum().reshape(-1), np.array([tab["t_c0"].reshape(-1), tab["t_e0"].reshape(-1)])))
    # marginalised likelihood over (c0, e0) with the custom basis in the device spec
    minfo = likelihood.MultipoleInfo.load(dr16["NGC_LRG_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20)
    spec = likelihood.build_spec([dict(basis=basis, co=co, nout=3 * nk, nterm=24, rows=rows, picc=binned._picc.reshape(-1))],
                                 minfo.data_vector, golden2["lrg_invcov"], gaussian=["t_c0", "t_e0"], jeffreys=True)
    from eftpipe_b200.engine import DeviceLikelihood
    import torch

    dl = DeviceLikelihood(spec)
    free = {k: v for k, v in params.items() if k in ("t_b1", "t_s")}
    nuis = likelihood.pack_nuisance(torch, [basis], free, [binned._f_bm], B, binned._T.shape[-1], spec=spec)
    logp, status, _ = dl.eval(B, [binned._T.contiguous().reshape(3 * nk, 24, -1)], [binned._f_bm], nuis)
    for i in range(B):
        png, pg = refs[i]
        assert _np(logp)[i] == pytest.approx(orc.marginalized_logp(png, pg, minfo.data_vector, golden2["lrg_invcov"], jeffreys=True), rel=1e-8)


def test_window_matrix_stage_on_the_device(golden2):
    """window.py:479-577 on a batched device Bird: the band-power window matrix changes the node grid (50 nodes -> 23 bands);
    the term arrays are set to the seeded inputs of tests/golden/make_golden_windowmatrix.py (point 0) and to twice them
    (point 1) and compared with what the unmodified reference returned"""
    import importlib.util

    from eftpipe_b200 import pybird, window

    spec = importlib.util.spec_from_file_location("mkwm", os.path.join(os.path.dirname(__file__), "golden", "make_golden_windowmatrix.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    gw = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "window_matrix.npz")))
    co = pybird.Common(Nl=3, kmax=0.3, with_NNLO=True)
    flat, terms = mk.inputs(co.Nl, co.Nk)
    ells, kmin, kmax = [int(x) for x in gw["ells"]], float(gw["kmin"]), float(gw["kmax"])
    cut = window.to_window_matrix(flat, window.PInfo((0, 2, 4), 0, 0.4, 400), window.PInfo((0, 1, 2, 3, 4), 0, 0.4, 40),
                                  ells_in=(0, 2, 4), kmax_in=co.k.max(), ells_out=tuple(ells), kmin_out=kmin, kmax_out=kmax)
    wm = window.WindowMatrix(cut, window.PolesInfo(co.Nl, 0, co.k.max(), cut.shape[3]), window.PolesInfo(len(ells), kmin, kmax, cut.shape[2]),
                             co=co, window_st=True)
    g = golden2
    bird = pybird.Bird(g["kin"], g["plin"][:2], g["f"][:2], co=co)
    pybird.NonLinear(load=False, save=False, co=co).PsCf(bird)
    bird.setPsCfl()
    for n in ("P11l", "Pctl", "Ploopl", "Pstl", "PctNNLOl"):
        setattr(bird, n, np.stack([terms[n], 2.0 * terms[n]]))
    wm.Window(bird)
    for n in ("P11l", "Pctl", "Ploopl", "Pstl", "PctNNLOl"):
        got, want = _np(getattr(bird, n)), gw["st1." + n]
        assert got.shape == (2,) + want.shape, n
        assert rowmax_rel(got[0], want) <= 1e-11 and rowmax_rel(got[1], 2.0 * want) <= 1e-11, n
    assert _np(bird.Picc).shape[-2:] == gw["st1.Picc"].shape and not _np(bird.Picc).any()
    # window_st=False cannot keep the stochastic rows on a different grid than the rest of one term array
    bird2 = pybird.Bird(g["kin"], g["plin"][:2], g["f"][:2], co=co)
    pybird.NonLinear(load=False, save=False, co=co).PsCf(bird2)
    bird2.setPsCfl()
    with pytest.raises(ValueError, match="stochastic"):
        window.WindowMatrix(cut, wm.inpoles, wm.outpoles, co=co, window_st=False).Window(bird2)
