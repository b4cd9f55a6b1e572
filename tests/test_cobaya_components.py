"""The Cobaya-facing components (eftpipe_b200.cobaya) driven by the same mini-Cobaya that drives the unmodified
reference (oracle/refshim/cobaya).  CPU: component wiring, requirement redirection, error conventions (no device work);
GPU: the production configuration end to end against the reference's own run of the same model
(tests/golden/config3_like.npz), single points (Cobaya proper) and batches, fast / slow caching, several products per
tracer, snapshots."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, rowmax_rel

sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
import refdriver  # noqa: E402


def _info(tmp, tables, likelihoods=("jeffreys",), **kw):
    paths = refdriver.write_dr16(os.path.join(str(tmp), "dr16"))
    return refdriver.config3_info(paths, tables, package="eftpipe_b200", likelihoods=likelihoods, **kw)


def test_components_resolve_without_a_gpu():
    import eftpipe_b200
    from cobaya.likelihood import Likelihood
    from cobaya.theory import Theory

    assert issubclass(eftpipe_b200.eftlss, Theory) and issubclass(eftpipe_b200.eftlike, Likelihood)
    for name, ref in (("get_nonlinear_Plk_grid", ("self", "tracer", "chained", "binned")),
                      ("get_nonlinear_Plk_gaussian_grid", ("self", "tracer", "chained", "binned")),
                      ("get_nonlinear_Plk_interpolator", ("self", "tracer", "chained")), ("get_snapshots", ("self", "tracer")),
                      ("get_eft_params_values_dict", ("self", "tracer")), ("get_bird_component", ("self", "tracer"))):
        import inspect

        assert tuple(inspect.signature(getattr(eftpipe_b200.eftlss, name)).parameters) == ref  # theory.py:244-267


def test_requirement_redirection_and_errors():
    """theory.py:165-194: product requirements fan out to the per-tracer leaf / kernel products; unknown tracers and
    incompatible binnings are refused like in the reference"""
    from cobaya.log import LoggedError

    import eftpipe_b200

    th = eftpipe_b200.eftlss(info=dict(tracers={"A": dict(z=0.7, km=0.7, kr=0.25, nd=4.5e-5, provider="refdriver.TableExtractor",
                                                          provider_kwargs=dict(table={}))}), name="eftpipe_b200.eftlss")
    assert th.get_requirements() == {"eftleaf_A_results": {}}
    helpers = th.get_helper_theories()
    assert set(helpers) == {"eftpipe_b200.eftlss.A", "eftpipe_b200.eftlss.A.kernel"}
    out = th.must_provide(nonlinear_Plk_grid={"A": {"ls": [0, 2], "chained": [False, True], "binned": False}})
    assert set(out) == {"eftleaf_A_results", "eftleaf_kernel_A_results"}
    req = th.core.requirements["A"]
    assert req["Nl"] == 2 and req["chained"] == [False, True] and req["binned"] == [False]
    th.must_provide(nonlinear_Plk_grid={"A": {"ls": [0, 2], "chained": True}})  # chained only: needs l = 4 too
    assert th.core.requirements["A"]["Nl"] == 3
    with pytest.raises(LoggedError):
        th.must_provide(nonlinear_Plk_grid={"B": {"ls": [0]}})
    with pytest.raises(LoggedError):
        th.must_provide(nonlinear_Plk_grid={"A": {"ls": [1]}})
    with pytest.raises(LoggedError):
        th.must_provide(nonlinear_Plk_grid={"A": {"ls": [0], "binned": True}})  # binned=True but missing binning
    th.must_provide(nonlinear_Plk_grid={"A": {"ls": [0], "binned": True, "binning": {"kout": np.arange(0.02, 0.2, 0.01)}}})
    with pytest.raises(LoggedError):
        th.must_provide(nonlinear_Plk_grid={"A": {"ls": [0], "binned": True, "binning": {"kout": np.arange(0.03, 0.2, 0.01)}}})
    leaf = helpers["eftpipe_b200.eftlss.A"]
    assert set(leaf.get_requirements()) == {"A_b1", "A_b2", "A_b4", "eftleaf_kernel_A_results"}  # theory.py:285-291 default prefix
    assert "A_cct" in leaf.get_can_support_params()
    with pytest.raises(LoggedError):
        eftpipe_b200.eftlss(info=dict(tracers={"X": dict(z=0.7, cross=["A", "B"])}), name="t")


@pytest.fixture(scope="module")
def golden3():
    return dict(np.load(os.path.join(GOLDEN, "config3_like.npz")))


def _tables(g, n=None):
    return {t: {k: g[f"{t}.{k}"][:n] for k in ("pkh", "f", "DA", "H", "h", "rdrag")} for t in ("LRG_NGC", "ELG_NGC", "X_NGC")}


@pytest.mark.gpu
def test_production_model_single_points_against_reference(golden3, tmp_path):
    """the production yaml through Cobaya's protocol, one point per evaluation: logp, derived chi2 / fullchi2 / best fit,
    product getters with the reference shapes, and the fast / slow split"""
    from cobaya.model import get_model

    g = golden3
    model = get_model(_info(tmp_path, _tables(g), likelihoods=("jeffreys", "gauss")))
    like, eft = model.likelihood["LEX_NGC"], model.theory["eftpipe_b200.eftlss"]
    pts = {k[3:]: v for k, v in g.items() if k.startswith("pt.")}
    kernels = [c for n, c in model.theory.items() if n.endswith(".kernel")]
    for i in (0, 3, 7):
        post = model.logposterior({k: v[i] for k, v in pts.items()})
        assert isinstance(post.loglikes[0], float)
        assert post.loglikes[0] == pytest.approx(g["LEX_NGC.logp"][i], rel=1e-9)
        assert post.loglikes[1] == pytest.approx(g["LEX_NGC_gauss.logp"][i], rel=1e-9)
        assert post.derived["LEX_NGC_chi2"] == pytest.approx(-2 * g["LEX_NGC.logp"][i], rel=1e-9)
        assert post.derived["LEX_NGC_fullchi2"] == pytest.approx(g["LEX_NGC.fullchi2"][i], rel=1e-7)
        assert post.derived["LEX_NGC_gauss_fullchi2"] == pytest.approx(g["LEX_NGC_gauss.fullchi2"][i], rel=1e-7)
        for t, ch in (("LRG_NGC", False), ("ELG_NGC", True), ("X_NGC", False)):
            ls, k, plk = eft.get_nonlinear_Plk_grid(t, chained=ch, binned=True)
            assert list(ls) == list(g[t + ".ls"]) and plk.shape == g[t + ".Plk"][i].shape
            assert rowmax_rel(plk, g[t + ".Plk"][i]) <= 1e-8
            ls, k, tab = eft.get_nonlinear_Plk_gaussian_grid(t, chained=ch, binned=True)
            assert all(v.shape == plk.shape for v in tab.values())
        assert rowmax_rel(like.PNG(), g["LEX_NGC.PNG"][i]) <= 1e-8
        assert rowmax_rel(like.PG(), g["LEX_NGC.PG"][i]) <= 1e-8
        names = [str(n) for n in g["gaussian_names"]]
        # the derived best-fit names are shared by the two likelihoods (the later one wins, as in Cobaya): check presence
        assert all("marg_" + n in post.derived for n in names)
    # fast / slow: a nuisance-only change re-runs no tracer pipeline
    n0 = [c.n_computed for c in kernels]
    p = {k: v[7] for k, v in pts.items()}
    p["LRG_NGC_b1"] = p["LRG_NGC_b1"] + 0.01
    lp = model.logposterior(p).loglikes[0]
    assert [c.n_computed for c in kernels] == n0 and lp != pytest.approx(g["LEX_NGC.logp"][7], rel=1e-6)
    # derived parameters of the kernels (theory.py:620-648)
    d = model.logposterior({k: v[1] for k, v in pts.items()}).derived
    from eftpipe_b200 import synthetic

    assert d["LRG_NGC_fz"] == pytest.approx(g["LRG_NGC.f"][1])
    ratio = 147.66 * 0.6777 / (g["LRG_NGC.rdrag"][1] * g["LRG_NGC.h"][1])
    assert d["LRG_NGC_alperp"] == pytest.approx(g["LRG_NGC.DA"][1] / synthetic.angular_distance(0.307115, 0.696) * ratio, rel=1e-12)


@pytest.mark.gpu
def test_batched_points_and_several_products(golden3, tmp_path):
    """B = 32 points per evaluation through the same components (a batched extractor), several (chained, binned) products
    of one tracer in one evaluation (theory.py:590-604) and the Bird snapshots (theory.py:260, :576-581)"""
    import torch
    from cobaya.model import get_model

    from eftpipe_b200 import boltzmann

    g = golden3
    tabs = _tables(g)
    info = _info(tmp_path, tabs, tracer_extra=dict(IRresum=dict(snapshot=True)), window_extra=dict(snapshot=True))
    for t in ("LRG_NGC", "ELG_NGC", "X_NGC"):
        tb = tabs[t]
        info["theory"]["eftpipe_b200.eftlss"]["tracers"][t].update(
            provider=boltzmann.ArrayExtractor(tb["pkh"], tb["f"], tb["DA"], tb["H"], h=tb["h"], rdrag=tb["rdrag"]), provider_kwargs={})
    info["theory"]["eftpipe_b200.eftlss"]["tracers"]["default"]["APeffect"]["snapshot"] = True
    info["params"].pop("point")
    # a second consumer asks LRG_NGC for more products - un-binned and binned, chained and not - plus the snapshots
    from cobaya.theory import Theory

    kout = refdriver_kout("NGC_LRG_P", 0.02)

    class Consumer(Theory):
        def get_requirements(self):
            return {"nonlinear_Plk_grid": {"LRG_NGC": {"ls": [0, 2, 4], "chained": [False, True], "binned": [False, True],
                                                       "binning": {"kout": kout}}},
                    "snapshots": {"LRG_NGC": None}}

    info["theory"]["consumer"] = {"class": Consumer}
    model = get_model(info)
    eft = model.theory["eftpipe_b200.eftlss"]
    pts = {k[3:]: v for k, v in g.items() if k.startswith("pt.") and k != "pt.point"}
    post = model.logposterior(pts)
    logp = post.loglikes[0]
    assert isinstance(logp, torch.Tensor) and tuple(logp.shape) == (32,)
    np.testing.assert_allclose(logp.cpu().numpy(), g["LEX_NGC.logp"], rtol=1e-9)
    ls, k, plk = eft.get_nonlinear_Plk_grid("LRG_NGC", chained=False, binned=True)
    assert rowmax_rel(plk, g["LRG_NGC.Plk"]) <= 1e-8
    ls_c, k_c, plk_c = eft.get_nonlinear_Plk_grid("LRG_NGC", chained=True, binned=True)
    assert list(ls_c) == [0, 2] and plk_c.shape == (32, 2, 18)
    # chained multipoles from the unchained ones: Q_l = P_l - A_l P_{l+2} (chained.py:13-28)
    np.testing.assert_allclose(plk_c[:, 0], plk[:, 0] + 2 / 5 * plk[:, 1], rtol=1e-10)
    np.testing.assert_allclose(plk_c[:, 1], plk[:, 1] + 20 / 27 * plk[:, 2], rtol=1e-10)
    ls_u, k_u, plk_u = eft.get_nonlinear_Plk_grid("LRG_NGC", chained=False, binned=False)
    assert plk_u.shape == (32, 3, 50) and k_u.size == 50
    snaps = eft.get_snapshots("LRG_NGC")
    assert set(snaps) == {"IRresum", "APeffect", "window"}
    win = snaps["window"]
    # the un-binned product IS the windowed bird reduced; the snapshot exposes its term arrays (B, Nl, 12, Nk)
    assert tuple(win.Ploopl.shape) == (32, 3, 12, 50) and tuple(snaps["IRresum"].Pctl.shape) == (32, 3, 6, 50)
    from eftpipe_b200 import plan as P

    Weff = P.window_effective_matrix(*_window(model, "LRG_NGC"))
    ap_terms = snaps["APeffect"].Ploopl.cpu().numpy()
    expect = np.einsum("akln,blin->baik", Weff, ap_terms)
    assert rowmax_rel(win.Ploopl.cpu().numpy(), expect) <= 1e-9


def refdriver_kout(name, kmin, kmax=0.20):
    fx = np.load(refdriver.FIXTURE)
    k = fx[name][:, 0]
    return k[(k >= kmin) & (k <= kmax)]


def _window(model, tracer):
    """(Wal, p, k, windowk) of the tracer's window as the core built it"""
    from eftpipe_b200 import window as W
    from eftpipe_b200 import pybird as pb

    fx = np.load(refdriver.FIXTURE)
    w = W.Window(window_configspace_array=fx["win_LRG"], co=pb.Common(Nl=3), accboost=4, windowk=0.1)
    return w.Wal, w.p, pb.Common(Nl=3).k, 0.1


@pytest.mark.gpu
def test_every_product_of_one_tracer_against_reference(tmp_path):
    """The reference's own regression test of the theory component (tests/regression/test_eftlss.py::test_ELG_NGC_reg): one
    evaluation serves interpolators (plain / chained), the four (chained, binned) grid products, their Gaussian tables and
    the derived parameters.  Same requirements, same points, compared with what the unmodified reference returned
    (tests/golden/products_elg.npz, made by make_golden_products.py) - single points (floats, Cobaya proper)."""
    import importlib.util

    from cobaya.model import get_model
    from cobaya.theory import Theory

    spec = importlib.util.spec_from_file_location("mkprod", os.path.join(GOLDEN, "make_golden_products.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    g = dict(np.load(os.path.join(GOLDEN, "products_elg.npz")))
    table = {k[4:]: v for k, v in g.items() if k.startswith("tab.")}
    pts = {k[3:]: v for k, v in g.items() if k.startswith("pt.")}
    paths = refdriver.write_dr16(os.path.join(str(tmp_path), "dr16"))

    class Consumer(Theory):
        def get_requirements(self):
            return mk.requirements()

    info = mk.build_info("eftpipe_b200", paths, table)
    info["theory"]["consumer"] = {"class": Consumer}
    model = get_model(info)
    eft = model.theory["eftpipe_b200.eftlss"]
    to_np = lambda v: v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    for i in range(mk.B):
        _, derived = model.loglikes({k: v[i] for k, v in pts.items()})
        got = mk.collect(eft)
        for key, val in got.items():
            want = g[key] if key.endswith((".ls", ".k")) else g[key][i]
            val = to_np(val)
            assert val.shape == want.shape, (key, val.shape, want.shape)
            if key.endswith(".ls"):
                assert list(val) == list(want), key
            elif key.endswith(".k"):
                np.testing.assert_allclose(val, want, rtol=1e-13, err_msg=key)
            else:
                assert rowmax_rel(val, want) <= 1e-8, (i, key, rowmax_rel(val, want))
        for d in ("alperp", "alpara"):
            assert float(derived[f"{mk.TRACER}_{d}"]) == pytest.approx(float(g["derived." + d][i]), rel=1e-12), d
        assert float(derived[mk.TRACER + "_fsigma8_z"]) == float(g["derived.fsigma8_z"][i])  # the table extractor's -1
