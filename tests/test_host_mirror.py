"""CPU: host-side mirror of the reference interface - precompute and configuration logic that needs no GPU."""
import json
import os

import numpy as np
import pytest

from conftest import rowmax_rel
from eftpipe_b200 import binning, chained, fftlog, likelihood, marginal, parambasis, plan as P, pybird, window

DATA = os.path.join(os.path.dirname(os.path.abspath(pybird.__file__)), "data", "dr16_ngc.npz")


@pytest.fixture(scope="module")
def dr16():
    return dict(np.load(DATA))


def test_common_validation():
    with pytest.raises(ValueError):
        pybird.Common(Nl=2, No=3)  # pybird.py:543-544
    with pytest.raises(ValueError):
        pybird.Common(IRcutoff=True)  # pybird.py:528-529
    co = pybird.Common(Nl=3, kmA=0.7, krA=0.25, ndA=4.5e-5)
    assert (co.Nk, co.Ns, co.Nkr, co.Nklow) == (50, 80, 43, 7)
    assert co.kmB == 0.7 and co.ndB == 4.5e-5
    assert pybird.Common(kmax=0.4).Nk == 84  # pybird.py:473-477


def test_fftlog_class_matches_reference_semantics(fftlog_kat):
    with pytest.raises(ValueError):
        fftlog.FFTLog(Nmax=255, xmin=1e-5, xmax=10, bias=-0.3)  # fftlog.py:61-62
    fl = fftlog.FFTLog(Nmax=256, xmin=1e-5, xmax=10, bias=-0.3)
    x, rows = fftlog_kat["x"], fftlog_kat["rows"]
    c = fl.coef_host(x, rows, extrap="padding", window=0.3)
    assert np.abs(c - fftlog_kat["coef"]).max() <= 1e-13 * np.abs(fftlog_kat["coef"]).max()
    # the linear operator reproduces the transform
    L = fl.operator(x, window=0.3)
    assert np.abs(rows @ L.T - fftlog_kat["coef"]).max() <= 1e-12 * np.abs(fftlog_kat["coef"]).max()


def test_window_precompute_matches_reference(dr16, golden2, tmp_path):
    co = pybird.Common(Nl=3, kmA=0.7, krA=0.25, ndA=4.5e-5)
    cache = tmp_path / "win_lrg.npy"
    w = window.Window(window_fourier_file=cache, window_configspace_array=dr16["win_LRG"], co=co, accboost=4, windowk=0.1)
    assert rowmax_rel(w.effective_matrix().reshape(150, 150), golden2["Weff_LRG"].reshape(150, 150)) <= 1e-9
    # cache round trip + strict meta check (window.py:204-260)
    assert cache.exists() and cache.with_suffix(".json").exists()
    w2 = window.Window(window_fourier_file=cache, window_configspace_array=dr16["win_LRG"], co=co, accboost=4, windowk=0.1)
    np.testing.assert_array_equal(w2.Wal, w.Wal)
    with pytest.raises(window.MetaInfoError):
        window.Window(window_fourier_file=cache, window_configspace_array=dr16["win_LRG"], co=co, accboost=4, bias=-1.5)
    with pytest.raises(ValueError):
        window.Window(co=co)


def test_binning_and_chained_operators(golden2):
    co = pybird.Common(Nl=3)
    b = binning.Binning(golden2["kout"], co=co)
    assert b.nbin == golden2["kout"].size == 18
    got = np.einsum("bk,nlik->nlib", b.matrix, golden2["win_Ploopl"])
    assert rowmax_rel(got, golden2["bin_Ploopl"]) <= 1e-9
    with pytest.raises(ValueError):
        binning.Binning(golden2["kout"], co=co, kstart=0.0)  # binning.py:89-90
    assert chained.chain_coeff(0) == pytest.approx(-0.4) and chained.chain_coeff(2) == pytest.approx(-20 / 27)
    with pytest.raises(NotImplementedError):
        chained.Chained().chained_matrix(5)


def test_likelihood_host_helpers(dr16):
    k = dr16["NGC_LRG_P"][:, 0]
    m = likelihood.parse_kmask(k, [0, 2, 4], 0.02, 0.20)
    assert all(s.stop - s.start == 18 for s in m.values())
    with pytest.raises(ValueError):
        likelihood.parse_kmask(k, [0, 2], [0.02], 0.2)
    cov = dr16["cov_NGC_L024E02X024_PQP"]
    kq = dr16["NGC_ELG_Q"][:, 0]
    masked = likelihood.mask_covariance(cov, [0, 2, 4], [0, 2, 4], k, 0.02, 0.2, [0, 2], [0, 2], kq, 0.03, 0.2,
                                        [0, 2, 4], [0, 2, 4], k, 0.02, 0.2)
    assert masked.shape == (142, 142)  # DR16 NGC production data vector
    with pytest.raises(ValueError):
        likelihood.mask_covariance(cov[:10, :10], [0], [0], k, None, None)
    assert likelihood.hartlap(1000, 142) == pytest.approx(0.8568568568568569)
    info = likelihood.MultipoleInfo.load(dr16["NGC_ELG_Q"], ls=[0, 2], kmin=0.03, kmax=0.20, symbol="Q")
    assert info.data_vector.size == 34 and info.kout.size == 17


def test_prior_handling():
    class M(marginal.Marginalizable):
        def marginalizable_params(self):
            return ["a", "b", "c"]

    m = M()
    m.setup_prior({"c": {"scale": 2}, "a": {"loc": 1, "scale": 4}})
    assert list(m.valid_prior) == ["a", "c"]  # sorted by marginalizable order (marginal.py:222-226)
    np.testing.assert_allclose(m.sigma_inv, np.diag([1 / 16, 1 / 4]))
    np.testing.assert_allclose(m.mu_G, [1, 0])
    m.setup_prior({"a": None, "b": {"scale": None}})
    assert not m.sigma_inv.any()  # infinite scales -> zero precision (marginal.py:74-75)
    with pytest.raises(marginal.LoggedError):
        m.setup_prior({"zzz": None})
    with pytest.raises(marginal.LoggedError):
        m.setup_prior({"a": {"scale": 1}, "b": None})  # marginal.py:227-231


def test_basis_names_and_descriptors():
    w = parambasis.WestCoastBasis(prefix="X_", cross_prefix=["L_", "E_"])
    assert w.non_gaussian_params() == ["L_b1", "L_b2", "L_b4", "E_b1", "E_b2", "E_b4"]
    assert w.gaussian_params()[-3:] == ["X_ce0", "X_cemono", "X_cequad"]
    co = pybird.Common(Nl=3, kmA=0.7, krA=0.25, ndA=4.5e-5, kmB=0.7, krB=0.25, ndB=2.3e-4)
    d = w.gaussian_descriptors(co)
    assert set(d) == set(w.gaussian_params())
    assert d["L_b3"][1][1] == parambasis.VAR_B1B and d["E_b3"][1][1] == parambasis.VAR_B1A
    with pytest.raises(NotImplementedError):
        parambasis.EastCoastBasis(prefix="a", cross_prefix=["b", "c"])
    assert parambasis.find_param_basis("eastcoast") is parambasis.EastCoastBasis


def test_linear_power_file_extractor_matches_reference():
    """boltzmann.py:246-309 mirror against the live reference's `LinearPowerFile.Pkh` (tests/golden/linear_power_file.npz,
    generated by tests/golden/make_golden_interp.py) + the batched protocol getters."""
    import json
    import os

    from conftest import GOLDEN
    from eftpipe_b200 import boltzmann

    g = dict(np.load(os.path.join(GOLDEN, "linear_power_file.npz")))
    ext = boltzmann.LinearPowerFile((g["k"], g["pk"]), gz=float(g["gz"]), prefix="t_")
    assert sorted(ext.get_requirements()) == json.loads(str(g["requirements"]))
    ext.initialize_with_provider({"t_f": [0.7, 0.8, 0.9], "t_alperp": [1.0, 1.01, 0.99], "t_alpara": [1.0, 0.98, 1.02]})
    pkh = ext.Pkh(g["kh"])
    assert pkh.shape == (3, 200)
    np.testing.assert_allclose(pkh[1], g["pkh"], rtol=1e-13)
    assert ext.DA() == 1 and ext.H() == 1  # the first calls seed the AP fiducial (boltzmann.py:287-297)
    np.testing.assert_allclose(ext.DA(), [1.0, 1.01, 0.99])
    np.testing.assert_allclose(ext.H(), 1 / np.array([1.0, 0.98, 1.02]))
    c = ext.cosmo()
    assert set(c) >= {"pkh", "f", "DA", "H"} and c["pkh"].shape == (3, 200)
    arr = boltzmann.ArrayExtractor(pkh, c["f"], c["DA"], c["H"])
    assert arr.cosmo()["pkh"] is pkh and boltzmann.find_boltzmann_extractor(arr) is arr
    with pytest.raises(NotImplementedError):
        boltzmann.find_boltzmann_extractor("classynu")


def test_custom_window_plugins_are_probed_into_operators(golden2):
    """theory.py:62-72 seam: a class found by dotted path with an in-place `.Window(bird)` becomes a fixed operator."""
    import helpers
    from eftpipe_b200 import plugins, pybird

    co = pybird.Common(Nl=3)
    Weff = golden2["Weff_LRG"]
    picc = -golden2["PSN"] * float(golden2["Pshot"])
    cls = plugins.find_window_constructor("helpers.MatrixWindow")
    assert cls is helpers.MatrixWindow and plugins.find_window_constructor("auto").__name__ == "Window"
    op = plugins.probe_linear_stage(cls(0.95 * Weff, picc=picc, co=co).Window, co)
    assert np.array_equal(op["matrix"], 0.95 * Weff) and op["matrix_st"] is None
    assert np.array_equal(op["picc"], picc)
    op = plugins.probe_linear_stage(cls(Weff, window_st=False, co=co).Window, co)  # stochastic terms left alone
    assert np.array_equal(op["matrix_st"].reshape(150, 150), np.eye(150))
    with pytest.raises(ValueError):
        plugins.probe_linear_stage(helpers.SquaringWindow(Weff, co=co).Window, co)


def test_reference_window_class_as_a_plugin(golden2):
    """The live reference's own Window class, driven as a plugin: probing reproduces the effective operator the
    reference applies (golden Weff_LRG).  Needs /root/reference (build container only)."""
    import os

    import refload
    from eftpipe_b200 import plugins

    if not refload.available():
        pytest.skip("reference tree not mounted")
    ref = refload.load()
    co = ref.pybird.Common(Nl=3, No=3, kmax=0.3, kmA=0.7, krA=0.25, ndA=4.5e-5)
    win = ref.window.Window(window_configspace_file=os.path.join(refload.REFERENCE_ROOT, "data", "DR16_noric", "win_NGC_LRG.txt"),
                            co=co, accboost=4, windowk=0.1, load=False, save=False)
    op = plugins.probe_linear_stage(win.Window, co)
    assert rowmax_rel(op["matrix"].reshape(150, 150), golden2["Weff_LRG"].reshape(150, 150)) <= 1e-12
    assert op["matrix_st"] is None and not op["picc"].any()


def test_plugin_configuration_is_validated_like_the_reference_initializer():
    """tools.py:176-205: unknown keyword / missing positional argument in a plugin's yaml sub-dict -> LoggedError"""
    from eftpipe_b200 import pybird, theory
    from eftpipe_b200.marginal import LoggedError

    co = pybird.Common(Nl=2)
    with pytest.raises(LoggedError, match="does not have keyword"):
        theory._construct(pybird.APeffect, dict(Om_AP=0.3, z_AP=0.5, nbins_mu=100), co=co)
    with pytest.raises(LoggedError, match="missing positional argument"):
        theory._construct(pybird.FiberCollision, dict(fs=0.6), co=co)
    ap = theory._construct(pybird.APeffect, dict(Om_AP=0.3, z_AP=0.5), co=co)
    assert ap.nbinsmu == 200


def test_likelihood_data_and_covariance_readers(tmp_path):
    """likelihood.py:26-62, :241-252, :337-347: yaml-style `path` / `reader` / `reader_kwargs`, multipole symbol and ells
    from the header, block-diagonal covariance from a list of paths."""
    import os

    from eftpipe_b200 import likelihood as lk

    d = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "eftpipe_b200", "data", "dr16_ngc.npz")))
    pq = tmp_path / "NGC_ELG_Q.txt"
    np.savetxt(pq, d["NGC_ELG_Q"], header="k Q0 Q2\nPshot=123.4", comments="# ")
    m = lk.MultipoleInfo.load(path=str(pq), ls=[0, 2], kmin=0.03, kmax=0.20)
    ref = lk.MultipoleInfo.load(table=d["NGC_ELG_Q"], ls=[0, 2], kmin=0.03, kmax=0.20, symbol="Q")
    assert m.symbol == "Q" and m.ls_tot == [0, 2] and np.array_equal(m.data_vector, ref.data_vector)
    assert lk.extract_multipole_info(["k", "P0", "P4", "P2"]) == ("P", [0, 2, 4])
    with pytest.raises(ValueError):
        lk.extract_multipole_info(["k", "P0", "Q2"])
    nohdr = tmp_path / "plain.txt"
    np.savetxt(nohdr, d["NGC_LRG_P"])
    assert lk.MultipoleInfo.load(path=str(nohdr), ls=[0, 2, 4], kmin=0.02, kmax=0.2).ls_tot == [0, 2, 4]
    # covariance: list of paths -> block diagonal; custom reader by dotted path
    c = d["cov_NGC_L024_P"]
    n = c.shape[0]
    pa, pb = tmp_path / "a.txt", tmp_path / "b.txt"
    np.savetxt(pa, c)
    np.savetxt(pb, 2.0 * c)
    like = lk.EFTLike(tracers=["LRG"], data=dict(path=str(nohdr), ls=[0, 2, 4], kmin=0.02, kmax=0.20, reader="numpy.loadtxt"),
                      cov=dict(path=str(pa), reader="numpy.loadtxt", Nreal=1000))
    like2 = lk.EFTLike(tracers=["LRG"], data=dict(table=d["NGC_LRG_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20), cov=dict(matrix=c, Nreal=1000))
    assert np.array_equal(like.invcov, like2.invcov) and np.array_equal(like.data_vector, like2.data_vector)
    two = lk.EFTLike(tracers=["A", "B"], data={t: dict(table=d["NGC_LRG_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20) for t in ("A", "B")},
                     cov=dict(path=[str(pa), str(pb)], Nreal=1000))
    nd = like2.ndata
    assert two.ndata == 2 * nd and np.allclose(two.invcov[nd:, nd:] * 2.0, two.invcov[:nd, :nd]) and not two.invcov[:nd, nd:].any()


def test_cobaya_params_block_is_resolved_for_a_batch():
    """the yaml's `params:` block: fixed values and `value: 'lambda ...'` inputs (b2, b4 from c2, c4) for (B,) arrays"""
    import yaml

    from eftpipe_b200 import cobaya_info

    info = yaml.safe_load("""
params:
  LRG_NGC_b1: {prior: {min: 0, max: 4}, ref: 2.1}
  LRG_NGC_c2: {prior: {min: -100, max: 100}, drop: true}
  LRG_NGC_c4: {value: 0, drop: true}
  LRG_NGC_b2: {value: 'lambda LRG_NGC_c2, LRG_NGC_c4: (LRG_NGC_c2 + LRG_NGC_c4) / np.sqrt(2.)'}
  LRG_NGC_b4: {value: 'lambda LRG_NGC_c2, LRG_NGC_c4: (LRG_NGC_c2 - LRG_NGC_c4) / np.sqrt(2.)'}
  S8: {derived: 'lambda omegam, sigma8: sigma8*np.sqrt(omegam/0.3)'}
""")
    c2 = np.array([0.3, 0.5, -0.2])
    p = cobaya_info.resolve_params(info, {"LRG_NGC_b1": [2.0, 2.1, 2.2], "LRG_NGC_c2": c2})
    np.testing.assert_allclose(p["LRG_NGC_b2"], c2 / np.sqrt(2.0))
    np.testing.assert_allclose(p["LRG_NGC_b4"], c2 / np.sqrt(2.0))
    assert float(p["LRG_NGC_c4"]) == 0.0 and "S8" not in p


CALLABLE_PRIOR = {
    "b3": {"loc": "lambda b1: 0.5 * b1", "scale": 4.0},
    "cct": {"loc": 0.0, "scale": "lambda b1, b2: 2.0 + np.abs(b2)"},
    "cr1": {"loc": "lambda b2, b4: b2 - b4", "scale": "lambda b4: np.exp(0.1 * b4) * 4"},
    "cr2": {"loc": 1.5, "scale": 4.0},
    "ce0": {"loc": 0, "scale": 2.0},
    "cequad": {"scale": 2.0},
}


def test_callable_priors_per_point_match_the_reference_evaluation():
    """marginal.py:13-20, :60-77: string `loc` / `scale` are eval'ed against the sampled EFT parameters at every
    evaluation.  The mirror evaluates them once per batch with arrays (`point_priors`); row by row this must equal the
    oracle's per-point restatement and, where the reference tree is mounted, the reference's own `mu_G` / `sigma_inv`."""
    import pybird_oracle as orc
    import refload
    from eftpipe_b200 import marginal

    names = list(CALLABLE_PRIOR)

    class M(marginal.Marginalizable):
        def marginalizable_params(self):
            return names

    m = M()
    m.setup_prior(CALLABLE_PRIOR)
    assert m.has_callable_prior()
    with pytest.raises(TypeError):
        m.mu_G
    rng = np.random.default_rng(5)
    B = 7
    env = {"b1": rng.normal(2.0, 0.3, B), "b2": rng.normal(0.5, 0.5, B), "b4": rng.normal(0.0, 0.5, B)}
    loc, sinv = m.point_priors(env, B)
    assert loc.shape == sinv.shape == (B, len(names))
    ref_cls = None
    if refload.available():
        ref = refload.load()

        class R(ref.marginal.Marginalizable):
            def marginalizable_params(self):
                return names

            def env(self):
                return self._env

        ref_cls = R()
        ref_cls.valid_prior = ref.marginal.Marginalizable.update_prior(ref_cls, CALLABLE_PRIOR)
        ref_cls._sigma_inv = np.zeros((len(names), len(names)))
    for i in range(B):
        point = {"np": np, **{k: float(v[i]) for k, v in env.items()}}
        mu_o, sig_o = orc.prior_mu_sigma_inv(m.valid_prior, point)
        np.testing.assert_allclose(loc[i], mu_o, rtol=1e-15)
        np.testing.assert_allclose(np.diag(sinv[i]), sig_o, rtol=1e-15)
        if ref_cls is not None:
            ref_cls._env = point
            np.testing.assert_allclose(loc[i], ref_cls.mu_G, rtol=1e-15)
            np.testing.assert_allclose(np.diag(sinv[i]), ref_cls.sigma_inv, rtol=1e-15)
    # one infinite scale at a point switches the whole prior off there (marginal.py:74-75)
    m2 = M()
    m2.valid_prior = {"b3": {"loc": 0, "scale": "lambda b1: np.where(b1 > 2.0, np.inf, 3.0)"}, "cct": {"loc": 1.0, "scale": 2.0}}
    _, s2 = m2.point_priors(env, B)
    off = env["b1"] > 2.0
    assert off.any() and (~off).any()
    assert not s2[off].any() and np.allclose(s2[~off], [1 / 9.0, 0.25])


def test_tracer_prefix_defaults_follow_the_reference():
    """theory.py:285-291: an absent `prefix` means `<tracer>_`, also for the parents of a cross tracer (ADVICE round 1:
    the first version read `LRG_b1` etc. as absent and evaluated with b1 = 0, without an error)"""
    from eftpipe_b200 import theory

    tr = {"LRG": dict(z=0.7, km=0.7, kr=0.25, nd=4.5e-5), "ELG": dict(z=0.85, km=0.7, kr=0.25, nd=2e-4, prefix=""),
          "X": dict(z=0.77, cross=["LRG", "ELG"])}
    th = theory.EFTLSS(tr)
    assert th.build_basis("LRG").prefix == "LRG_" and th.build_basis("ELG").prefix == ""
    x = th.build_basis("X")
    assert x.prefix == "X_" and x.cross_prefix == ["LRG_", "ELG_"]  # an EMPTY parent prefix falls back too (theory.py:290-291)
    assert x.non_gaussian_params()[:3] == ["LRG_b1", "LRG_b2", "LRG_b4"]
    assert theory.tracer_prefix("A", {"prefix": "p_"}) == "p_"
    # with_icc / provider_kwargs are reference keys (theory.py:341-344, :384-388): accepted; icc on a cross is refused at build
    theory.EFTLSS({"A": dict(z=0.7, km=0.7, nd=1e-4, with_icc=True, icc={}, provider_kwargs={})})
    with pytest.raises(marginal.LoggedError):
        theory.EFTLSS({"A": dict(z=0.7, km=0.7, nd=1e-4, bogus_key=1)})


def _window_matrix_case():
    """seeded inputs of tests/golden/make_golden_windowmatrix.py + the reference's outputs"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("mkwm", os.path.join(os.path.dirname(__file__), "golden", "make_golden_windowmatrix.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "window_matrix.npz")))
    assert int(g["seed"]) == mk.SEED
    return mk, g


def test_window_matrix_stage_against_reference_golden():
    """window.py:418-577: the band-power window matrix (cut out of a flat file matrix, cubic interpolation onto the band grid,
    contraction) on a numpy bird-like, against what the unmodified reference returned for the same seeded inputs"""
    import types

    from eftpipe_b200 import pybird, window

    mk, g = _window_matrix_case()
    co = pybird.Common(Nl=3, kmax=0.3, with_NNLO=True)
    flat, terms = mk.inputs(co.Nl, co.Nk)
    ells, kmin, kmax = [int(x) for x in g["ells"]], float(g["kmin"]), float(g["kmax"])
    cut = window.to_window_matrix(flat, window.PInfo((0, 2, 4), 0, 0.4, 400), window.PInfo((0, 1, 2, 3, 4), 0, 0.4, 40),
                                  ells_in=(0, 2, 4), kmax_in=co.k.max(), ells_out=tuple(ells), kmin_out=kmin, kmax_out=kmax)
    assert cut.shape == tuple(g["cut_shape"])
    assert np.array_equal(cut[:, :, ::4, ::25], g["cut_sub"])
    np.testing.assert_allclose(cut.sum(axis=3), g["cut_rowsum"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(cut.sum(axis=2), g["cut_colsum"], rtol=0, atol=1e-12)
    for st in (False, True):
        wm = window.WindowMatrix(cut, window.PolesInfo(co.Nl, 0, co.k.max(), cut.shape[3]), window.PolesInfo(len(ells), kmin, kmax, cut.shape[2]),
                                 co=co, window_st=st)
        bird = types.SimpleNamespace(co=co, create_snapshot=lambda name: None, **{k: v.copy() for k, v in terms.items()})
        wm.Window(bird)
        for n in ("P11l", "Pctl", "Ploopl", "Pstl", "PctNNLOl", "Picc"):
            want = g[f"st{int(st)}.{n}"]
            assert np.shape(getattr(bird, n)) == want.shape, (st, n)
            np.testing.assert_allclose(getattr(bird, n), want, rtol=1e-11, atol=1e-11 * np.abs(want).max(), err_msg=f"{st} {n}")
    # the reference's validation errors (window.py:490-506)
    with pytest.raises(ValueError, match="matrix shape"):
        window.WindowMatrix(cut[:, :, :-1], window.PolesInfo(3, 0, 0.3, 300), window.PolesInfo(3, kmin, kmax, cut.shape[2]), co=co)
    with pytest.raises(ValueError, match="input poles"):
        window.WindowMatrix(cut[:, :2], window.PolesInfo(2, 0, 0.3, 300), window.PolesInfo(3, kmin, kmax, cut.shape[2]), co=co)
    with pytest.raises(NotImplementedError):
        window.WindowMatrix(cut, window.PolesInfo(3, 0, 0.3, 300), window.PolesInfo(3, kmin, kmax, cut.shape[2]), co=co, icc=object())
