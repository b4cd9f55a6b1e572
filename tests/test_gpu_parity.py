"""GPU parity: every CUDA stage called through the C ABI (ctypes -> libeftb200.so) against the golden
vectors of the live reference and against the oracle on seeded synthetic inputs.

Tolerance (north star): multipoles <= 1e-8 relative, chi^2 <= 1e-6.  "Relative" is measured against the
largest entry of each k-row (rtol plus atol = tol * max|row|) because loop terms cross
zero; all arithmetic is fp64."""
import numpy as np
import pytest

import helpers
from conftest import rowmax_rel

pytestmark = pytest.mark.gpu
TOL = 1e-8


@pytest.fixture(scope="module")
def gpu(golden2):
    import torch

    from eftpipe_b200 import engine

    assert torch.cuda.is_available()
    pl = helpers.config2_plan(golden2)
    dp = engine.DevicePlan(pl)
    return pl, dp


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def stages(gpu, golden2):
    """Run the pipeline stage by stage through the stage-level entry points."""
    import torch

    pl, dp = gpu
    g = golden2
    B = g["plin"].shape[0]
    F = dp.front(g["plin"])
    D = dp.antidiag(F, B)
    P22, Cs = dp.spectral(D, B)
    f_bm = dp.to_batch_minor(g["f"])[0]
    T, Cr = dp.group(F, P22, Cs, f_bm, B)
    T_pre = T.clone()
    dp.resum(F, Cr, f_bm, T, B)
    T_res = T.clone()
    DA_bm, H_bm = dp.to_batch_minor(g["DA"])[0], dp.to_batch_minor(g["H"])[0]
    T_ap = dp.ap(T, DA_bm, H_bm, B)
    out = dp.project(T_ap, B)
    # the fused pipeline's order: Legendre/f grouping of D before the D -> C(s) transform
    P22g, Crg = dp.spectral_grouped(D, f_bm, B)
    Tg, Crg = dp.group(F, P22g, None, f_bm, B, Cr=Crg)
    dp.resum(F, Crg, f_bm, Tg, B)
    outg = dp.project(dp.ap(Tg, DA_bm, H_bm, B), B)
    torch.cuda.synchronize()
    return dict(B=B, F=_np(F), D=_np(D), P22=_np(P22), Cs=_np(Cs), T_pre=_np(T_pre), Cr=_np(Cr), T_res=_np(T_res),
                T_ap=_np(T_ap), out=_np(out), Crg=_np(Crg), outg=_np(outg))


def test_front(gpu, stages, golden2):
    pl, _ = gpu
    g, F, B = golden2, stages["F"], stages["B"]
    rows = lambda name: F[pl.front.rows[name][0] : pl.front.rows[name][0] + pl.front.rows[name][1], :B]
    c = rows("cre") + 1j * rows("cim")
    assert np.abs(c.T - g["coef"][:, :129]).max() <= 1e-12 * np.abs(g["coef"]).max()
    assert rowmax_rel(rows("P11").T, g["P11"]) <= TOL
    assert rowmax_rel(rows("X").T, g["X"]) <= TOL
    assert rowmax_rel(rows("Y").T, g["Y"]) <= TOL
    assert rowmax_rel(rows("C11").reshape(3, 80, B).transpose(2, 0, 1), g["C11"]) <= TOL
    assert rowmax_rel(rows("Cct").reshape(3, 80, B).transpose(2, 0, 1), g["Cct"]) <= TOL


def test_loops(stages, golden2):
    B = stages["B"]
    assert rowmax_rel(stages["P22"][..., :B].transpose(2, 0, 1), golden2["P22"]) <= TOL
    assert rowmax_rel(stages["Cs"][:, :28, :, :B].transpose(3, 0, 1, 2), golden2["C22"]) <= TOL
    assert rowmax_rel(stages["Cs"][:, 28:, :, :B].transpose(3, 0, 1, 2), golden2["C13"]) <= TOL


def test_antidiag_matches_emulation(gpu, stages):
    import emulate as E

    pl, _ = gpu
    B = stages["B"]
    F = stages["F"][:, :B]
    D = E.antidiag(pl, E.rows(pl, F, "cre"), E.rows(pl, F, "cim"))
    got = stages["D"][..., :B]
    assert np.abs(got - D).max() <= 1e-12 * np.abs(D).max()


@pytest.mark.parametrize("stage,prefix", [("T_pre", "pre_"), ("T_res", "res_"), ("T_ap", "ap_")])
def test_terms(stages, golden2, stage, prefix):
    B = stages["B"]
    T = helpers.split_terms(stages[stage][..., :B].transpose(3, 0, 2, 1))
    for name, arr in T.items():
        key = prefix + name
        if key in golden2:
            assert rowmax_rel(arr, golden2[key]) <= TOL, key


def test_cloopl(stages, golden2):
    B = stages["B"]
    assert rowmax_rel(stages["Cr"][:B, :, 2:14, :], golden2["pre_Cloopl"]) <= TOL
    # grouped variant (what the fused pipeline runs): same rows, and C11/Cct identical
    assert rowmax_rel(stages["Crg"][:B, :, 2:14, :], golden2["pre_Cloopl"]) <= TOL
    np.testing.assert_array_equal(stages["Crg"][:B, :, :2, :], stages["Cr"][:B, :, :2, :])


def test_projection(gpu, stages, golden2):
    pl, _ = gpu
    B = stages["B"]
    out = stages["out"][..., :B]
    nl, nk = pl.out_shape
    T = helpers.split_terms(out.reshape(nl, nk, out.shape[1], B).transpose(3, 0, 2, 1))
    for name, arr in T.items():
        assert rowmax_rel(arr, golden2["bin_" + name]) <= TOL, name


def test_fused_pipeline_equals_stages(gpu, stages, golden2):
    import torch

    pl, dp = gpu
    g = golden2
    pm, bm = dp.eval_terms(g["plin"], g["f"], g["DA"], g["H"], want_bm=True)
    torch.cuda.synchronize()
    B = stages["B"]
    np.testing.assert_array_equal(_np(bm)[..., :B], stages["outg"][..., :B])
    assert rowmax_rel(stages["outg"][..., :B], stages["out"][..., :B]) <= 1e-11
    T = helpers.split_terms(_np(pm))
    for name, arr in T.items():
        assert rowmax_rel(arr, g["bin_" + name]) <= TOL, name


def test_ragged_batches_and_padding(gpu, golden2):
    """B not a multiple of 32 / B = 1 / B > 32: pad lanes must not leak into results."""
    import torch

    pl, dp = gpu
    g = golden2
    reps = 37
    plin = np.tile(g["plin"], (reps, 1))[:70]
    f, DA, H = (np.tile(g[k], reps)[:70] for k in ("f", "DA", "H"))
    pm70, _ = dp.eval_terms(plin, f, DA, H)
    pm1, _ = dp.eval_terms(plin[:1], f[:1], DA[:1], H[:1])
    torch.cuda.synchronize()
    a, b = _np(pm70), _np(pm1)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[3], a[0])
    np.testing.assert_array_equal(a[69], a[0])


def test_against_oracle_on_seeded_inputs():
    """Fresh seeded cosmologies (not in the goldens): CUDA path vs the oracle, Nl=3 and Nl=2."""
    import torch

    import pybird_oracle as orc
    from eftpipe_b200 import engine, plan, synthetic

    batch = synthetic.make_batch(5, 0.7, seed=99, unique=5)
    for Nl in (3, 2):
        dp = engine.DevicePlan(plan.build_tracer_plan(Nl=Nl))
        pm, _ = dp.eval_terms(batch.plin, batch.f)
        torch.cuda.synchronize()
        got = _np(pm)
        co = orc.Common(Nl=Nl)
        nl, rs = orc.NonLinear(co), orc.Resum(co)
        for i in (0, 4):
            b = orc.Bird(co, batch.kin, batch.plin[i], batch.f[i])
            nl.PsCf(b)
            orc.set_PsCfl(b)
            rs.Ps(b)
            ref = np.concatenate([b.P11l, b.Pctl, b.Ploopl, b.Pstl], axis=1)
            assert rowmax_rel(got[i], ref) <= TOL, (Nl, i)


def test_linearity_property_full_size():
    """Size-independent property at BASELINE's batch size (1024): without AP the linear and counter terms
    scale linearly with the amplitude of P_lin at fixed shape, the 22/13 loop terms quadratically
    (before resummation mixes them)."""
    import torch

    from eftpipe_b200 import engine, plan, synthetic

    batch = synthetic.make_batch(1024, 0.7, seed=3, unique=8)
    dp = engine.DevicePlan(plan.build_tracer_plan(Nl=3, with_resum=False))
    a, _ = dp.eval_terms(batch.plin, batch.f)
    b, _ = dp.eval_terms(2.0 * batch.plin, batch.f)
    torch.cuda.synchronize()
    a, b = _np(a), _np(b)
    assert np.isfinite(a).all()
    assert rowmax_rel(b[:, :, 0:9], 2.0 * a[:, :, 0:9]) <= 1e-12
    assert rowmax_rel(b[:, :, 9:21], 4.0 * a[:, :, 9:21]) <= 1e-11


def test_error_paths(gpu):
    from eftpipe_b200 import _lib

    pl, dp = gpu
    lib = _lib.load()
    assert lib.eftb_eval_terms(dp.handle, 4, None, None, None, None, None, None, None, 0, None) == -1
    ws = dp.torch.empty(16, dtype=dp.torch.float64, device="cuda")
    x = dp.torch.ones((4, 200), dtype=dp.torch.float64, device="cuda")
    s = dp.torch.ones(4, dtype=dp.torch.float64, device="cuda")
    rc = lib.eftb_eval_terms(dp.handle, 4, x.data_ptr(), s.data_ptr(), s.data_ptr(), s.data_ptr(), None, None, ws.data_ptr(), 128, None)
    assert rc == -4
