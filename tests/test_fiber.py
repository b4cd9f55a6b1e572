"""Fibre-collision correction (FiberCollision.fibcolWindow, pybird.py:1631-1809): the oracle and
the plan's fixed operator against the live-reference golden, the composition order with window / binning, and the
CUDA path (fused pipeline and the reference-named mirror class) against the oracle."""
import json

import numpy as np
import pytest

from conftest import rowmax_rel

TOL = 1e-8


def test_oracle_matches_reference_golden(fiber_kat):
    import pybird_oracle as orc

    kw = json.loads(str(fiber_kat["fiber"]))
    got = orc.fiber_dPcorr(orc.Common(Nl=3), fiber_kat["PS"], **kw)
    assert rowmax_rel(got, fiber_kat["dPcorr"]) <= 1e-13


def test_plan_operator_matches_reference_golden(fiber_kat):
    from eftpipe_b200 import plan as P

    kw = json.loads(str(fiber_kat["fiber"]))
    g = P.GridConfig(Nl=3)
    F = P.fiber_matrix(g.k, 3, kw["fs"], kw["Dfc"], kw["ktrust"])
    got = np.einsum("aklm,ljm->ajk", F, fiber_kat["PS"])
    assert rowmax_rel(got, fiber_kat["dPcorr"]) <= 1e-13


def test_composition_order_and_stochastic_exemption(fiber_kat):
    """window -> fibre -> binning (theory.py:584-604); the stochastic rows skip the fibre operator unless fiberst."""
    from eftpipe_b200 import plan as P

    kw = json.loads(str(fiber_kat["fiber"]))
    g = P.GridConfig(Nl=3)
    rng = np.random.default_rng(3)
    W = np.eye(150).reshape(3, 50, 3, 50) + 0.01 * rng.normal(size=(3, 50, 3, 50))
    F = P.fiber_matrix(g.k, 3, kw["fs"], kw["Dfc"], kw["ktrust"])
    binm, _, _, _ = P.binning_matrix(g.k, np.arange(0.025, 0.2, 0.01))
    proj = P.compose_projection(g, window=W, fiber=F, binning=binm)
    x = rng.normal(size=(3, 50))
    xw = np.einsum("akln,ln->ak", W, x)
    ref = (xw + np.einsum("akln,ln->ak", F, xw)) @ binm.T
    assert np.abs(proj["matrix"] @ x.reshape(-1) - ref.reshape(-1)).max() <= 1e-12 * np.abs(ref).max()
    ref_st = xw @ binm.T
    assert proj["matrix_st"] is not None
    assert np.abs(proj["matrix_st"] @ x.reshape(-1) - ref_st.reshape(-1)).max() <= 1e-12 * np.abs(ref_st).max()
    assert P.compose_projection(g, window=W, fiber=F, binning=binm, fiber_st=True)["matrix_st"] is None


def test_window_st_false_keeps_stochastic_terms_out_of_the_window():
    """window.py:401-403: with window_st=False the stochastic terms skip the window but are still binned."""
    from eftpipe_b200 import plan as P

    g = P.GridConfig(Nl=3)
    rng = np.random.default_rng(4)
    W = np.eye(150).reshape(3, 50, 3, 50) + 0.01 * rng.normal(size=(3, 50, 3, 50))
    binm, _, _, _ = P.binning_matrix(g.k, np.arange(0.025, 0.2, 0.01))
    proj = P.compose_projection(g, window=W, binning=binm, window_st=False)
    x = rng.normal(size=(3, 50))
    assert np.abs(proj["matrix_st"] @ x.reshape(-1) - (x @ binm.T).reshape(-1)).max() <= 1e-13 * np.abs(x).max()
    assert P.compose_projection(g, window=W, binning=binm)["matrix_st"] is None


@pytest.mark.gpu
@pytest.mark.parametrize("fiberst", [False, True])
def test_gpu_fibre_collisions_against_oracle(fiber_kat, fiberst):
    import torch

    import pybird_oracle as orc
    from eftpipe_b200 import engine, plan, pybird, synthetic

    kw = json.loads(str(fiber_kat["fiber"]))
    batch = synthetic.make_batch(4, 0.7, seed=21, unique=4)
    g = plan.GridConfig(Nl=3)
    F = plan.fiber_matrix(g.k, 3, kw["fs"], kw["Dfc"], kw["ktrust"])
    proj = plan.compose_projection(g, fiber=F, fiber_st=fiberst)
    dp = engine.DevicePlan(plan.build_tracer_plan(Nl=3, projection=proj))
    got, _ = dp.eval_terms(batch.plin, batch.f)
    # the same through the reference-named classes
    co = pybird.Common(Nl=3)
    nl, rs = pybird.NonLinear(load=False, save=False, co=co), pybird.Resum(co=co)
    fc = pybird.FiberCollision(co=co, fiberst=fiberst, **kw)
    bird = pybird.Bird(batch.kin, batch.plin, batch.f, co=co)
    nl.PsCf(bird)
    bird.setPsCfl()
    rs.Ps(bird)
    fc.fibcolWindow(bird)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    mirror = bird.terms_point_major().cpu().numpy()
    oco = orc.Common(Nl=3)
    onl, ors = orc.NonLinear(oco), orc.Resum(oco)
    for i in (0, 3):
        b = orc.Bird(oco, batch.kin, batch.plin[i], batch.f[i])
        onl.PsCf(b)
        orc.set_PsCfl(b)
        ors.Ps(b)
        orc.fibcol_window(b, fiberst=fiberst, **kw)
        ref = np.concatenate([b.P11l, b.Pctl, b.Ploopl, b.Pstl], axis=1)
        assert rowmax_rel(got[i], ref) <= TOL, i
        assert rowmax_rel(mirror[i], ref) <= TOL, i
