"""GPU parity on the other BASELINE configurations and on edge cases of the AP / resummation kernels, against
the oracle on seeded inputs (sizes the oracle finishes in seconds):

  config 5   NFFT=512 loop matrices, kmax=0.4 (Nk=84), fine binning
  NNLO       with_NNLO=True (PctNNLOl rows, CctNNLO resummation)
  Nl=2 + AP  the 2-multipole instantiations of the resum / AP kernels
  extreme AP distortions of +-20-30 % (wide B-spline windows, both directions of k'(mu)) and the exact fiducial
             (k' == k: degenerate window)

Tolerance as in test_gpu_parity.py: 1e-8 of the row maximum."""
import numpy as np
import pytest

from conftest import rowmax_rel

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _np(t):
    return t.detach().cpu().numpy()


def _oracle_terms(orc, co, nl, rs, batch, i, ap=None, DA=None, H=None):
    b = orc.Bird(co, batch.kin, batch.plin[i], batch.f[i], DA=None if DA is None else DA[i], H=None if H is None else H[i])
    nl.PsCf(b)
    orc.set_PsCfl(b)
    if rs is not None:
        rs.Ps(b)
    if ap is not None:
        ap.AP(b)
    parts = [b.P11l, b.Pctl, b.Ploopl, b.Pstl]
    if co.with_NNLO:
        parts.append(b.PctNNLOl)
    return np.concatenate(parts, axis=1)


def test_config5_high_resolution():
    """BASELINE config 5: NFFT=512, kmax=0.4 -> Nk=84, + fine k-binning through the projection operator."""
    import torch

    import pybird_oracle as orc
    from eftpipe_b200 import engine, plan, synthetic

    batch = synthetic.make_batch(3, 0.7, seed=512, unique=3)
    g = plan.GridConfig(Nl=3, kmax=0.4, NFFT=512)
    kout = np.arange(0.0125, 0.39, 0.005)
    binm, keff, _, _ = plan.binning_matrix(g.k, kout, decimals=3)
    proj = plan.compose_projection(g, binning=binm)
    dp_raw = engine.DevicePlan(plan.build_tracer_plan(Nl=3, kmax=0.4, NFFT=512))
    dp_bin = engine.DevicePlan(plan.build_tracer_plan(Nl=3, kmax=0.4, NFFT=512, projection=proj))
    raw, _ = dp_raw.eval_terms(batch.plin, batch.f)
    binned, _ = dp_bin.eval_terms(batch.plin, batch.f)
    torch.cuda.synchronize()
    raw, binned = _np(raw), _np(binned)
    co = orc.Common(Nl=3, kmax=0.4)
    assert co.Nk == 84
    nl, rs = orc.NonLinear(co, NFFT=512), orc.Resum(co)
    ob = orc.Binning(kout, co, decimals=3)
    for i in (0, 2):
        ref = _oracle_terms(orc, co, nl, rs, batch, i)
        assert rowmax_rel(raw[i], ref) <= TOL, i
        refb = np.array([[ob.integrate(ref[l, t]) for t in range(ref.shape[1])] for l in range(3)])
        assert rowmax_rel(binned[i], refb) <= TOL, i


def test_nnlo_terms():
    import torch

    import pybird_oracle as orc
    from eftpipe_b200 import engine, plan, synthetic

    batch = synthetic.make_batch(4, 0.7, seed=77, unique=4)
    dp = engine.DevicePlan(plan.build_tracer_plan(Nl=3, with_NNLO=True))
    got, _ = dp.eval_terms(batch.plin, batch.f)
    torch.cuda.synchronize()
    got = _np(got)
    assert got.shape[2] == 27
    co = orc.Common(Nl=3, with_NNLO=True)
    nl, rs = orc.NonLinear(co), orc.Resum(co)
    for i in (1, 3):
        assert rowmax_rel(got[i], _oracle_terms(orc, co, nl, rs, batch, i)) <= TOL, i


@pytest.mark.parametrize("Nl", [2, 3])
def test_ap_extreme_and_degenerate_geometry(Nl):
    """q_perp, q_par far from 1 in both orders (k' increasing / decreasing with mu, windows of ~10-20 B-splines)
    and the exact fiducial cosmology (k' == k for every mu)."""
    import torch

    import pybird_oracle as orc
    from eftpipe_b200 import engine, plan, synthetic

    Om_AP, z_AP = 0.307115, 0.696
    DA0, H0 = synthetic.angular_distance(Om_AP, z_AP), synthetic.hubble(Om_AP, z_AP)
    batch = synthetic.make_batch(6, 0.7, seed=11, unique=6)
    DA = DA0 * np.array([1.0, 1.25, 0.8, 1.1, 0.92, 1.0])
    H = H0 * np.array([1.0, 0.8, 1.3, 1.1, 1.0, 0.9])
    for APst in (False, True):
        dp = engine.DevicePlan(plan.build_tracer_plan(Nl=Nl, ap=dict(DA=DA0, H=H0, APst=APst)))
        got, _ = dp.eval_terms(batch.plin, batch.f, DA, H)
        torch.cuda.synchronize()
        got = _np(got)
        assert np.isfinite(got).all()
        co = orc.Common(Nl=Nl)
        nl, rs = orc.NonLinear(co), orc.Resum(co)
        ap = orc.APeffect(co, DA=DA0, H=H0, APst=APst)
        for i in range(6):
            ref = _oracle_terms(orc, co, nl, rs, batch, i, ap=ap, DA=DA, H=H)
            assert rowmax_rel(got[i], ref) <= TOL, (Nl, APst, i)


def test_large_batch_is_consistent_with_small_batches():
    """Size-independent property at 4x BASELINE's batch (4096 points, several AP chunks): every point of a
    large batch equals the same point evaluated in a batch of 32 (no cross-talk between lanes, CTAs, chunks)."""
    import torch

    from eftpipe_b200 import engine, plan, synthetic

    Om_AP, z_AP = 0.307115, 0.696
    DA0, H0 = synthetic.angular_distance(Om_AP, z_AP), synthetic.hubble(Om_AP, z_AP)
    B = 4096
    batch = synthetic.make_batch(B, 0.7, seed=5, unique=64)
    dp = engine.DevicePlan(plan.build_tracer_plan(Nl=3, ap=dict(DA=DA0, H=H0, APst=True)))
    big, _ = dp.eval_terms(batch.plin, batch.f, batch.DA, batch.H)
    torch.cuda.synchronize()
    big = _np(big).copy()
    assert np.isfinite(big).all()
    for lo in (0, 1504, 4064):
        sl = slice(lo, lo + 32)
        small, _ = dp.eval_terms(batch.plin[sl], batch.f[sl], batch.DA[sl], batch.H[sl])
        torch.cuda.synchronize()
        np.testing.assert_array_equal(_np(small), big[sl])


def test_batch_beyond_one_ap_chunk():
    """12288 points are more than one launch of the AP kernels holds (the dense overflow operator is capped at 2 GB =
    11930 cosmologies of 50 nodes): two chunks, the second one with the batch padding; points of both chunks and across
    the chunk boundary equal the same points evaluated in batches of 32"""
    import torch

    from eftpipe_b200 import engine, plan, synthetic

    Om_AP, z_AP = 0.307115, 0.696
    DA0, H0 = synthetic.angular_distance(Om_AP, z_AP), synthetic.hubble(Om_AP, z_AP)
    B = 12288 - 7  # ragged: the last chunk ends inside a group of 32
    batch = synthetic.make_batch(B, 0.7, seed=6, unique=64)
    for APst in (True, False):
        dp = engine.DevicePlan(plan.build_tracer_plan(Nl=3, ap=dict(DA=DA0, H=H0, APst=APst)))
        big, _ = dp.eval_terms(batch.plin, batch.f, batch.DA, batch.H)
        torch.cuda.synchronize()
        big = _np(big).copy()
        assert np.isfinite(big).all()
        for lo in (0, 11914, B - 32):
            sl = slice(lo, lo + 32)
            small, _ = dp.eval_terms(batch.plin[sl], batch.f[sl], batch.DA[sl], batch.H[sl])
            torch.cuda.synchronize()
            np.testing.assert_array_equal(_np(small), big[sl])
        del dp, big
        torch.cuda.empty_cache()


def test_odd_node_count_with_nnlo():
    """kmax = 0.305 gives 65 k nodes; with NNLO (27 term rows) the per-cosmology coefficient block has an odd number of
    doubles - the TMA-staged AP kernels pad the stride.  Fused pipeline against the oracle."""
    import pybird_oracle as orc
    from eftpipe_b200 import engine, plan, synthetic

    Om_AP, z_AP = 0.307115, 0.696
    DA0, H0 = synthetic.angular_distance(Om_AP, z_AP), synthetic.hubble(Om_AP, z_AP)
    batch = synthetic.make_batch(3, 0.7, seed=21)
    pl = plan.build_tracer_plan(Nl=3, kmax=0.305, with_NNLO=True, ap=dict(DA=DA0, H=H0, APst=True))
    assert pl.grid.Nk == 65 and pl.grid.nterm == 27
    got, _ = engine.DevicePlan(pl).eval_terms(batch.plin, batch.f, batch.DA, batch.H)
    got = _np(got)
    co = orc.Common(Nl=3, kmax=0.305, with_NNLO=True)
    nl, rs, ap = orc.NonLinear(co), orc.Resum(co), orc.APeffect(co, Om_AP=Om_AP, z_AP=z_AP, APst=True)
    for i in range(2):
        b = orc.Bird(co, batch.kin, batch.plin[i], batch.f[i], batch.DA[i], batch.H[i], 0.7)
        nl.PsCf(b); orc.set_PsCfl(b); rs.Ps(b); ap.AP(b)
        ref = np.concatenate([b.P11l, b.Pctl, b.Ploopl, b.Pstl, b.PctNNLOl], axis=1)
        assert rowmax_rel(got[i], ref) <= 1e-8


_SCALAR_SCRIPT = r"""
import sys
import numpy as np
from eftpipe_b200 import engine, plan, synthetic
out = sys.argv[1]
batch = synthetic.make_batch(5, 0.7, seed=77)
res = {}
for name, kw in (("nl3", dict(Nl=3)), ("nl2", dict(Nl=2)), ("nnlo", dict(Nl=3, with_NNLO=True)), ("opti", dict(Nl=3, optiresum=True))):
    got, _ = engine.DevicePlan(plan.build_tracer_plan(**kw)).eval_terms(batch.plin, batch.f)
    res[name] = got.detach().cpu().numpy()
np.savez(out, **res)
"""


def test_resum_scalar_fallback_agrees_with_dmma_form(tmp_path):
    """The a = 1 half of the resummation kernel has two forms: the DMMA row contraction (default) and the scalar sweep
    that remains as the fallback for unaligned / odd row sizes (and sweeps the left-over columns).  Same inputs through
    both (the fallback forced by EFTB_RESUM_DOTS=scalar in a fresh process, the knob is read once per process):
    Nl = 3, Nl = 2, NNLO (third n-tile) and optiresum (52 s nodes padded to 64)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "run.py"
    script.write_text(_SCALAR_SCRIPT)
    outs = {}
    for mode in ("dmma", "scalar"):
        env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        env.pop("EFTB_RESUM_DOTS", None)
        if mode == "scalar":
            env["EFTB_RESUM_DOTS"] = "scalar"
        out = tmp_path / f"{mode}.npz"
        subprocess.run([sys.executable, str(script), str(out)], check=True, env=env, cwd=root, timeout=600)
        outs[mode] = np.load(out)
    for name in outs["dmma"].files:
        a, b = outs["dmma"][name], outs["scalar"][name]
        assert np.isfinite(a).all() and a.shape == b.shape
        assert rowmax_rel(a, b) <= 1e-11, name  # same algebra, different summation order


@pytest.mark.parametrize("kmax,NFFT", [(0.4, 256), (0.35, 512)])
def test_ap_on_the_larger_k_grids(kmax, NFFT):
    """AP resampling on the Nk = 84 / 74 node grids (kmax = 0.4 / 0.35) with both loop-matrix sizes: larger operator and
    coefficient blocks per cosmology in ap_apply_kernel, the 8-warp anti-diagonal CTA (NFFT = 512), 77 / 67 resummed nodes"""
    import torch

    import pybird_oracle as orc
    from eftpipe_b200 import engine, plan, synthetic

    Om_AP, z_AP = 0.307115, 0.696
    DA0, H0 = synthetic.angular_distance(Om_AP, z_AP), synthetic.hubble(Om_AP, z_AP)
    batch = synthetic.make_batch(5, 0.7, seed=23, unique=5)
    DA = DA0 * np.array([1.0, 1.06, 0.95, 1.02, 0.9])
    H = H0 * np.array([1.0, 0.97, 1.04, 1.1, 1.0])
    dp = engine.DevicePlan(plan.build_tracer_plan(Nl=3, kmax=kmax, NFFT=NFFT, ap=dict(DA=DA0, H=H0, APst=True)))
    got, _ = dp.eval_terms(batch.plin, batch.f, DA, H)
    torch.cuda.synchronize()
    got = _np(got)
    co = orc.Common(Nl=3, kmax=kmax)
    assert got.shape[-1] == co.Nk and np.isfinite(got).all()
    nl, rs = orc.NonLinear(co, NFFT=NFFT), orc.Resum(co)
    ap = orc.APeffect(co, DA=DA0, H=H0, APst=True)
    for i in (0, 1, 4):
        assert rowmax_rel(got[i], _oracle_terms(orc, co, nl, rs, batch, i, ap=ap, DA=DA, H=H)) <= TOL, (kmax, NFFT, i)
