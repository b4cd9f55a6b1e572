"""GPU: the input side - the on-device Eisenstein-Hu producer behind the BoltzmannExtractor seam
(boltzmann.py:22-101) against the host model every synthetic input comes from, and the extractor branch of
`EFTLSS.calculate` (theory.py:559-565)."""
import numpy as np
import pytest

from conftest import rowmax_rel

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def test_eisenstein_hu_on_device_matches_host_model():
    from eftpipe_b200 import boltzmann, synthetic

    z = 0.696
    host = synthetic.make_batch(6, z, seed=20261018 + 9)       # scipy.quad growth / distance integrals
    fast = synthetic.make_batch_fast(70, z, seed=20261018 + 9)  # Gauss-Legendre, the same nodes as the kernel
    ex = boltzmann.EisensteinHu(rdrag=synthetic.RDRAG)
    ex.initialize(zeff=z)
    ex.initialize_with_provider(dict(omegam=fast.theta[:, 0], h=fast.theta[:, 1], sigma8=fast.theta[:, 2]))
    ex.calculate()
    c = ex.cosmo()
    assert c["pkh"].is_cuda and tuple(c["pkh"].shape) == (70, 200)
    # transcendental functions of the device math library agree with numpy's to a few ulp; the spectrum spans 6 decades
    assert np.max(np.abs(_np(c["pkh"]) / fast.plin - 1)) <= 1e-11
    for n in ("f", "DA", "H"):
        assert np.max(np.abs(_np(c[n]) / getattr(fast, n) - 1)) <= 1e-12, n
    assert np.max(np.abs(_np(c["pkh"])[:6] / host.plin - 1)) <= 1e-9
    assert np.max(np.abs(_np(c["f"])[:6] / host.f - 1)) <= 1e-9
    np.testing.assert_array_equal(_np(c["h"]), fast.theta[:, 1])


def test_extractor_branch_of_the_theory(golden2):
    """theory.py:559-565: `calculate` pulls Pkh / f / DA / H through an extractor; with the device producer the sampled
    cosmology is the only per-point input.  Same multipole terms as with the host-generated arrays."""
    from eftpipe_b200 import boltzmann, synthetic, theory

    z, B = 0.7, 40
    batch = synthetic.make_batch_fast(B, z, seed=5)
    tr = {"LRG": dict(prefix="", z=z, km=0.7, kr=0.25, nd=4.5e-5, with_IRresum=True, with_APeffect=True,
                      APeffect=dict(Om_AP=0.307115, z_AP=0.696, APst=True))}
    th = theory.EFTLSS(tr).must_provide({"nonlinear_Plk_grid": {"LRG": {"ls": [0, 2, 4], "binned": False}}}).initialize()
    ex = boltzmann.EisensteinHu()
    ex.initialize(zeff=z)
    ex.initialize_with_provider(dict(omegam=batch.theta[:, 0], h=batch.theta[:, 1], sigma8=batch.theta[:, 2]))
    ex.calculate()
    params = dict(b1=2.0, b2=0.5, b4=0.5)
    th.calculate({"LRG": ex})
    _, _, a = th.get_nonlinear_Plk_grid("LRG", params)
    th.calculate({"LRG": boltzmann.ArrayExtractor(batch.plin, batch.f, batch.DA, batch.H)})
    _, _, b = th.get_nonlinear_Plk_grid("LRG", params)
    assert tuple(a.shape) == (B, 3, 50)
    assert rowmax_rel(_np(a), _np(b)) <= 1e-9
    th.calculate({"LRG": dict(pkh=batch.plin, f=batch.f, DA=batch.DA, H=batch.H)})
    _, _, c = th.get_nonlinear_Plk_grid("LRG", params)
    assert np.array_equal(_np(b), _np(c))


def test_sampled_cosmology_through_cobaya(tmp_path):
    """Cobaya samples (omegam, h, sigma8, b1, c2): the kernel helper theories produce P_lin on the device.  logp against the
    same model fed with host-generated tables, B = 16 points per evaluation; a nuisance-only step re-runs no producer."""
    import os
    import sys

    from conftest import ROOT

    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
    import refdriver
    from cobaya.model import get_model

    from eftpipe_b200 import boltzmann, synthetic

    B = 16
    theta = synthetic.draw_cosmologies(B, 20261018 + 3)
    paths = refdriver.write_dr16(os.path.join(str(tmp_path), "dr16"))
    tables = {}
    for name, z in refdriver.TRACERS:
        b = synthetic.make_batch_fast(B, z, seed=20261018 + 3)
        tables[name] = dict(pkh=b.plin, f=b.f, DA=b.DA, H=b.H, h=b.h, rdrag=b.rdrag)
    pts = refdriver.draw_points(B)
    nuis = {k: v for k, v in pts.items() if k != "point"}

    def run(device_producer):
        info = refdriver.config3_info(paths, tables, package="eftpipe_b200")
        info["params"].pop("point")
        tr = info["theory"]["eftpipe_b200.eftlss"]["tracers"]
        for name, z in refdriver.TRACERS:
            if device_producer:
                tr[name].update(provider="eftpipe_b200.boltzmann.EisensteinHu", provider_kwargs=dict(rdrag=synthetic.RDRAG))
            else:
                t = tables[name]
                tr[name].update(provider=boltzmann.ArrayExtractor(t["pkh"], t["f"], t["DA"], t["H"], h=t["h"], rdrag=t["rdrag"]),
                                provider_kwargs={})
        if device_producer:
            for p in ("omegam", "h", "sigma8"):
                info["params"][p] = {"prior": {"min": 0, "max": 2}}
        model = get_model(info)
        point = dict(nuis)
        if device_producer:
            point.update(omegam=theta[:, 0], h=theta[:, 1], sigma8=theta[:, 2])
        return model, point, model.logposterior(point)

    _, _, ref = run(False)
    model, point, got = run(True)
    np.testing.assert_allclose(_np(got.loglikes[0]), _np(ref.loglikes[0]), rtol=1e-8)
    np.testing.assert_allclose(got.derived["LRG_NGC_alperp"], ref.derived["LRG_NGC_alperp"], rtol=1e-11)
    kernels = [c for n, c in model.theory.items() if n.endswith(".kernel")]
    n0 = [c.n_computed for c in kernels]
    point["ELG_NGC_b1"] = point["ELG_NGC_b1"] + 0.02
    model.logposterior(point)
    assert [c.n_computed for c in kernels] == n0
    point["sigma8"] = point["sigma8"] * 1.01
    model.logposterior(point)
    assert [c.n_computed for c in kernels] == [n + 1 for n in n0]


def test_shared_sigma8_between_tracers():
    """the sigma8 normalisation integral is redshift independent: a second tracer's producer reuses the first one's when it
    was just evaluated on the very same parameter tensors - and only then"""
    import torch

    from eftpipe_b200 import boltzmann, synthetic

    theta = synthetic.draw_cosmologies(40, 3)
    dev = [torch.as_tensor(theta[:, i].copy(), device="cuda") for i in range(3)]
    a = boltzmann.EisensteinHu()
    b = boltzmann.EisensteinHu(share_sigma8_with=a)
    ref = boltzmann.EisensteinHu()
    a.initialize(zeff=0.696)
    b.initialize(zeff=0.849)
    ref.initialize(zeff=0.849)
    kw = dict(omegam=dev[0], h=dev[1], sigma8=dev[2])
    a.calculate(**kw)
    b.calculate(**kw)
    ref.calculate(**kw)
    assert b._sig2 is a._sig2                                   # reused
    assert torch.equal(b.cosmo()["pkh"], ref.cosmo()["pkh"]) and torch.equal(b.cosmo()["f"], ref.cosmo()["f"])
    b.calculate(**kw)                                           # the primary was not re-evaluated since: no stale reuse
    assert b._sig2 is not a._sig2 and torch.equal(b.cosmo()["pkh"], ref.cosmo()["pkh"])
    other = dict(kw, sigma8=dev[2] * 1.1)                       # different tensors: no reuse
    a.calculate(**kw)
    b.calculate(**other)
    assert b._sig2 is not a._sig2
    np.testing.assert_allclose(_np(b.cosmo()["pkh"]), 1.21 * _np(ref.cosmo()["pkh"]), rtol=1e-13)
