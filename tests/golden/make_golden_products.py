"""Golden vectors modelled on the reference's own regression test of the theory component
(`tests/regression/test_eftlss.py::test_ELG_NGC_reg`): one tracer, every product the provider offers, asked for in one
evaluation - interpolators (plain and chained), grid products for the four (chained, binned) combinations, the Gaussian
(derivative) tables for the same, and the derived parameters - computed by the UNMODIFIED reference through
oracle/refshim/cobaya.  The Boltzmann code of the reference's yaml is replaced by the table extractor
(`refdriver.TableExtractor`) so that both implementations consume identical linear spectra.

Build container only:  python tests/golden/make_golden_products.py  ->  tests/golden/products_elg.npz
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import refdriver  # noqa: E402

B = 4
KOUT = np.arange(25) * 0.01 + 0.005          # the reference test's binning
TRACER = "ELG_NGC"
COMBOS = [(False, False), (True, False), (False, True), (True, True)]  # (chained, binned)
EFT = dict(b1=1.589354637, c2=1.266786101, b3=1.675518719e-01, cct=-7.109794839e-01, cr1=-2.635512354, cr2=-3.667395306e-01,
           ce0=1.634998078e-02, cequad=-4.125790673e-02)  # the reference test's sampled point (first point; the others vary)


def requirements():
    opts = {"ls": [0, 2, 4], "chained": [False, True], "binned": [False, True], "binning": {"kout": KOUT}}
    return {"nonlinear_Plk_interpolator": {TRACER: {"ls": [0, 2, 4], "chained": [False, True]}},
            "nonlinear_Plk_grid": {TRACER: dict(opts)}, "nonlinear_Plk_gaussian_grid": {TRACER: dict(opts)},
            TRACER + "_fsigma8_z": None, TRACER + "_alperp": None, TRACER + "_alpara": None}


def build_info(package, paths, table, cache_dir=None):
    """one auto tracer with the production settings of the ELG sample + a consumer that asks for every product"""
    win = dict(window_configspace_file=paths["win_ELG"], accboost=4, windowk=0.1)
    if cache_dir:
        win["window_fourier_file"] = os.path.join(cache_dir, f"{package}_win_NGC_ELG_acc4.npy")
    tracers = {TRACER: dict(prefix=TRACER + "_", z=0.849, nd=2.3e-4, km=0.7, kr=0.25, use_cb=True, with_IRresum=True, with_APeffect=True,
                            with_window=True, APeffect=dict(Om_AP=0.307115, rdrag_AP=147.66, h_AP=0.6777, APst=True), window=win,
                            provider="refdriver.TableExtractor", provider_kwargs=dict(table=table))}
    uniform = lambda lo, hi: {"prior": {"min": lo, "max": hi}}
    pre = TRACER + "_"
    params = {"point": uniform(0, 1e9), pre + "b1": uniform(0, 4), pre + "c2": dict(uniform(-4, 4), drop=True),
              pre + "b2": {"value": f"lambda {pre}c2: {pre}c2 / np.sqrt(2.)"}, pre + "b4": {"value": f"lambda {pre}c2: {pre}c2 / np.sqrt(2.)"}}
    for n in ("b3", "cct", "cr1", "cr2", "ce0", "cequad"):
        params[pre + n] = uniform(-100, 100)
    params[pre + "cemono"] = 0.0
    for d in ("fsigma8_z", "alperp", "alpara"):
        params[pre + d] = None
    theory = {package + ".eftlss": dict(tracers=tracers, **({"cache_dir_path": cache_dir} if cache_dir else {}))}
    return dict(theory=theory, likelihood={}, params=params)


def points():
    rng = np.random.default_rng(7)
    pts = {"point": np.arange(B, dtype=float)}
    for n, v in EFT.items():
        pts[f"{TRACER}_{n}"] = v * (1.0 + 0.1 * rng.standard_normal(B))
        pts[f"{TRACER}_{n}"][0] = v
    return pts


def collect(eft, provider=None):
    """everything the reference's regression test reads, from the component `eft` (the provider interface)"""
    out = {}
    out["plk"] = np.asarray(eft.get_nonlinear_Plk_interpolator(TRACER)([0, 2, 4], KOUT))
    out["plk_chained"] = np.asarray(eft.get_nonlinear_Plk_interpolator(TRACER, chained=True)([0, 2], KOUT))
    for ch, bn in COMBOS:
        tag = f"c{int(ch)}b{int(bn)}"
        ls, k, P = eft.get_nonlinear_Plk_grid(TRACER, chained=ch, binned=bn)
        out[f"grid_{tag}.ls"], out[f"grid_{tag}.k"], out[f"grid_{tag}.P"] = np.array(ls), np.asarray(k), np.asarray(P)
        ls, k, table = eft.get_nonlinear_Plk_gaussian_grid(TRACER, chained=ch, binned=bn)
        for name, v in table.items():
            out[f"gauss_{tag}.{name}"] = np.asarray(v)
    return out


def main():
    refdriver.use_minicobaya()
    from cobaya.theory import Theory

    paths = refdriver.write_dr16("/tmp/dr16txt")
    tabs = refdriver.synthetic_tables(B)[TRACER]
    cache = "/tmp/eftpipe_b200_bench_cache/ref"
    os.makedirs(cache, exist_ok=True)

    class Consumer(Theory):
        def get_requirements(self):
            return requirements()

    info = build_info("eftpipe", paths, tabs, cache_dir=cache)
    info["theory"]["consumer"] = {"class": Consumer}
    model = refdriver.reference_model(info)
    eft = model.theory["eftpipe.eftlss"]
    pts = points()
    out = dict(meta=json.dumps(dict(numpy=np.__version__, scipy=scipy.__version__, generated_by="tests/golden/make_golden_products.py")),
               kout=KOUT)
    for k, v in tabs.items():
        out["tab." + k] = v
    for k, v in pts.items():
        out["pt." + k] = v
    acc = {}
    for i in range(B):
        _, derived = model.loglikes({k: v[i] for k, v in pts.items()})
        got = collect(eft)
        for d in ("fsigma8_z", "alperp", "alpara"):
            got["derived." + d] = np.float64(derived[f"{TRACER}_{d}"])
        for k, v in got.items():
            acc.setdefault(k, []).append(v)
    for k, v in acc.items():
        out[k] = np.array(v[0]) if k.endswith((".ls", ".k")) else np.stack(v)
    path = os.path.join(HERE, "products_elg.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.shape(v) for k, v in out.items() if not k.startswith(("tab.", "pt."))})


if __name__ == "__main__":
    main()
