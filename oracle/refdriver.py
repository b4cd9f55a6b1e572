"""Drive a Cobaya-style model (the UNMODIFIED reference, or this repository's own Cobaya classes) over the DR16 NGC
LRG x ELG x cross configuration (BASELINE config 3 / 4; cobaya/yamls/DR16_noric_LEX_..._kmax0.20.yaml in the reference)
through `oracle/refshim/cobaya` (mini-Cobaya).

TEST INFRASTRUCTURE: used by the golden generators, the tests and bench.py's CPU arm only.  The data are the reference's
own DR16 files, read from the compact fixture eftpipe_b200/data/dr16_ngc.npz and written back to text so that the
reference's readers (reader.py, window.py) load them the way they load the originals.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, ".."))
SHIM = os.path.join(HERE, "refshim")
FIXTURE = os.path.join(ROOT, "eftpipe_b200", "data", "dr16_ngc.npz")
TRACERS = (("LRG_NGC", 0.696), ("ELG_NGC", 0.849), ("X_NGC", 0.763))


def use_minicobaya():
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    import cobaya  # noqa: F401

    return cobaya


def write_dr16(dirpath):
    """text files in the reference's formats from the fixture; returns {name: path}"""
    os.makedirs(dirpath, exist_ok=True)
    fx = np.load(FIXTURE)
    out = {}
    for name, header in (("NGC_LRG_P", "k P0 P2 P4"), ("NGC_ELG_Q", "k Q0 Q2"), ("NGC_X_P", "k P0 P2 P4"),
                         ("NGC_ELG_P", "k P0 P2 P4")):
        out[name] = os.path.join(dirpath, name + ".txt")
        if not os.path.exists(out[name]):
            np.savetxt(out[name], fx[name], header=header, fmt="%.17e")
    for name in ("cov_NGC_L024E02X024_PQP", "cov_NGC_L024_P"):
        out[name] = os.path.join(dirpath, name + ".txt")
        if not os.path.exists(out[name]):
            np.savetxt(out[name], fx[name], fmt="%.17e")
    for t in ("LRG", "ELG", "X"):
        out["win_" + t] = os.path.join(dirpath, f"win_NGC_{t}.txt")
        if not os.path.exists(out["win_" + t]):
            np.savetxt(out["win_" + t], fx["win_" + t], fmt="%.17e")
    return out


class TableExtractor:
    """A `BoltzmannExtractor` (boltzmann.py:22-101) serving precomputed linear spectra: the sampled parameter `point`
    selects row int(point) of the table, so that every implementation driven through Cobaya consumes byte-identical
    inputs.  `table`: dict(pkh=(n, 200) on kh = logspace(-5, 0, 200), f=, DA=, H= (n,) [, h, rdrag]) or a path to an npz
    holding `<key>pkh`, ...; this stands where CLASS / CAMB stand in a real run."""

    def __init__(self, table, key="", param="point"):
        if isinstance(table, (str, os.PathLike)):
            z = np.load(table)
            table = {n: z[key + n] for n in ("pkh", "f", "DA", "H", "h", "rdrag") if key + n in z.files}
        self.table, self.param = table, param
        self.kh = np.logspace(-5, 0, 200)

    def initialize(self, zeff, use_cb=False, zextra=(), **kwargs):
        self.zeff = zeff

    def initialize_with_provider(self, provider):
        self.provider = provider

    def get_requirements(self):
        return {self.param: None}

    def calculate(self, **params_values_dict):
        self.i = int(round(float(self.provider.get_param(self.param))))

    def Pkh(self, kh):
        if not np.array_equal(kh, self.kh):
            raise ValueError("TableExtractor serves kh = logspace(-5, 0, 200) only (theory.py:562)")
        return self.table["pkh"][self.i].copy()

    def f(self):
        return float(self.table["f"][self.i])

    def DA(self):
        return float(self.table["DA"][self.i])

    def H(self):
        return float(self.table["H"][self.i])

    def h(self):
        return float(self.table["h"][self.i]) if "h" in self.table else None

    def rdrag(self):
        return float(self.table["rdrag"][self.i]) if "rdrag" in self.table else None

    def fsigma8_z(self):
        return -1


WEST_MARG = ("b3", "cct", "cr1", "cr2", "ce0", "cequad")  # the production yaml excludes cemono


def marg_block(scales=None):
    """`marg:` of the production yaml (:91-111): prefix form for the two auto tracers, flat names for the cross.
    scales=None: infinite (Jeffreys yaml); else the `_gauss.yaml` scales per name."""
    s = (lambda n: {"scale": None}) if scales is None else (lambda n: {"scale": scales[n]})
    west = {n: s(n) for n in WEST_MARG}
    return {"LRG_NGC_": west, "ELG_NGC_": dict(west), "X_NGC_ce0": s("ce0"), "X_NGC_cequad": s("cequad")}


GAUSS_SCALES = dict(b3=4, cct=2, cr1=4, cr2=4, ce0=2, cequad=2)


def config3_info(paths, tables, package="eftpipe", cache_dir=None, likelihoods=("jeffreys",), window_extra=None,
                 tracer_extra=None):
    """Cobaya info of the DR16 NGC LRG x ELG x X likelihood.  `tables[tracer]`: the TableExtractor table of that tracer
    (its own redshift).  package: "eftpipe" (reference) or "eftpipe_b200" (this repository's Cobaya classes)."""
    win = lambda t: dict({"window_configspace_file": paths["win_" + t]},
                         **({"window_fourier_file": os.path.join(cache_dir, f"{package}_win_NGC_{t}_acc4.npy")} if cache_dir else {}),
                         **(window_extra or {}))
    prov = lambda t: dict(provider="refdriver.TableExtractor", provider_kwargs=dict(table=tables[t]))
    tracers = {
        "LRG_NGC": dict(prefix="LRG_NGC_", z=0.696, nd=4.5e-5, window=win("LRG"), **prov("LRG_NGC")),
        "ELG_NGC": dict(prefix="ELG_NGC_", z=0.849, nd=2.3e-4, window=win("ELG"), **prov("ELG_NGC")),
        "X_NGC": dict(prefix="X_NGC_", z=0.763, cross=["LRG_NGC", "ELG_NGC"], window=win("X"), **prov("X_NGC")),
        "default": dict(km=0.7, kr=0.25, use_cb=True, with_IRresum=True, with_APeffect=True, with_window=True,
                        APeffect=dict(Om_AP=0.307115, rdrag_AP=147.66, h_AP=0.6777, APst=True),
                        window=dict(accboost=4, windowk=0.1), **(tracer_extra or {})),
    }
    like = dict(
        tracers=["LRG_NGC", "ELG_NGC", "X_NGC"], chained=[False, True, False],
        data={"LRG_NGC": dict(path=paths["NGC_LRG_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20),
              "ELG_NGC": dict(path=paths["NGC_ELG_Q"], ls=[0, 2], kmin=0.03, kmax=0.20),
              "X_NGC": dict(path=paths["NGC_X_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20)},
        cov=dict(path=paths["cov_NGC_L024E02X024_PQP"], Nreal=1000), with_binning=True)
    like["class"] = package + ".eftlike"
    likes = {}
    if "jeffreys" in likelihoods:
        likes["LEX_NGC"] = dict(like, jeffreys=True, marg=marg_block())
    if "gauss" in likelihoods:
        likes["LEX_NGC_gauss"] = dict(like, jeffreys=False, marg=marg_block(GAUSS_SCALES))
    uniform = lambda lo, hi: {"prior": {"min": lo, "max": hi}}
    params = {"point": uniform(0, 1e9)}
    for pre in ("LRG_NGC_", "ELG_NGC_"):
        params[pre + "b1"] = uniform(0, 4)
        params[pre + "c2"] = dict(uniform(-4, 4), drop=True)
        params[pre + "b2"] = {"value": f"lambda {pre}c2: {pre}c2 / np.sqrt(2.)"}  # yaml :213-220 with c4 = 0
        params[pre + "b4"] = {"value": f"lambda {pre}c2: {pre}c2 / np.sqrt(2.)"}
    for name in likes:  # derived: chi2, fullchi2, best-fit marginalised parameters
        params[name + "_chi2"], params[name + "_fullchi2"] = None, None
    for pre in ("LRG_NGC_", "ELG_NGC_"):
        for n in WEST_MARG:
            params["marg_" + pre + n] = None
    params["marg_X_NGC_ce0"], params["marg_X_NGC_cequad"] = None, None
    theory = {package + ".eftlss": dict(tracers=tracers, **({"cache_dir_path": cache_dir} if cache_dir else {}))}
    return dict(theory=theory, likelihood=likes, params=params)


def synthetic_tables(B, seed=20261018 + 3, unique=None):
    """per-tracer TableExtractor tables of B synthetic cosmologies (the same cosmology index across tracers, each tracer
    at its own redshift) - SURVEY.md section 8d"""
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    from eftpipe_b200 import synthetic

    out = {}
    for name, z in TRACERS:
        b = synthetic.make_batch(B, z, seed=seed, unique=unique)
        out[name] = dict(pkh=b.plin, f=b.f, DA=b.DA, H=b.H, h=b.h, rdrag=b.rdrag)
    return out


def draw_points(B, seed=11):
    """sampled nuisance values of B points: b1, c2 per auto tracer (yaml :201-250)"""
    rng = np.random.default_rng(seed)
    pts = {"point": np.arange(B, dtype=float)}
    for pre, b1 in (("LRG_NGC_", 2.1), ("ELG_NGC_", 1.4)):
        pts[pre + "b1"] = b1 + 0.05 * rng.standard_normal(B)
        pts[pre + "c2"] = 0.7 + 0.1 * rng.standard_normal(B)
    return pts


def reference_model(info):
    """mini-Cobaya model over the unmodified reference (imported by refload from /root/reference or baseline/_ref)"""
    use_minicobaya()
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import refload

    refload.load()
    from cobaya.model import get_model

    return get_model(info)
