"""Input side of the pipeline - mirror of `eftpipe.boltzmann` (boltzmann.py:20-363): the `BoltzmannExtractor` protocol
eftlss pulls its cosmology-dependent inputs through (`Pkh, f, DA, H, h, rdrag, fsigma8_z`), with a leading batch axis.

CLASS / CAMB / Matryoshka are not available in this image (they are input *producers*); what is
mirrored is what runs without them:
  * `LinearPowerFile` (boltzmann.py:246-309): a fixed linear power spectrum from a file (template fits) with the
    per-point parameters `{prefix}f, {prefix}alperp, {prefix}alpara` - same constructor, same low-k power-law extension,
    same log-log cubic interpolation;
  * `ArrayExtractor`: the batched hand-over for any external producer (an emulator, a Boltzmann code run elsewhere).
`theory.EFTLSS.calculate` accepts an extractor per tracer in place of the `dict(pkh=, f=, DA=, H=)`.
"""
from __future__ import annotations

import importlib

import numpy as np
from scipy.interpolate import interp1d

KH = np.logspace(-5, 0, 200)  # the grid eftlss samples the linear power on (theory.py:562)


class BoltzmannExtractor:
    """boltzmann.py:20-101 - same method names; every getter returns one value per point of the batch"""

    def initialize(self, zeff, use_cb=False, zextra=(), **kwargs):
        self.zeff, self.use_cb, self.zextra = zeff, use_cb, list(zextra)

    def initialize_with_provider(self, provider):
        self.provider = provider

    def get_requirements(self):
        return {}

    def calculate(self, **params_values_dict):
        pass

    def Pkh(self, kh):
        raise NotImplementedError

    def f(self):
        raise NotImplementedError

    def DA(self):
        raise NotImplementedError

    def H(self):
        raise NotImplementedError

    def h(self):
        return None

    def rdrag(self):
        return None

    def fsigma8_z(self):
        return -1  # boltzmann.py:100-101

    def cosmo(self, kh=KH):
        """the per-tracer input of `theory.EFTLSS.calculate` (what theory.py:559-565 collects)"""
        out = dict(pkh=self.Pkh(kh), f=self.f(), DA=self.DA(), H=self.H())
        for name in ("h", "rdrag", "fsigma8_z"):
            v = getattr(self, name)()
            if v is not None and not (np.isscalar(v) and v == -1):
                out[name] = v
        return out


class ArrayExtractor(BoltzmannExtractor):
    """Batched hand-over: pkh (B, len(kh)) sampled on `kh` (default: eftlss's grid), f / DA / H [/ h, rdrag, fsigma8_z] (B,)"""

    def __init__(self, pkh, f, DA, H, h=None, rdrag=None, fsigma8_z=None, kh=KH):
        self._kh, self._pkh = np.asarray(kh, float), pkh
        self._vals = dict(f=f, DA=DA, H=H, h=h, rdrag=rdrag, fsigma8_z=fsigma8_z)

    def Pkh(self, kh):
        kh = np.asarray(kh, float)
        if kh.shape == self._kh.shape and np.array_equal(kh, self._kh):
            return self._pkh
        p = np.asarray(self._pkh.detach().cpu().numpy() if hasattr(self._pkh, "detach") else self._pkh, float)
        return np.exp(interp1d(np.log(self._kh), np.log(p), kind="cubic", axis=-1)(np.log(kh)))

    f = lambda self: self._vals["f"]
    DA = lambda self: self._vals["DA"]
    H = lambda self: self._vals["H"]
    h = lambda self: self._vals["h"]
    rdrag = lambda self: self._vals["rdrag"]

    def fsigma8_z(self):
        return -1 if self._vals["fsigma8_z"] is None else self._vals["fsigma8_z"]


class LinearPowerFile(BoltzmannExtractor):
    """boltzmann.py:246-309.  `path`: a two-column text file (k, P) or a (k, P) pair of arrays.  The provider is any
    mapping / object with `get_param` holding `{prefix}f, {prefix}alperp, {prefix}alpara` as scalars or (B,) arrays.
    As in the reference, the first `DA()` / `H()` calls return 1 (they seed the AP fiducial), later ones return
    `alperp` and `1 / alpara`."""

    def __init__(self, path, gz=1, prefix=""):
        k, pk = (np.asarray(path[0], float), np.asarray(path[1], float)) if isinstance(path, (tuple, list)) else np.loadtxt(path, unpack=True)
        pk = pk * gz**2
        self.prefix = prefix
        if k[0] > 1e-5:  # boltzmann.py:260-267: power-law extension down to k = 1e-5
            ns = (np.log(pk[1]) - np.log(pk[0])) / (np.log(k[1]) - np.log(k[0]))
            lowk = np.geomspace(1e-5, k[0], 100, endpoint=False)
            k, pk = np.hstack((lowk, k)), np.hstack((pk[0] * (lowk / k[0]) ** ns, pk))
        fn = interp1d(np.log(k), np.log(pk), kind="cubic")
        self.plin = lambda kk: np.exp(fn(np.log(kk)))
        self._returned_DA = self._returned_H = False
        self.provider = None

    def initialize(self, zeff, use_cb=False, zextra=(), **kwargs):
        super().initialize(zeff, use_cb, zextra)

    def get_requirements(self):
        return {self.prefix + "f": None, self.prefix + "alperp": None, self.prefix + "alpara": None}

    def _param(self, name):
        p = self.provider
        v = p.get_param(self.prefix + name) if hasattr(p, "get_param") else p[self.prefix + name]
        return np.atleast_1d(np.asarray(v, float))

    def Pkh(self, kh):
        """(B, len(kh)): the same template for every point of the batch"""
        row = self.plin(np.asarray(kh, float))
        B = self._param("f").size if self.provider is not None else 1
        return np.broadcast_to(row, (B, row.size)).copy()

    def f(self):
        return self._param("f")

    def DA(self):
        if self._returned_DA:
            return self._param("alperp")
        self._returned_DA = True
        return 1

    def H(self):
        if self._returned_H:
            return 1 / self._param("alpara")
        self._returned_H = True
        return 1

    def h(self):
        return 1

    def rdrag(self):
        return 1


def find_boltzmann_extractor(name, kwargs=None):
    """boltzmann.py:351-363; the Cobaya-backed CLASS / CAMB / Matryoshka extractors are not available here"""
    if not isinstance(name, str):
        return name
    if name in ("camb", "classy", "classynu", "matryoshka"):
        raise NotImplementedError(f"the {name} provider needs Cobaya and a Boltzmann code, which this build does not ship; "
                                  "hand the linear power over with boltzmann.ArrayExtractor or LinearPowerFile")
    module_name, class_name = name.rsplit(".", 1)
    return getattr(importlib.import_module(module_name), class_name)(**(kwargs or {}))
