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


class EisensteinHu(BoltzmannExtractor):
    """A batched, on-device producer of the pipeline's inputs: the Eisenstein & Hu (1998) linear power of
    flat LCDM at the tracer's redshift, sigma8-normalised, with the matching f, DA, H - `eftb_eh_power`, one CTA per point.
    It stands where CLASS / CAMB stand in the reference (neither exists in this image) and is the model every synthetic
    input here comes from (synthetic.linear_power), so it is checked against that host code.

    Sampled parameters (names configurable through `params`): Omega_m, h, sigma8, scalars or (B,) arrays, read from the
    provider like the reference's extractors read theirs (boltzmann.py:289-305).  Everything it returns is a CUDA tensor:
    P_lin never crosses PCIe; per point the host hands over three numbers."""

    def __init__(self, params=("omegam", "h", "sigma8"), omega_b=0.02214, ns=0.9611, Tcmb=2.7255, rdrag=None, prefix="", ngl=96,
                 nsig=2000, share_sigma8_with=None):
        """share_sigma8_with: another EisensteinHu of the same evaluation (another tracer, another redshift).  The sigma8
        normalisation integral does not depend on the redshift and is ten times the work of the 200 output nodes; this
        extractor reuses the other one's when that one has just been evaluated on the very same parameter tensors."""
        self.names = [prefix + p for p in params]
        self.omega_b, self.ns, self.Tcmb, self._rdrag, self.ngl, self.nsig = omega_b, ns, Tcmb, rdrag, ngl, nsig
        self.provider, self._out = None, None
        self._primary, self._gen, self._seen, self._key, self._sig2, self._event = share_sigma8_with, 0, -1, None, None, None

    def get_requirements(self):
        return {n: None for n in self.names}

    def _param(self, name):
        p = self.provider
        return p.get_param(name) if hasattr(p, "get_param") else p[name]

    def calculate(self, **params_values_dict):
        import ctypes as C

        from . import _lib

        torch = _lib.require_cuda()
        lib = _lib.load()
        vals = [params_values_dict[n] if n in params_values_dict else self._param(n) for n in self.names]
        cols = [v.to("cuda", torch.float64).reshape(-1) if isinstance(v, torch.Tensor) else
                torch.as_tensor(np.atleast_1d(np.asarray(v, float)), device="cuda") for v in vals]
        B = max(c.numel() for c in cols)
        theta = torch.stack([c.expand(B) for c in cols], dim=1).contiguous()
        if getattr(self, "_consts", None) is None:
            u, w = np.polynomial.legendre.leggauss(self.ngl)
            self._consts = tuple(torch.as_tensor(a, device="cuda") for a in (KH, 0.5 * (u + 1.0), 0.5 * w))
        kh, gu, gw = self._consts
        pkh = torch.empty((B, kh.numel()), dtype=torch.float64, device="cuda")
        f, DA, H = (torch.empty(B, dtype=torch.float64, device="cuda") for _ in range(3))
        p = lambda t: C.c_void_p(t.data_ptr())
        # the same parameter tensors (identity, not value: checked per call) and a primary evaluated since we last looked
        key = tuple((v.data_ptr(), tuple(v.shape)) if isinstance(v, torch.Tensor) else id(v) for v in vals) + (self.omega_b, self.ns, self.Tcmb, self.nsig)
        pr = self._primary
        reuse = pr is not None and pr._sig2 is not None and pr._key == key and pr._gen != self._seen and all(isinstance(v, torch.Tensor) for v in vals)
        if reuse:
            sig2, self._seen = pr._sig2, pr._gen
            # ordering across streams: the primary's kernel may run on another tracer's stream
            if pr._event is not None:
                torch.cuda.current_stream().wait_event(pr._event)
        else:
            sig2 = torch.empty(B, dtype=torch.float64, device="cuda")
        _lib.check(lib.eftb_eh_power(B, p(theta), float(self.zeff), float(self.omega_b), float(self.ns), float(self.Tcmb), p(kh),
                                     kh.numel(), p(gu), p(gw), self.ngl, self.nsig, p(pkh), p(f), p(DA), p(H), p(sig2), int(reuse),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "eftb_eh_power")
        self._sig2, self._key, self._gen = sig2, key, self._gen + 1
        self._event = torch.cuda.Event()
        self._event.record(torch.cuda.current_stream())
        self._out = dict(pkh=pkh, f=f, DA=DA, H=H, h=theta[:, 1].contiguous())
        if self._rdrag is not None:
            self._out["rdrag"] = torch.full((B,), float(self._rdrag), dtype=torch.float64, device="cuda")

    def Pkh(self, kh):
        if not np.array_equal(np.asarray(kh, float), KH):
            raise ValueError("EisensteinHu serves kh = logspace(-5, 0, 200) (theory.py:562)")
        return self._out["pkh"]

    f = lambda self: self._out["f"]
    DA = lambda self: self._out["DA"]
    H = lambda self: self._out["H"]
    h = lambda self: self._out["h"]
    rdrag = lambda self: self._out.get("rdrag")

    def cosmo(self, kh=KH):
        return dict(self._out)


def find_boltzmann_extractor(name, kwargs=None):
    """boltzmann.py:351-363; the Cobaya-backed CLASS / CAMB / Matryoshka extractors are not available here"""
    if not isinstance(name, str):
        return name
    if name in ("camb", "classy", "classynu", "matryoshka"):
        raise NotImplementedError(f"the {name} provider needs Cobaya and a Boltzmann code, which this build does not ship; "
                                  "hand the linear power over with boltzmann.ArrayExtractor or LinearPowerFile")
    module_name, class_name = name.rsplit(".", 1)
    return getattr(importlib.import_module(module_name), class_name)(**(kwargs or {}))
