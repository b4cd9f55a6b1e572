"""Simplified interface to the batched EFTofLSS multipoles - mirror of `eftpipe.model.EFTModel` (model.py:15-459).

Same builder methods and keywords (`set_IRresum`, `set_window`, `set_APeffect`, `set_icc`, `done(ellmax)`, `clone`,
`Plk_mm`, `__call__(b1A, c2A, b3A, c4A, cctA, cr1A, cr2A, ce0, cemono, cequad, b1B, ...)` returning a
`PlkInterpolator`).  The reference builds a Cobaya model with CLASS as the Boltzmann provider (`set_cosmology`,
model.py:137-169); neither exists in this image and the linear power spectrum is an *input* of the B200 path, so the
cosmology enters through `set_linear_power(pkh, f, DA, H)` instead - arrays with a leading batch axis, one row per
cosmology, on the reference's grid kh = logspace(-5, 0, 200) (theory.py:562).  Every call evaluates the whole batch.
"""
from __future__ import annotations

import math
from copy import deepcopy

import numpy as np

from .theory import EFTLSS, PlkInterpolator


class EFTModel:
    def __init__(self, z, ndA=1e-4, ndB=None, kmA=0.7, krA=0.25, kmB=0.7, krB=0.25, cache_dir_path=None, use_cb=True,
                 with_RSD=True, IRcutoff=False, kIR=None, Nl=None):
        self._done = False
        self.z, self.use_cb = z, use_cb
        self.cross = ndB is not None
        if self.cross:  # model.py:78-83
            self.tracers = {"A": {"prefix": "A_", "z": z, "nd": ndA, "km": kmA, "kr": krA},
                            "B": {"prefix": "B_", "z": z, "nd": ndB, "km": kmB, "kr": krB},
                            "x": {"prefix": "x_", "z": z, "cross": ["A", "B"]}}
        else:
            self.tracers = {"x": {"prefix": "x_", "z": z, "nd": ndA, "km": kmA, "kr": krA}}
        self.tracers["default"] = {"with_IRresum": False, "IRcutoff": IRcutoff, "kIR": kIR, "with_RSD": with_RSD}
        if Nl is not None:
            self.tracers["default"]["Nl"] = Nl
        self.cache_dir_path = cache_dir_path
        self._cosmo = None

    # ---- the reference's builder methods (model.py:171-343): same keywords, forwarded to the tracer configuration
    def set_cosmology(self, *args, **kwargs):
        raise NotImplementedError("no Boltzmann code in this build: pass the linear power with set_linear_power(pkh, f, DA, H)")

    def set_linear_power(self, pkh, f, DA=None, H=None, rdrag=None, h=None):
        """pkh (B, 200) on kh = logspace(-5, 0, 200) in (Mpc/h)^3; f, DA, H (B,) - the quantities the reference pulls from
        its Boltzmann provider (theory.py:559-565)."""
        self._cosmo = dict(pkh=pkh, f=f, DA=DA, H=H, rdrag=rdrag, h=h)
        return self

    def set_IRresum(self, optiresum=False, NFFT=192):
        d = self.tracers["x"]
        d["with_IRresum"], d["IRresum"], d["optiresum"] = True, {"NFFT": NFFT}, optiresum
        return self

    def set_window(self, window_fourier_file=None, window_configspace_file=None, Na=None, Nl=None, Nq=3, pmax=0.3,
                   accboost=1, withmask=True, windowk=0.05, Nmax=4096, xmin_factor=1.0, xmax_factor=100.0, bias=-1.6,
                   window_param=1, window_st=True, window_configspace_array=None):
        d = self.tracers["x"]
        d["with_window"] = True
        d["window"] = dict(window_fourier_file=window_fourier_file, window_configspace_file=window_configspace_file, Na=Na,
                           Nl=Nl, Nq=Nq, pmax=pmax, accboost=accboost, withmask=withmask, windowk=windowk, Nmax=Nmax,
                           xmin_factor=xmin_factor, xmax_factor=xmax_factor, bias=bias, window_param=window_param,
                           window_st=window_st)
        if window_configspace_array is not None:
            d["window"]["window_configspace_array"] = window_configspace_array
        return self

    def set_APeffect(self, Om_AP, z_AP=None, rdrag_AP=None, h_AP=None, nbinsmu=200, accboost=1, Nlmax=None, APst=False):
        d = self.tracers["x"]
        d["with_APeffect"] = True
        d["APeffect"] = dict(Om_AP=Om_AP, z_AP=z_AP or d["z"], rdrag_AP=rdrag_AP, h_AP=h_AP, nbinsmu=nbinsmu,
                             accboost=accboost, Nlmax=Nlmax, APst=APst)
        return self

    def set_icc(self, Pshot, icc_fourier_file=None, **kwargs):
        d = self.tracers["x"]
        d["with_icc"] = True  # model.py:332
        d["icc"] = dict(Pshot=Pshot, icc_fourier_file=icc_fourier_file, **kwargs)
        return self

    def done(self, ellmax=2, debug=False, logging=False, zextra=()):
        if self._done:
            raise RuntimeError("already done")
        ls = list(range(0, ellmax + 1, 2))
        self.theory = EFTLSS(deepcopy(self.tracers), cache_dir_path=self.cache_dir_path)
        self.theory.must_provide({"nonlinear_Plk_grid": {"x": {"ls": ls, "binned": False}}}).initialize()
        self.ls = ls
        self._done = True
        return self

    def clone(self):
        ret = type(self)(1.0)
        ret.z, ret.use_cb, ret.cross = self.z, self.use_cb, self.cross
        ret.tracers, ret.cache_dir_path = deepcopy(self.tracers), self.cache_dir_path
        return ret

    def f(self):
        return None if self._cosmo is None else self._cosmo["f"]

    def Plk_mm(self, cct=0.0, cr1=0.0, cr2=0.0):
        return self(b1A=1, c2A=math.sqrt(2) / 2, b3A=1, c4A=math.sqrt(2) / 2, cctA=cct, cr1A=cr1, cr2A=cr2)

    def __call__(self, b1A, c2A, b3A, c4A, cctA, cr1A, cr2A, ce0=0.0, cemono=0.0, cequad=0.0, b1B=0.0, c2B=0.0,
                 b3B=0.0, c4B=0.0, cctB=0.0, cr1B=0.0, cr2B=0.0) -> PlkInterpolator:
        """model.py:429-459: parameters may be scalars or (B,) arrays; returns the batched `PlkInterpolator`."""
        if not self._done:
            raise RuntimeError("need to call done() first")
        if self._cosmo is None:
            raise RuntimeError("need the linear power: call set_linear_power(pkh, f, DA, H)")
        b24 = lambda c2, c4: ((np.asarray(c2) + np.asarray(c4)) / np.sqrt(2.0), (np.asarray(c2) - np.asarray(c4)) / np.sqrt(2.0))
        b2A, b4A = b24(c2A, c4A)  # model.py:100-103
        if self.cross:
            b2B, b4B = b24(c2B, c4B)
            params = dict(A_b1=b1A, A_b2=b2A, A_b3=b3A, A_b4=b4A, A_cct=cctA, A_cr1=cr1A, A_cr2=cr2A,
                          B_b1=b1B, B_b2=b2B, B_b3=b3B, B_b4=b4B, B_cct=cctB, B_cr1=cr1B, B_cr2=cr2B,
                          x_ce0=ce0, x_cemono=cemono, x_cequad=cequad)
        else:
            params = dict(x_b1=b1A, x_b2=b2A, x_b3=b3A, x_b4=b4A, x_cct=cctA, x_cr1=cr1A, x_cr2=cr2A,
                          x_ce0=ce0, x_cemono=cemono, x_cequad=cequad)
        cosmo = {k: v for k, v in self._cosmo.items() if v is not None}
        self.theory.calculate({"x": cosmo})
        return self.theory.get_nonlinear_Plk_interpolator("x", params)
