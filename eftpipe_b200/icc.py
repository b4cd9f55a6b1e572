"""Integral-constraint correction, APPLY STEP - mirror of `eftpipe.icc.IntegralConstraint`
(icc.py:119-497).  The Fourier-space matrices (`PSN[Na,Nk]`, `Wal[Na,Nl,Nk,Np]`) are loaded from the
reference's own cache format (`icc_fourier_file` .npz + .json meta, icc.py:298-357) or passed in directly;
building them from configuration-space files (2-D FFTLog, icc.py:359-446) is out of scope (SURVEY.md 2 #9/#10:
precompute only, and broken in the reference under SciPy >= 1.14)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from .plan import window_effective_matrix, window_pgrid
from .window import MetaInfoError


class IntegralConstraint:
    def __init__(self, Pshot, icc_fourier_file=None, icc_configspace_SN_file=None, icc_configspace_IC_file=None,
                 inorder=False, co=None, load=True, save=True, check_meta=True, Na=None, Nl=None, pmax=0.3, accboost=1,
                 withmask=True, windowk=0.05, name="eftpipe.icc", snapshot=False, PSN=None, Wal=None, **fft_options):
        from .pybird import common

        self.co = co if co is not None else common
        if icc_fourier_file is None and PSN is None:
            if icc_configspace_SN_file or icc_configspace_IC_file:
                raise NotImplementedError("computing the ICC matrices from configuration space is out of scope; "
                                          "provide icc_fourier_file (reference cache format) or PSN/Wal arrays")
            raise ValueError("No ICC file specified")
        self.withmask, self.windowk = withmask, windowk
        Na = Na or self.co.Nl
        Nl = Nl or self.co.Nl
        if Na > self.co.Nl or Nl > self.co.Nl:
            raise ValueError(f"request Na={Na}, Nl={Nl} while bird only compute Nl up to {self.co.Nl}")
        if Na > Nl:
            raise ValueError(f"dangerous settings Na={Na} > Nl={Nl}")
        self.p = window_pgrid(kmax=pmax, accboost=accboost)
        self.Pshot = Pshot
        if PSN is None:
            path = Path(icc_fourier_file).resolve()
            data = np.load(path)  # OSError propagates like the reference's failed load + no config-space files
            PSN, Wal = data["PSN"], data["Wal"]
            meta_file = path.with_suffix(".json")
            if check_meta and meta_file.exists():
                with meta_file.open("r") as fh:
                    meta = json.load(fh)
                for key, val in dict(Na=Na, Nl=Nl, pmax=pmax, accboost=accboost, k=self.co.k.tolist()).items():
                    if key in meta and meta[key] != val:
                        raise MetaInfoError(f"inconsistent meta info for {key}: {meta[key]} != {val}")
        self.Wal = np.asarray(Wal, float)
        self.PSN = np.asarray(PSN, float) * Pshot  # icc.py:296 "always need Pshot"
        self.snapshot = snapshot

    def effective_matrix(self):
        """(Na, Nk, Nl, Nk) operator of `integrWindow` (icc.py:448-484)."""
        return window_effective_matrix(self.Wal, self.p, self.co.k, windowk=self.windowk, withmask=self.withmask)
