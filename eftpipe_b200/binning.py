"""k-bin averaging - mirror of `eftpipe.binning.Binning` (binning.py:17-162).

Every bin average  int k^2 spline(P)(k) dk / int k^2 dk  (100*accboost trapezoid nodes per bin) is linear in
the node values, so the stage is one fixed (nbin x Nk) matrix per multipole, applied by the DMMA GEMM."""
from __future__ import annotations

import numpy as np

from . import plan as P
from .pybird import apply_node_operator, common
from .transformer import PlainBird, f_batch_minor


class Binning:
    def __init__(self, kout, accboost=1, decimals=2, co=common, name="pybird.binning", kstart=None, kend=None,
                 nbins=None):
        self.kout = np.array(kout)
        self.co, self.accboost, self.decimals = co, accboost, decimals
        kedges = None
        if not (kstart is None and kend is None and nbins is None):
            if kstart is None or kend is None or nbins is None:
                raise ValueError("need specify kstart, kend and nbins together")
            kedges = np.linspace(kstart, kend, nbins + 1)  # binning.py:89-96
            il = np.searchsorted(kedges, self.kout[0]) - 1
            ir = np.searchsorted(kedges, self.kout[-1], side="right") + 1
            kedges = kedges[il:ir]
        self.matrix, self.keff, self.binmin, self.binmax = P.binning_matrix(
            co.k, self.kout, accboost=accboost, decimals=decimals, kedges=kedges)
        self.nbin = self.matrix.shape[0]
        self._full = None

    def node_matrix(self, Nl):
        """block-diagonal (Nl*nbin, Nl*Nk) operator."""
        return np.kron(np.eye(Nl), self.matrix)

    def integrBinning(self, P_nodes):
        """host helper for cosmology-independent arrays (e.g. Picc): (.., Nk) -> (.., nbin)."""
        return np.asarray(P_nodes) @ self.matrix.T

    def kbinning(self, bird):
        Nl = bird._T.shape[0]
        if self._full is None or self._full.shape[1] != Nl * self.co.Nk:
            self._full = self.node_matrix(Nl)
            self._node_operator = None
        T = apply_node_operator(bird, self._full, Nl, stochastic=True, cache_owner=self, in_place=False)
        picc = None if bird._picc is None else self.integrBinning(bird._picc)
        return PlainBird(bird.f, bird.co, T, picc, bird.B, bird._squeeze, f_batch_minor(bird))

    def transform(self, birdlike):
        return self.kbinning(birdlike)
