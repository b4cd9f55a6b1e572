"""Chained multipoles Q_l = P_l - A_l P_{l+2} - mirror of `eftpipe.chained` (chained.py:13-68)."""
from __future__ import annotations

import numpy as np
from scipy.special import eval_legendre

from . import plan as P
from .pybird import apply_node_operator
from .transformer import PlainBird, f_batch_minor


def chain_coeff(l: int) -> float:
    return ((2 * l + 1) * eval_legendre(l, 0.0)) / ((2 * l + 5) * eval_legendre(l + 2, 0.0))


class Chained:
    def __init__(self):
        self._ops = {}

    def chained_matrix(self, Nl: int):
        if Nl not in (2, 3, 4):
            raise NotImplementedError
        return P.chained_matrix(Nl)

    def transform(self, birdlike):
        Nl, nk = birdlike._T.shape[0], birdlike._T.shape[1]
        mat = self.chained_matrix(Nl)
        key = (Nl, nk)
        if key not in self._ops:
            holder = type("_Holder", (), {})()
            self._ops[key] = (np.kron(mat, np.eye(nk)), holder)
        full, holder = self._ops[key]
        T = apply_node_operator(birdlike, full, Nl - 1, stochastic=True, cache_owner=holder, in_place=False)
        picc = None if birdlike._picc is None else mat @ birdlike._picc
        return PlainBird(birdlike.f, birdlike.co, T, picc, birdlike.B, birdlike._squeeze, f_batch_minor(birdlike))
