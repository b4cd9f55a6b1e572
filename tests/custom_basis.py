"""A reference-style EFT parameter basis by dotted path (parambasis.py:139-162, :457-465): plain numpy code with the
`EFTBasis` protocol and none of this package's kernel hooks.  "b1 / sigma-style": P = b1^2 P11l[0] + b1 f P11l[1] + f^2
P11l[2] + b1 * s * Ploopl[1] + Pctl[0] * c0 / km^2 + Pstl[0] * e0 / nd, Gaussian parameters c0, e0."""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class _Component:
    Plin: np.ndarray
    Ploop: np.ndarray
    Pct: np.ndarray
    Pst: np.ndarray
    Picc: np.ndarray

    def sum(self):
        return self.Plin + self.Ploop + self.Pct + self.Pst + self.Picc


@dataclass(frozen=True)
class ToyBasis:
    prefix: str = ""
    cross_prefix: list = field(default_factory=list)

    @classmethod
    def get_name(cls):
        return "toy"

    @classmethod
    def counterform(cls):
        return "westcoast"

    def non_gaussian_params(self):
        return [self.prefix + "b1", self.prefix + "s"]

    def gaussian_params(self):
        return [self.prefix + "c0", self.prefix + "e0"]

    def default(self):
        return {p: 0.0 for p in self.gaussian_params()}

    def reduce_Plk(self, bird, params_values_dict):
        p = dict(self.default(), **params_values_dict)
        b1, s, c0, e0 = (p[self.prefix + n] for n in ("b1", "s", "c0", "e0"))
        No, f, co = bird.co.No, bird.f, bird.co
        b11 = np.array([b1 * b1, b1 * f, f * f])
        Plin = np.einsum("b,lbx->lx", b11, bird.P11l[:No])
        Ploop = b1 * s * bird.Ploopl[:No, 1] + np.sin(s) * bird.Ploopl[:No, 5]
        Pct = c0 / co.kmA**2 * bird.Pctl[:No, 0]
        Pst = e0 / co.ndA * bird.Pstl[:No, 0]
        return _Component(Plin, Ploop, Pct, Pst, bird.Picc[:No])

    def reduce_Plk_gaussian_table(self, bird, params_values_dict, requires=None):
        No, co = bird.co.No, bird.co
        out = {self.prefix + "c0": bird.Pctl[:No, 0] / co.kmA**2, self.prefix + "e0": bird.Pstl[:No, 0] / co.ndA}
        return {k: v for k, v in out.items() if requires is None or k in requires}
