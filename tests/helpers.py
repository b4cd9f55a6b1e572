"""Shared builders for the test-suite."""
import json

import numpy as np

from eftpipe_b200 import plan as P
from eftpipe_b200 import synthetic


def config2_plan(g, with_projection=True, chained=False):
    """The plan of BASELINE config 2 from the golden fixture `g` (tests/golden/config2_chain.npz)."""
    apk = json.loads(str(g["ap"]))
    ap = dict(DA=synthetic.angular_distance(apk["Om_AP"], apk["z_AP"]), H=synthetic.hubble(apk["Om_AP"], apk["z_AP"]),
              APst=apk["APst"])
    grid = P.GridConfig(Nl=3)
    proj = None
    if with_projection:
        Weff = g["Weff_LRG"]
        binm, keff, _, _ = P.binning_matrix(grid.k, g["kout"])
        proj = P.compose_projection(grid, window=Weff,
                                    icc=dict(matrix=0.05 * Weff, PSN_times_Pshot=g["PSN"] * float(g["Pshot"])),
                                    binning=binm, chained=chained)
        proj["kout"] = keff
    return P.build_tracer_plan(Nl=3, ap=ap, projection=proj)


def split_terms(T):
    """(.., l, i, k) -> dict of the reference's term arrays."""
    return dict(P11l=T[..., 0:3, :], Pctl=T[..., 3:9, :], Ploopl=T[..., 9:21, :], Pstl=T[..., 21:24, :])


def option_variants(g):
    """{name: (Common kwargs, Resum kwargs)} of the options fixture `g` (tests/golden/options_resum.npz)."""
    return {k: (v["common"], v["resum"]) for k, v in json.loads(str(g["variants"])).items()}


def option_plan(common, resum):
    """host plan of one option variant (Nl=3 unless the variant says otherwise)"""
    return P.build_tracer_plan(Nl=common.get("Nl", 3), optiresum=common.get("optiresum", False),
                               ircutoff=common.get("IRcutoff", False), kIR=common.get("kIR"),
                               lambda_ir=resum.get("LambdaIR", 0.2))


class MatrixWindow:
    """A reference-style custom window plugin (theory.py:62-72): numpy code with an in-place `.Window(bird)`, applying a
    given (Na, Nk, Nl, Nk) operator to every term array except - optionally - the stochastic ones, and a constant to Picc."""

    def __init__(self, matrix, picc=None, window_st=True, co=None, icc=None, name=None):
        self.matrix, self.picc, self.window_st, self.co, self.icc, self.name = np.asarray(matrix, float), picc, window_st, co, icc, name

    def Window(self, bird):
        f = lambda T: np.einsum("akln,lin->aik", self.matrix, T)
        bird.P11l, bird.Pctl, bird.Ploopl = f(bird.P11l), f(bird.Pctl), f(bird.Ploopl)
        if bird.co.with_NNLO:
            bird.PctNNLOl = f(bird.PctNNLOl)
        if self.window_st:
            bird.Pstl = f(bird.Pstl)
        if self.picc is not None:
            bird.Picc = bird.Picc + self.picc


class SquaringWindow(MatrixWindow):
    """not linear: must be refused at plan build"""

    def Window(self, bird):
        bird.Ploopl = bird.Ploopl**2
