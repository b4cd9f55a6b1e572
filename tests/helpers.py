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
