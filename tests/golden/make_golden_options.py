"""Golden vectors for the non-default options of the loop + resummation rows:
`optiresum=True` (pybird.py:553-554, :1235-1244, :1382-1400), `IRcutoff` in {"all", "loop", "resum"} with `kIR`
(pybird.py:1127-1160, :1320-1334) and a non-default `LambdaIR`.  Runs the UNMODIFIED reference (mounted at
/root/reference) and the oracle on the same synthetic inputs, checks them against each other and writes the
reference's outputs to tests/golden/options_resum.npz.  Build-container only.
Usage:  python tests/golden/make_golden_options.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import ROOT, compare, meta, orc, orc_chain, ref_chain, refload, synthetic  # noqa: E402,F401

VARIANTS = {
    "optiresum": dict(common=dict(optiresum=True), resum=dict()),
    "optiresum_lambda1": dict(common=dict(optiresum=True), resum=dict(LambdaIR=1.0)),
    "ircut_all": dict(common=dict(IRcutoff="all", kIR=3e-3), resum=dict()),
    "ircut_loop": dict(common=dict(IRcutoff="loop", kIR=3e-3), resum=dict()),
    "ircut_resum": dict(common=dict(IRcutoff="resum", kIR=3e-3), resum=dict()),
    "nl2_optiresum_ircut": dict(common=dict(Nl=2, No=2, optiresum=True, IRcutoff=True, kIR=1e-3), resum=dict()),
}
KEEP = ("coef", "P22", "P13", "C11", "Cct", "pre_Cloopl", "X", "Y", "res_P11l", "res_Pctl", "res_Ploopl")


def main():
    ref = refload.load()
    pb = ref.pybird
    z = 0.7
    batch = synthetic.make_batch(2, z, seed=20261018 + 7)
    gold = dict(meta=meta(), kin=batch.kin, plin=batch.plin, f=batch.f, DA=batch.DA, H=batch.H, z=z,
                variants=json.dumps(VARIANTS))
    worst = 0.0
    for name, v in VARIANTS.items():
        kw = dict(Nl=3, No=3, kmax=0.3, kmA=0.7, krA=0.25, ndA=4.5e-5)
        kw.update(v["common"])
        co_r, co_o = pb.Common(**kw), orc.Common(**kw)
        nl_r, nl_o = pb.NonLinear(load=False, save=False, co=co_r), orc.NonLinear(co_o)
        rs_r, rs_o = pb.Resum(co=co_r, **v["resum"]), orc.Resum(co_o, **v["resum"])
        for i in range(len(batch)):
            args = (batch.kin, batch.plin[i], batch.f[i], batch.DA[i], batch.H[i], z)
            r = ref_chain(ref, co_r, nl_r, rs_r, *args)
            if co_r.IRcutoff:  # ref_chain's "coef" is the uncut set; record the set the k-space loops used
                r["coef"] = nl_r.Coef(r["_bird"], window=0.2, IRcut=co_r.IRcutoff in ("all", "loop"))
            o = orc_chain(co_o, nl_o, rs_o, *args)
            worst = max(worst, compare(f"{name} cosmology {i}", r, o))
            for key in KEEP:
                gold.setdefault(f"{name}__{key}", []).append(r[key])
    for key in list(gold):
        if "__" in key:
            gold[key] = np.array(gold[key])
    np.savez_compressed(os.path.join(HERE, "options_resum.npz"), **gold)
    print("worst oracle-vs-reference error: %.3e" % worst)
    assert worst < 1e-9, "oracle does not reproduce the reference"


if __name__ == "__main__":
    main()
