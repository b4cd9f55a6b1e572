"""CPU: the plan operators built by eftpipe_b200/plan.py, driven by the numpy emulation of the CUDA
kernels (tests/emulate.py), against the reference goldens.  This validates every precomputed operator
and the restructured algebra (anti-diagonal sums, resummation polynomial form, B-spline AP, composed
projection) without a GPU."""
import numpy as np
import pytest

import emulate as E
import helpers
from conftest import rowmax_rel
from eftpipe_b200 import plan as P

TOL = 1e-9  # parity bar is 1e-8; the emulation sits 2-3 orders below it


@pytest.fixture(scope="module")
def chain(golden2):
    pl = helpers.config2_plan(golden2)
    out = E.run_chain(pl, golden2["plin"], golden2["f"], golden2["DA"], golden2["H"])
    return pl, out


def test_front_products(chain, golden2):
    pl, r = chain
    g, F, B = golden2, chain[1]["F"], 3
    c = E.rows(pl, F, "cre") + 1j * E.rows(pl, F, "cim")
    assert np.abs(c.T - g["coef"][:, :129]).max() <= 1e-13 * np.abs(g["coef"]).max()
    assert rowmax_rel(E.rows(pl, F, "P11").T, g["P11"]) <= TOL
    assert rowmax_rel(E.rows(pl, F, "X").T, g["X"]) <= TOL
    assert rowmax_rel(E.rows(pl, F, "Y").T, g["Y"]) <= TOL
    assert rowmax_rel(E.rows(pl, F, "C11").reshape(3, 80, B).transpose(2, 0, 1), g["C11"]) <= TOL
    assert rowmax_rel(E.rows(pl, F, "Cct").reshape(3, 80, B).transpose(2, 0, 1), g["Cct"]) <= TOL


def test_antidiagonal_loops(chain, golden2):
    _, r = chain
    assert rowmax_rel(r["P22"].transpose(2, 0, 1), golden2["P22"]) <= TOL
    assert rowmax_rel(r["Cs"][:, :28].transpose(3, 0, 1, 2), golden2["C22"]) <= TOL
    assert rowmax_rel(r["Cs"][:, 28:].transpose(3, 0, 1, 2), golden2["C13"]) <= TOL


@pytest.mark.parametrize("stage,prefix", [("T_pre", "pre_"), ("T_res", "res_"), ("T_ap", "ap_")])
def test_terms(chain, golden2, stage, prefix):
    _, r = chain
    T = helpers.split_terms(r[stage].transpose(3, 0, 2, 1))
    for name, arr in T.items():
        key = prefix + name
        if key in golden2:
            assert rowmax_rel(arr, golden2[key]) <= TOL, key


def test_projection(chain, golden2):
    pl, r = chain
    out = r["out"]  # (nout, nterm, B)
    nl, nk = pl.out_shape
    T = helpers.split_terms(out.reshape(nl, nk, out.shape[1], -1).transpose(3, 0, 2, 1))
    for name, arr in T.items():
        assert rowmax_rel(arr, golden2["bin_" + name]) <= TOL, name
    assert rowmax_rel(pl.picc_out.reshape(nl, nk), golden2["bin_Picc"][0]) <= TOL


def test_chained_projection(golden2):
    pl = helpers.config2_plan(golden2, chained=True)
    assert pl.out_shape == (2, golden2["kout"].size)
    T = golden2["ap_Ploopl"][0]  # (l, i, k) after AP
    X = T.transpose(0, 2, 1).reshape(3 * 50, 12)
    got = (pl.project @ X).reshape(2, -1, 12).transpose(0, 2, 1)
    assert rowmax_rel(got, golden2["chn_Ploopl"][0]) <= TOL


def test_pair_table_counts():
    pl = P.build_tracer_plan(Nl=2, with_resum=False)
    n = pl.Nmax
    assert pl.pair_offsets[-1] == pl.pair_table.shape[0] == sum(t // 2 + 1 for t in range(n + 1))
    assert pl.pair_table.shape[1] == P.NCH


# ---- non-default options of the loop / resummation rows: optiresum, IRcutoff, LambdaIR ----
VARIANTS = ["optiresum", "optiresum_lambda1", "ircut_all", "ircut_loop", "ircut_resum", "nl2_optiresum_ircut"]


@pytest.mark.parametrize("name", VARIANTS)
def test_option_variants(golden_opts, name):
    g = golden_opts
    common, resum = helpers.option_variants(g)[name]
    pl = helpers.option_plan(common, resum)
    grid = pl.grid
    r = E.run_chain(pl, g["plin"], g["f"])
    B, Nl = g["plin"].shape[0], grid.Nl
    ref = lambda key: g[f"{name}__{key}"]
    on_sr = lambda C: C if grid.E is None else np.einsum("rs,...s->...r", grid.E, C)  # the device holds E.C only
    F = r["F"]
    c = E.rows(pl, F, "cre") + 1j * E.rows(pl, F, "cim")
    assert np.abs(c.T - ref("coef")[:, :129]).max() <= 1e-13 * np.abs(ref("coef")).max()
    assert ("cre_cf" in pl.front.rows) == (common.get("IRcutoff") in ("loop", "resum"))
    assert rowmax_rel(r["P22"].transpose(2, 0, 1), ref("P22")) <= TOL
    assert rowmax_rel(E.rows(pl, F, "X").T, ref("X")) <= TOL
    assert rowmax_rel(E.rows(pl, F, "Y").T, ref("Y")) <= TOL
    assert rowmax_rel(E.rows(pl, F, "C11").reshape(Nl, grid.Ns, B).transpose(2, 0, 1), on_sr(ref("C11"))) <= TOL
    assert rowmax_rel(E.rows(pl, F, "Cct").reshape(Nl, grid.Ns, B).transpose(2, 0, 1), on_sr(ref("Cct"))) <= TOL
    assert rowmax_rel(r["Cr"][:, 2:14].transpose(3, 0, 1, 2), on_sr(ref("pre_Cloopl"))) <= TOL
    T = helpers.split_terms(r["T_res"].transpose(3, 0, 2, 1))
    for key in ("P11l", "Pctl", "Ploopl"):
        assert rowmax_rel(T[key], ref("res_" + key)) <= TOL, key


def test_option_errors():
    with pytest.raises(ValueError):
        P.build_tracer_plan(Nl=2, ircutoff="all")  # kIR missing (pybird.py:528-529)
    with pytest.raises(ValueError):
        P.build_tracer_plan(Nl=2, ircutoff="sometimes", kIR=1e-3)  # pybird.py:1160
