"""Extract the closed-form constant tables of the PyBird one-loop calculation into a
compact numeric file, `eftpipe_b200/data/pybird_tables.npz`.

Run ONCE in the build container (needs /root/reference and sympy); the output is
committed.  The tables are mathematical constants of the EFTofLSS one-loop calculation
(Perko et al. 2016, D'Amico et al. 2020) that the reference stores as Python lambdas:

* `M22b[0..27](n1, n2)`, `M13b[0..9](n1)`      (reference pybird.py:98-148)
* `Qa`  (Nl=2, NIR=8)                           (reference pybird.py:179-469)
* `Qawithhex` (Nl=3, NIR=16)                    (reference resumfactor.py:595-2377, :4632)

They are re-expressed here as exact polynomial coefficient arrays (symbolic expansion, no
floating point until the final cast):

  m22_num[b, i, j], m22_den[b, i, j] :  M22b_b = sum n1^i n2^j num / sum n1^i n2^j den
  m13_num[b, i],    m13_den[b, i]    :  M13b_b = sum n1^i num / sum n1^i den
  q_nl2[a, l, lp, u, d], q_nl3[...]  :  Q[a][2l][2lp][u](f) = sum_d q f^d
                                         (a is the reference's FIRST index, i.e. "N-j")
"""
from __future__ import annotations

import os
import sys

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import refload  # noqa: E402


def poly2d(expr, x, y):
    num, den = sp.fraction(sp.together(sp.nsimplify(expr, rational=True)))
    pn = sp.Poly(sp.expand(num), x, y)
    pd = sp.Poly(sp.expand(den), x, y)
    return pn, pd


def to_array2d(polys, x, y):
    dmax = max(max(max(m) for m in p.monoms()) for p in polys) + 1
    out = np.zeros((len(polys), dmax, dmax))
    for b, p in enumerate(polys):
        for (i, j), c in zip(p.monoms(), p.coeffs()):
            out[b, i, j] = float(c)
            assert float(c) == c or abs(float(c) - c) < 1e-300, "non-representable coefficient"
    return out


def main():
    ref = refload.load()
    pb, rf = ref.pybird, ref.resumfactor
    n1, n2, f = sp.symbols("n1 n2 f")

    nums, dens = [], []
    for b in range(28):
        pn, pd = poly2d(pb.M22b[b](n1, n2), n1, n2)
        nums.append(pn)
        dens.append(pd)
    m22_num = to_array2d(nums, n1, n2)
    m22_den = to_array2d(dens, n1, n2)

    nums, dens = [], []
    for b in range(10):
        expr = sp.nsimplify(pb.M13b[b](n1), rational=True)
        num, den = sp.fraction(sp.together(expr))
        nums.append(sp.Poly(sp.expand(num), n1))
        dens.append(sp.Poly(sp.expand(den), n1))
    d13 = max(max(p.degree() for p in nums), max(p.degree() for p in dens)) + 1
    m13_num = np.zeros((10, d13))
    m13_den = np.zeros((10, d13))
    for b in range(10):
        for (i,), c in zip(nums[b].monoms(), nums[b].coeffs()):
            m13_num[b, i] = float(c)
        for (i,), c in zip(dens[b].monoms(), dens[b].coeffs()):
            m13_den[b, i] = float(c)

    def qtable(table, Nl, Nn):
        polys = {}
        dmax = 0
        for a in range(2):
            for l in range(Nl):
                for lp in range(Nl):
                    for u in range(Nn):
                        expr = sp.nsimplify(table[a][2 * l][2 * lp][u](f), rational=True)
                        p = sp.Poly(sp.expand(expr), f)
                        polys[a, l, lp, u] = p
                        dmax = max(dmax, p.degree())
        out = np.zeros((2, Nl, Nl, Nn, dmax + 1))
        for key, p in polys.items():
            for (d,), c in zip(p.monoms(), p.coeffs()):
                out[key + (d,)] = float(sp.Rational(c))
        return out

    q_nl2 = qtable(pb.Qa, 2, 32)
    q_nl3 = qtable(rf.Qawithhex, 3, 96)

    mu = np.array([[pb.mu[p][l] for l in (0, 2, 4)] for p in (0, 2, 4, 6, 8)])
    out = os.path.join(HERE, "..", "eftpipe_b200", "data", "pybird_tables.npz")
    np.savez_compressed(
        out, m22_num=m22_num, m22_den=m22_den, m13_num=m13_num, m13_den=m13_den,
        q_nl2=q_nl2, q_nl3=q_nl3, mu_to_legendre=mu,
        kbird=pb.get_kbird(0.3), sbird=pb.sbird,
    )
    print("wrote", out, {k: v.shape for k, v in dict(
        m22_num=m22_num, m22_den=m22_den, m13_num=m13_num, q_nl2=q_nl2, q_nl3=q_nl3).items()})


if __name__ == "__main__":
    main()
