"""EFT parameter bases - mirror of `eftpipe.parambasis` (parambasis.py:30-465): `BirdComponent`,
`reduce_Plk`, `WestCoastBasis`, `EastCoastBasis`, `find_param_basis`, batched on the device.

A basis answers two questions for the CUDA bias-reduction kernel (`like_vectors_kernel`):
  * `kernel_columns(params, f)`: the NPAR = 19 west-coast-form inputs b1A..cr2A, b1B..cr2B, ce0, cemono, cequad
    (parambasis.py:76-83) and the two NNLO counterterm coefficients (cr4, cr6 west / ctilde, 0 east; :96-107) as
    per-point tensors / floats (0 = absent);
  * `gaussian_descriptors(co)`: for every Gaussian (analytically marginalisable) parameter the derivative
    dP/dg as  sum_q coef_q * var_q * term[i_q]  with var = b1A^pa b1B^pb f^pf (`var_code`)
    (parambasis.py:249-316 west, :403-454 east).
"""
from __future__ import annotations

import importlib
from dataclasses import dataclass, field

import numpy as np

NPAR = 19  # EFTB_NPAR of include/eftb200.h


def var_code(pa=0, pb=0, pf=0):
    """b1A^pa * b1B^pb * f^pf as the integer the kernel decodes (eftb200.h `g_var`)"""
    if not (0 <= pa <= 3 and 0 <= pb <= 3 and 0 <= pf <= 7):
        raise ValueError("derivative factor outside b1^<=3 f^<=7")
    return pa | pb << 2 | pf << 4


VAR_ONE, VAR_B1A, VAR_B1B, VAR_F, VAR_F2 = var_code(), var_code(pa=1), var_code(pb=1), var_code(pf=1), var_code(pf=2)
T11, TCT, TLOOP, TST, TNNLO = 0, 3, 9, 21, 24  # offsets of P11l, Pctl, Ploopl, Pstl, PctNNLOl rows on the term axis


@dataclass
class BirdComponent:
    """parambasis.py:30-39 (arrays are (B, No, nk) device tensors)."""

    Plin: object
    Ploop: object
    Pct: object
    Pst: object
    Picc: object

    def sum(self):
        return self.Plin + self.Ploop + self.Pct + self.Pst + self.Picc


def _stoch_factors(co):
    x1 = 0.5 * (1.0 / co.ndA + 1.0 / co.ndB)
    x2 = 0.5 * (1.0 / co.ndA / co.kmA**2 + 1.0 / co.ndB / co.kmB**2)
    return x1, x2


@dataclass(frozen=True)
class WestCoastBasis:
    prefix: str = ""
    cross_prefix: list = field(default_factory=list)

    def default(self):
        return {p: 0.0 for p in self.gaussian_params()}

    def is_cross(self):
        return bool(self.cross_prefix)

    def bsA(self):
        pre = self.cross_prefix[0] if self.is_cross() else self.prefix
        return [pre + p for p in ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2")]

    def bsB(self):
        if not self.is_cross():
            return []
        return [self.cross_prefix[1] + p for p in ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2")]

    def es(self):
        return [self.prefix + p for p in ("ce0", "cemono", "cequad")]

    def cnnloA(self):
        return [self.prefix + p for p in ("cr4", "cr6")]

    @classmethod
    def get_name(cls):
        return "westcoast"

    @classmethod
    def counterform(cls):
        return "westcoast"

    def non_gaussian_params(self):
        names = ("b1", "b2", "b4")
        if self.is_cross():
            return [x + p for x in self.cross_prefix for p in names]
        return [self.prefix + p for p in names]

    def gaussian_params(self):
        names, st = ("b3", "cct", "cr1", "cr2"), ("ce0", "cemono", "cequad")
        if self.is_cross():
            return [x + p for x in self.cross_prefix for p in names] + [self.prefix + p for p in st]
        return [self.prefix + p for p in names + st] + self.cnnloA()

    # ---- kernel interface ----
    def kernel_columns(self, params, f=None):
        get = lambda n: params.get(n, 0.0)
        A = [get(n) for n in self.bsA()]
        Bv = [get(n) for n in self.bsB()] if self.is_cross() else A
        # the NNLO coefficients carry the tracer's own prefix, also for a cross (parambasis.py:192-193, :242)
        return A + Bv + [get(n) for n in self.es()] + [get(n) for n in self.cnnloA()]

    def gaussian_descriptors(self, co):
        x1, x2 = _stoch_factors(co)
        out = {}
        if self.is_cross():
            pa, pb = self.cross_prefix
            for pre, other, km, kr in ((pa, VAR_B1B, co.kmA, co.krA), (pb, VAR_B1A, co.kmB, co.krB)):
                out[pre + "b3"] = [(TLOOP + 3, VAR_ONE, 0.5), (TLOOP + 7, other, 0.5)]
                out[pre + "cct"] = [(TCT + 0, other, 1.0 / km**2), (TCT + 3, VAR_F, 1.0 / km**2)]
                out[pre + "cr1"] = [(TCT + 1, other, 1.0 / kr**2), (TCT + 4, VAR_F, 1.0 / kr**2)]
                out[pre + "cr2"] = [(TCT + 2, other, 1.0 / kr**2), (TCT + 5, VAR_F, 1.0 / kr**2)]
        else:
            pre, km, kr = self.prefix, co.kmA, co.krA
            out[pre + "b3"] = [(TLOOP + 3, VAR_ONE, 1.0), (TLOOP + 7, VAR_B1A, 1.0)]
            out[pre + "cct"] = [(TCT + 0, VAR_B1A, 2.0 / km**2), (TCT + 3, VAR_F, 2.0 / km**2)]
            out[pre + "cr1"] = [(TCT + 1, VAR_B1A, 2.0 / kr**2), (TCT + 4, VAR_F, 2.0 / kr**2)]
            out[pre + "cr2"] = [(TCT + 2, VAR_B1A, 2.0 / kr**2), (TCT + 5, VAR_F, 2.0 / kr**2)]
            if co.with_NNLO:  # parambasis.py:303-307 (auto spectra only, like the reference)
                out[pre + "cr4"] = [(TNNLO + 0, var_code(pa=2), 0.25 / kr**4)]
                out[pre + "cr6"] = [(TNNLO + 1, VAR_B1A, 0.25 / kr**4)]
        out[self.prefix + "ce0"] = [(TST + 0, VAR_ONE, x1)]
        out[self.prefix + "cemono"] = [(TST + 1, VAR_ONE, x2)]
        out[self.prefix + "cequad"] = [(TST + 2, VAR_ONE, x2)]
        return out

    # ---- reference API ----
    def reduce_Plk(self, bird, params_values_dict):
        from .likelihood import reduce_on_device

        return reduce_on_device(self, bird, params_values_dict)[0]

    def reduce_Plk_gaussian_table(self, bird, params_values_dict, requires=None):
        from .likelihood import reduce_on_device

        table = reduce_on_device(self, bird, params_values_dict, want_table=True)[1]
        return {k: v for k, v in table.items() if requires is None or k in requires}


@dataclass(frozen=True)
class EastCoastBasis:
    """parambasis.py:319-454 (arXiv:2106.12580, 2208.05929); auto-spectra only, like the reference."""

    prefix: str = ""
    cross_prefix: list = field(default_factory=list)

    def __post_init__(self):
        if self.cross_prefix:
            raise NotImplementedError("EastCoastBasis does not support cross yet")

    def default(self):
        return {p: 0.0 for p in self.gaussian_params()}

    def is_cross(self):
        return False

    def bsA(self):
        return [self.prefix + p for p in ("b1", "b2", "bG2", "bGamma3", "c0", "c2", "c4")]

    def es(self):
        return [self.prefix + p for p in ("Pshot", "a0", "a2")]

    def cnnloA(self):
        return [self.prefix + "ctilde"]

    @classmethod
    def get_name(cls):
        return "eastcoast"

    @classmethod
    def counterform(cls):
        return "eastcoast"

    def non_gaussian_params(self):
        return [self.prefix + p for p in ("b1", "b2", "bG2")]

    def gaussian_params(self):
        return [self.prefix + p for p in ("bGamma3", "c0", "c2", "c4", "Pshot", "a0", "a2")] + self.cnnloA()

    def kernel_columns(self, params, f=None):
        get = lambda n: params.get(self.prefix + n, 0.0)
        b1, b2, bG2, bGamma3, c0, c2, c4 = (get(n) for n in ("b1", "b2", "bG2", "bGamma3", "c0", "c2", "c4"))
        A = [b1, b1 + 7 / 2 * bG2, b1 + 15 * bG2 + 6 * bGamma3, 1 / 2 * b2 - 7 / 2 * bG2,
             c0 - f / 3 * c2 + 3 / 35 * f**2 * c4, c2 - 6 / 7 * f * c4, c4]  # parambasis.py:384-392
        Pshot, a0, a2 = get("Pshot"), get("a0"), get("a2")
        return A + A + [Pshot, a0 + 1 / 3 * a2, 2 / 3 * a2, get("ctilde"), 0.0]  # :395-398

    def gaussian_descriptors(self, co):
        x1, x2 = _stoch_factors(co)
        pre = self.prefix
        nnlo = {}
        if co.with_NNLO:  # parambasis.py:439-445
            nnlo[pre + "ctilde"] = [(TNNLO + 0, var_code(pa=2, pf=4), -1.0), (TNNLO + 1, var_code(pa=1, pf=5), -2.0),
                                    (TNNLO + 2, var_code(pf=6), -1.0)]
        return {
            pre + "bGamma3": [(TLOOP + 3, VAR_ONE, 6.0), (TLOOP + 7, VAR_B1A, 6.0)],
            pre + "c0": [(TCT + 0, VAR_ONE, -2.0)],
            pre + "c2": [(TCT + 0, VAR_F, 2.0 / 3.0), (TCT + 1, VAR_F, -2.0)],
            pre + "c4": [(TCT + 0, VAR_F2, -6.0 / 35.0), (TCT + 1, VAR_F2, 12.0 / 7.0), (TCT + 2, VAR_F2, -2.0)],
            pre + "Pshot": [(TST + 0, VAR_ONE, x1)],
            pre + "a0": [(TST + 1, VAR_ONE, x2)],
            pre + "a2": [(TST + 1, VAR_ONE, x2 / 3.0), (TST + 2, VAR_ONE, 2.0 * x2 / 3.0)],
            **nnlo,
        }

    def reduce_Plk(self, bird, params_values_dict):
        from .likelihood import reduce_on_device

        return reduce_on_device(self, bird, params_values_dict)[0]

    def reduce_Plk_gaussian_table(self, bird, params_values_dict, requires=None):
        from .likelihood import reduce_on_device

        table = reduce_on_device(self, bird, params_values_dict, want_table=True)[1]
        return {k: v for k, v in table.items() if requires is None or k in requires}


def reduce_Plk(bird, bsA, bsB=None, es=(0.0, 0.0, 0.0), cnnloA=(0.0, 0.0), cnnloB=None):
    """Function form (parambasis.py:42-136): explicit bias lists instead of a parameter dictionary."""
    from .likelihood import reduce_on_device

    names = ("b1", "b2", "b3", "b4", "cct", "cr1", "cr2")
    if bsB is None:
        basis = WestCoastBasis(prefix="")
        params = dict(zip(names, bsA))
    else:
        basis = WestCoastBasis(prefix="X_", cross_prefix=["A_", "B_"])
        params = {**{"A_" + n: v for n, v in zip(names, bsA)}, **{"B_" + n: v for n, v in zip(names, bsB)}}
    params.update({basis.prefix + n: v for n, v in zip(("ce0", "cemono", "cequad"), es)})
    if bird.co.with_NNLO:
        if bird.co.counterform == "eastcoast":
            # the function form reads ctilde = cnnloA[0] and takes bsA as already-converted west-form values (:102-105)
            raise NotImplementedError("function-form reduce_Plk with counterform='eastcoast' and NNLO: use EastCoastBasis.reduce_Plk")
        params.update({basis.prefix + n: v for n, v in zip(("cr4", "cr6"), cnnloA)})
    return reduce_on_device(basis, bird, params)[0]


class _UnitBird:
    """A numpy birdlike (pybird.py:598-608) whose term arrays are unit vectors along a probe axis: what a linear reduction
    returns for it are its coefficients"""

    def __init__(self, co, f, nterm):
        self.co, self.f = co, f
        eye = np.eye(nterm)
        row = lambda a, b: np.ascontiguousarray(eye[a:b][None])  # (1, n, nterm): multipole axis of length 1
        self.P11l, self.Pctl, self.Ploopl, self.Pstl = row(0, 3), row(3, 9), row(9, 21), row(21, 24)
        self.PctNNLOl = row(24, 27) if nterm > 24 else np.zeros((1, 3, nterm))
        self.Picc = np.zeros((1, nterm))


class ProbedBasis:
    """A reference-style `EFTBasis` (parambasis.py:139-162) given by dotted path - numpy code with `reduce_Plk(bird, params)
    -> BirdComponent` and `reduce_Plk_gaussian_table(bird, params, requires)` - on the batched device path.

    Both reductions are linear in the term arrays, with coefficients that depend on the point (its parameters and growth
    rate) only.  Per point, the basis' own code is run once on a unit birdlike; what comes back are the coefficient of every
    term row in P_l(k) and in every Gaussian-parameter derivative.  These go to the device as explicit bias columns
    (eftb_like_constants.mode = 1) and the kernels do the arithmetic on the real term arrays.  A basis whose reduction is
    not linear, or depends on the multipole, is refused when it is first probed."""

    kernel_mode = "explicit"

    def __init__(self, inner):
        self.inner = inner
        self.prefix, self.cross_prefix = inner.prefix, list(getattr(inner, "cross_prefix", []) or [])

    def __getattr__(self, name):  # get_name, counterform, gaussian_params, non_gaussian_params, default, ...
        return getattr(self.inner, name)

    def counterform(self):
        return self.inner.counterform()

    def is_cross(self):
        return bool(self.cross_prefix)

    def explicit_columns(self, params, f, B, co, nterm, gaussian):
        """(bias (B, nterm), {gaussian name: (B, nterm)}) for the B points of `params` (scalars or (B,) arrays)"""
        from types import SimpleNamespace

        host = lambda v: v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v, float)
        cols = {k: np.broadcast_to(host(v).reshape(-1), (B,)) if np.ndim(host(v)) else np.full(B, float(host(v))) for k, v in params.items()}
        fB = np.broadcast_to(host(f).reshape(-1)[:B] if np.ndim(host(f)) else np.full(B, float(host(f))), (B,))
        pco = SimpleNamespace(**{k: getattr(co, k) for k in ("kmA", "krA", "ndA", "kmB", "krB", "ndB", "counterform", "with_NNLO")}, No=1)
        bias = np.zeros((B, nterm))
        table = {g: np.zeros((B, nterm)) for g in gaussian}
        mine = set(self.inner.gaussian_params())
        for i in range(B):
            p = {k: float(v[i]) for k, v in cols.items()}
            bird = _UnitBird(pco, float(fB[i]), nterm)
            bias[i] = np.asarray(self.inner.reduce_Plk(bird, p).sum(), float).reshape(-1)
            if gaussian:
                want = [g for g in gaussian if g in mine]
                full = dict(self.inner.default(), **p) if hasattr(self.inner, "default") else p
                tab = self.inner.reduce_Plk_gaussian_table(bird, full, requires=want) if want else {}
                for g, arr in tab.items():
                    if g in table:
                        table[g][i] = np.asarray(arr, float).reshape(-1)
        return bias, table

    # ---- reference API on device birds ----
    def reduce_Plk(self, bird, params_values_dict):
        from .likelihood import reduce_on_device

        return reduce_on_device(self, bird, params_values_dict)[0]

    def reduce_Plk_gaussian_table(self, bird, params_values_dict, requires=None):
        from .likelihood import reduce_on_device

        table = reduce_on_device(self, bird, params_values_dict, want_table=True)[1]
        return {k: v for k, v in table.items() if requires is None or k in requires}


def find_param_basis(name: str):
    """parambasis.py:457-465.  A class by dotted path that is not one of this package's own bases is wrapped into a
    `ProbedBasis` when it is instantiated."""
    if name == "westcoast":
        return WestCoastBasis
    if name == "eastcoast":
        return EastCoastBasis
    module_name, class_name = name.rsplit(".", 1)
    cls = getattr(importlib.import_module(module_name), class_name)
    if hasattr(cls, "kernel_columns"):
        return cls

    def construct(prefix="", cross_prefix=None):
        return ProbedBasis(cls(prefix=prefix, cross_prefix=list(cross_prefix or [])))

    construct.__name__ = class_name
    return construct
