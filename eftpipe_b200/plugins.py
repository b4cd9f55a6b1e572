"""Custom stage plugins - the reference's "window by dotted path" seam (theory.py:62-72, :370-377): `with_window:
"pkg.mod.Class"` names a class constructed as `Class(**cfg, co=co, icc=icc, name=...)` whose `.Window(bird)` mutates
the Bird's term arrays in place.

Any such stage is linear in the term arrays (plus a constant on `Picc`), and on the batched path a linear stage is a
fixed operator composed into the tracer's projection.  So a custom plugin is *probed* once at plan build: its
`.Window` is applied to host-side numpy probe birds that carry the unit vectors of the (multipole, k-node) grid, and the
responses are the columns of the operator.  The plugin itself stays ordinary reference-style numpy code; it never
runs per evaluation.
"""
from __future__ import annotations

import functools
import importlib

import numpy as np


def find_window_constructor(name: str):
    """theory.py:62-72: "auto" / "default" -> the built-in Window, otherwise a dotted path to a class"""
    if name in ("auto", "default"):
        from .window import Window

        return Window
    module_name, class_name = name.rsplit(".", 1)
    try:
        module = importlib.import_module(module_name)
    except ModuleNotFoundError:
        module_name, tmp = module_name.rsplit(".", 1)
        class_name = f"{tmp}.{class_name}"
        module = importlib.import_module(module_name)
    return functools.reduce(getattr, class_name.split("."), module)


class ProbeBird:
    """Host-side stand-in for a Bird with the attributes of the `BirdLike` protocol (pybird.py:598-608)"""

    def __init__(self, co, f=0.0):
        Nl, Nk = co.Nl, co.Nk
        self.co, self.f = co, f
        self.P11l = np.zeros((Nl, 3, Nk))
        self.Pctl = np.zeros((Nl, 6, Nk))
        self.Ploopl = np.zeros((Nl, 12, Nk))
        self.Pstl = np.zeros((Nl, 3, Nk))
        self.PctNNLOl = np.zeros((Nl, 3, Nk))
        self.Picc = np.zeros((Nl, Nk))
        self.snapshots = {}

    def create_snapshot(self, name):
        pass


def probe_linear_stage(apply, co):
    """Operator of an in-place linear stage `apply(bird)`.

    Returns dict(matrix=(Na, Nk', Nl, Nk), matrix_st=(Na, Nk', Nl, Nk) or None when the stochastic terms see the same
    operator, picc=(Na, Nk') the constant it adds to Picc).  The 21 rows of P11l | Pctl | Ploopl carry 21 unit vectors
    per call; Pstl is probed separately (window_st / fiberst style exemptions); linearity is verified on a random
    combination."""
    Nl, Nk = co.Nl, co.Nk
    n = Nl * Nk
    cols = []
    names = (("P11l", 3), ("Pctl", 6), ("Ploopl", 12))
    per_call = sum(c for _, c in names)
    shape = None
    for start in range(0, n, per_call):
        bird = ProbeBird(co)
        idx = start
        for name, cnt in names:
            arr = getattr(bird, name)
            for i in range(cnt):
                if idx < n:
                    arr[idx // Nk, i, idx % Nk] = 1.0
                idx += 1
        apply(bird)
        idx = start
        for name, cnt in names:
            arr = np.asarray(getattr(bird, name))
            for i in range(cnt):
                if idx < n:
                    cols.append(arr[:, i, :].copy())
                idx += 1
    shape = cols[0].shape  # (Na, Nk')
    matrix = np.stack(cols, axis=-1).reshape(shape + (Nl, Nk))
    # constant part and the stochastic rows
    zero = ProbeBird(co)
    apply(zero)
    picc = np.asarray(zero.Picc, float).copy()
    st_cols = []
    for start in range(0, n, 3):
        bird = ProbeBird(co)
        for i in range(3):
            if start + i < n:
                bird.Pstl[(start + i) // Nk, i, (start + i) % Nk] = 1.0
        apply(bird)
        for i in range(3):
            if start + i < n:
                st_cols.append(np.asarray(bird.Pstl)[:, i, :].copy())
    if st_cols[0].shape != shape:
        raise ValueError("the stage maps the stochastic terms to a different grid than the other terms")
    matrix_st = np.stack(st_cols, axis=-1).reshape(shape + (Nl, Nk))
    # linearity / row-independence check on a random bird
    rng = np.random.default_rng(0)
    bird = ProbeBird(co)
    for name, _ in names:
        getattr(bird, name)[...] = rng.normal(size=getattr(bird, name).shape)
    ref = {name: getattr(bird, name).copy() for name, _ in names}
    apply(bird)
    for name, _ in names:
        want = np.einsum("akln,lin->aik", matrix, ref[name])
        got = np.asarray(getattr(bird, name))
        if not np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max()):
            raise ValueError(f"the stage is not a fixed linear map of {name}: it cannot be composed into the plan")
    same = np.allclose(matrix_st, matrix, rtol=1e-12, atol=1e-12 * np.abs(matrix).max())
    return dict(matrix=matrix, matrix_st=None if same else matrix_st, picc=picc)
