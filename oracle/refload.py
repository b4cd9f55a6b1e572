"""Import the UNMODIFIED reference (zhaoruiyang98/eftpipe at /root/reference) for
golden-vector generation.  TEST INFRASTRUCTURE ONLY - build container only; the GPU box
has no /root/reference and nothing at run time may depend on this module.

Recipe: a 4-module `cobaya` stub on sys.path plus a synthetic
`eftpipe` package object whose __path__ points at the reference tree, so that sub-modules
import without executing eftpipe/__init__.py (which pulls in CLASS through cobaya).
"""
from __future__ import annotations

import os
import sys
import types
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
# the reference tree: $EFTPIPE_REFERENCE, /root/reference (build container), or the unmodified install that
# baseline/install_ref.sh leaves under baseline/_ref (git-ignored; this is what exists on the GPU box)
_CANDIDATES = [os.environ.get("EFTPIPE_REFERENCE"), "/root/reference", os.path.join(_HERE, "..", "baseline", "_ref")]
REFERENCE_ROOT = next((os.path.abspath(c) for c in _CANDIDATES if c and os.path.isdir(os.path.join(c, "eftpipe", "pybird"))),
                      "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "eftpipe", "pybird"))


def load():
    """Return a namespace with the reference modules (pybird, fftlog, window, ...)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import importlib

    if "eftpipe" not in sys.modules or not hasattr(sys.modules["eftpipe"], "__path__"):
        # the package is imported the normal way, running its unmodified __init__.py (the shim carries an import stub for
        # cobaya.theories.classy); should that fail, sub-modules are still importable through a bare package object
        sys.path.insert(0, REFERENCE_ROOT)
        try:
            importlib.import_module("eftpipe")
        except Exception:
            sys.modules.pop("eftpipe", None)
            pkg = types.ModuleType("eftpipe")
            pkg.__path__ = [os.path.join(REFERENCE_ROOT, "eftpipe")]
            sys.modules["eftpipe"] = pkg
        finally:
            sys.path.remove(REFERENCE_ROOT)

    ns = types.SimpleNamespace()
    ns.pybird = importlib.import_module("eftpipe.pybird.pybird")
    ns.fftlog = importlib.import_module("eftpipe.pybird.fftlog")
    ns.resumfactor = importlib.import_module("eftpipe.pybird.resumfactor")
    ns.window = importlib.import_module("eftpipe.window")
    ns.binning = importlib.import_module("eftpipe.binning")
    ns.chained = importlib.import_module("eftpipe.chained")
    ns.parambasis = importlib.import_module("eftpipe.parambasis")
    ns.marginal = importlib.import_module("eftpipe.marginal")
    ns.transformer = importlib.import_module("eftpipe.transformer")
    try:
        ns.icc = importlib.import_module("eftpipe.icc")
    except Exception as ex:  # numba / pandas API drift: ICC apply step is still restated
        ns.icc = None
        ns.icc_error = repr(ex)
    try:
        ns.likelihood = importlib.import_module("eftpipe.likelihood")
    except Exception as ex:
        ns.likelihood = None
        ns.likelihood_error = repr(ex)
    try:
        ns.theory = importlib.import_module("eftpipe.theory")
        ns.boltzmann = importlib.import_module("eftpipe.boltzmann")
    except Exception as ex:
        ns.theory = None
        ns.theory_error = repr(ex)
    pkg = sys.modules["eftpipe"]  # the names eftpipe/__init__.py exports (Cobaya resolves `eftpipe.eftlss` through them)
    if not hasattr(pkg, "eftlss") and ns.theory is not None:
        pkg.eftlss = ns.theory.EFTLSS
    if not hasattr(pkg, "eftlike") and ns.likelihood is not None:
        pkg.eftlike = ns.likelihood.EFTLike
    return ns
