"""eftpipe_b200 - the PyBird one-loop multipole + marginalised-likelihood hot path of eftpipe on B200 (sm_100a).

Host-side mirror of the reference's interface (`pybird`, `window`, `icc`, `binning`, `chained`, `parambasis`,
`marginal`, `likelihood`, `theory`, `model`, `boltzmann`) over `libeftb200.so` (C ABI in include/eftb200.h).
The CUDA library is required: there is no CPU path (see `_lib.require_cuda`)."""

__version__ = "0.1.0"


def __getattr__(name):
    """`eftpipe_b200.eftlss` / `eftpipe_b200.eftlike`: the Cobaya components (eftpipe/__init__.py:2-4), imported on first use
    because they need the `cobaya` package"""
    if name in ("eftlss", "eftlike"):
        from . import cobaya as _cobaya

        return getattr(_cobaya, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
