"""`BirdTransformer` protocol and `PlainBird` (reference transformer.py:11-38), batched on the device."""
from __future__ import annotations

from .pybird import _TermsView


class PlainBird(_TermsView):
    """Result of Binning / Chained transforms: same read-only term accessors as `Bird`
    (P11l, Pctl, Ploopl, Pstl, PctNNLOl, Picc), arrays shaped (B, Nl_out, nrow, nk_out)."""

    def __init__(self, f, co, T, picc, B, squeeze=False, f_bm=None):
        self.f, self.co, self._T, self._picc, self.B, self._squeeze = f, co, T, picc, B, squeeze
        self._f_bm = f_bm


class BirdCopier:
    """transformer.py:27-38"""

    def transform(self, birdlike):
        return PlainBird(birdlike.f, birdlike.co, birdlike._T.clone(),
                         None if birdlike._picc is None else birdlike._picc.copy(), birdlike.B, birdlike._squeeze,
                         getattr(birdlike, "_f_bm", None))


def f_batch_minor(birdlike):
    """growth rate as a (Bp,) batch-minor tensor, for the bias-reduction kernels."""
    fb = getattr(birdlike, "_f_bm", None)
    if fb is None:
        fb = birdlike._bm_scalar("f")
    return fb
