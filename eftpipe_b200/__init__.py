"""eftpipe_b200 - the PyBird one-loop multipole + marginalised-likelihood hot path of eftpipe on B200 (sm_100a).

Host-side mirror of the reference's interface (`pybird`, `window`, `icc`, `binning`, `chained`, `parambasis`,
`marginal`, `likelihood`, `theory`, `model`, `boltzmann`) over `libeftb200.so` (C ABI in include/eftb200.h).
The CUDA library is required: there is no CPU path (see `_lib.require_cuda`)."""

__version__ = "0.1.0"
