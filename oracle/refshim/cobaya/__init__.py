"""Mini-Cobaya: a small stand-in for the `cobaya` package (absent in this image, no network).

TEST INFRASTRUCTURE ONLY - nothing in the product imports it.  It implements the part of Cobaya's component protocol
that zhaoruiyang98/eftpipe uses, closely enough to run the UNMODIFIED reference end to end
(`EFTLSS -> EFTLeafKernel -> EFTLeaf -> EFTLike`, theory.py / likelihood.py) and, through the very same driver, this
repository's own Cobaya-facing classes (`eftpipe_b200.cobaya`):

  * `cobaya.theory.Theory / HelperTheory / Provider`, `cobaya.likelihood.Likelihood`: construction from an info
    dictionary over the class defaults (`<file_base_name>.yaml` next to the class), `initialize`, `get_requirements`,
    `must_provide` (with requirement redirection through its return value), `get_can_provide[_params]`,
    `get_can_support_params`, `get_helper_theories`, `initialize_with_provider`, `calculate(state, want_derived,
    **params)`, one-deep state caching keyed on the component's own and inherited input parameters (the fast / slow
    split);
  * `cobaya.model.get_model(info)`: component instantiation, parameter assignment, dependency resolution, ordered
    evaluation, `loglikes`, `logposterior`;
  * `cobaya.log`, `cobaya.mpi`, `cobaya.typing`, `cobaya.theories.classy` (an import stub: the reference's
    `eftpipe/__init__.py` imports its CLASS wrapper unconditionally).

What it deliberately is not: samplers, priors beyond bounds bookkeeping, MPI, output files, CLASS / CAMB.
Parameter values may be numpy arrays / tensors of shape (B,) - real Cobaya passes floats; batching is this
repository's extension and the reference is only ever driven with floats.
"""
__version__ = "0.0-mini"
