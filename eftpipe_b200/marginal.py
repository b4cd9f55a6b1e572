"""Analytic Gaussian marginalisation - mirror of `eftpipe.marginal.Marginalizable` (marginal.py:31-232).

The per-point algebra (F2, F1, F0, Cholesky, log-det, back-substitution; marginal.py:100-137) runs in
`like_finish_kernel` (csrc/like.cu).  This module keeps the reference's prior handling: `update_prior`
(sorting, infinite scales all-or-none, marginal.py:198-232) and the `mu_G` / `sigma_inv` construction
(marginal.py:60-77).  Callable (string) priors, which the reference `eval`s per evaluation, are supported
for host-side evaluation only when they do not depend on sampled parameters."""
from __future__ import annotations

import numpy as np


class LoggedError(Exception):
    pass


def valid_prior_config(config) -> bool:
    if config is None:
        return True
    return isinstance(config, dict) and ("loc" in config or "scale" in config)


class Marginalizable:
    valid_prior: dict

    def marginalizable_params(self) -> list:
        raise NotImplementedError

    def update_prior(self, prior: dict) -> dict:
        allowed = self.marginalizable_params()
        for key in prior:
            if key not in allowed:
                raise LoggedError(f"key <{key}> is not marginalizable")
        new, ninf = {}, 0
        for name, dct in prior.items():
            loc = None if dct is None else dct.get("loc", None)
            scale = None if dct is None else dct.get("scale", None)
            if scale is None or scale == np.inf:
                scale = np.inf
                ninf += 1
            new[name] = {"loc": 0 if loc is None else loc, "scale": scale}
        order = sorted(new, key=allowed.index)
        out = {name: new[name] for name in order}
        if ninf != 0 and ninf != len(out):
            raise LoggedError("only support setting infinite scale for all parameters")
        return out

    def setup_prior(self, prior: dict) -> None:
        self.valid_prior = self.update_prior(prior)

    @property
    def mu_G(self):
        locs = [d["loc"] for d in self.valid_prior.values()]
        if any(isinstance(x, str) for x in locs):
            raise NotImplementedError("callable prior locations are not supported on the batched device path")
        return np.array(locs, dtype=np.float64)

    @property
    def sigma_inv(self):
        std = [d["scale"] for d in self.valid_prior.values()]
        if any(isinstance(x, str) for x in std):
            raise NotImplementedError("callable prior scales are not supported on the batched device path")
        n = len(std)
        if np.inf in std:
            return np.zeros((n, n))  # marginal.py:74-75
        return np.diag(1.0 / np.array(std, dtype=np.float64) ** 2)
