"""Analytic Gaussian marginalisation - mirror of `eftpipe.marginal.Marginalizable` (marginal.py:31-232).

The per-point algebra (F2, F1, F0, Cholesky, log-det, back-substitution; marginal.py:100-137) runs in
`like_finish_kernel` (csrc/like.cu).  This module keeps the reference's prior handling: `update_prior`
(sorting, infinite scales all-or-none, marginal.py:198-232) and the `mu_G` / `sigma_inv` construction
(marginal.py:60-77).  Callable (string) priors, which the reference `eval`s at every evaluation against the sampled
EFT parameters (marginal.py:13-20, likelihood.py:560-564), are evaluated here once per batch with the parameters as
arrays (`point_priors`) and handed to the kernel as a per-point location / inverse variance."""
from __future__ import annotations

import inspect

import numpy as np


def eval_callable(s: str, env: dict):
    """marginal.py:13-20: `eval` the string, call it with the entries of `env` its positional arguments name"""
    fn = eval(s, env)
    argnames = inspect.getfullargspec(fn).args
    return fn(*(env[p] for p in argnames))


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

    def has_callable_prior(self) -> bool:
        return any(isinstance(d["loc"], str) or isinstance(d["scale"], str) for d in self.valid_prior.values())

    def point_priors(self, env: dict, B: int):
        """(loc (B, nG), diag Sigma^-1 (B, nG)) for a batch: marginal.py:60-77 with every entry of `env` an array of B
        values (or a scalar).  As in the reference, one infinite scale at a point switches the whole prior off there."""
        env = dict(env)
        env.setdefault("np", np)
        col = lambda x: np.broadcast_to(np.asarray(eval_callable(x, env) if isinstance(x, str) else x, dtype=np.float64), (B,))
        loc = np.stack([col(d["loc"]) for d in self.valid_prior.values()], axis=1)
        std = np.stack([col(d["scale"]) for d in self.valid_prior.values()], axis=1)
        with np.errstate(divide="ignore"):
            sinv = 1.0 / std**2
        sinv[np.isinf(std).any(axis=1)] = 0.0  # marginal.py:74-75
        return np.ascontiguousarray(loc), np.ascontiguousarray(sinv)

    @property
    def mu_G(self):
        locs = [d["loc"] for d in self.valid_prior.values()]
        if any(isinstance(x, str) for x in locs):
            raise TypeError("callable prior locations depend on the point: use point_priors(env, B)")
        return np.array(locs, dtype=np.float64)

    @property
    def sigma_inv(self):
        std = [d["scale"] for d in self.valid_prior.values()]
        if any(isinstance(x, str) for x in std):
            raise TypeError("callable prior scales depend on the point: use point_priors(env, B)")
        n = len(std)
        if np.inf in std:
            return np.zeros((n, n))  # marginal.py:74-75
        return np.diag(1.0 / np.array(std, dtype=np.float64) ** 2)
