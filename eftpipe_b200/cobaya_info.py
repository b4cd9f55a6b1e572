"""Build the batched theory + likelihoods straight from a Cobaya input (the reference's yaml files, e.g.
cobaya/yamls/DR16_noric_LEX_NS_LP024_kmax0.20_EQ02_kmax0.20_XP024_kmax0.20.yaml): the `theory: eftpipe.eftlss: tracers:`
block configures `theory.EFTLSS`, every `likelihood:` entry of class `eftpipe.eftlike` becomes a `likelihood.EFTLike`.
The Boltzmann provider and the sampler blocks are not used: the linear power is an input of `EFTLSS.calculate`.
"""
from __future__ import annotations

import os
from copy import deepcopy

from .likelihood import EFTLike
from .theory import EFTLSS

_LIKE_KEYS = ("tracers", "data", "cov", "chained", "with_binning", "binning", "marg", "jeffreys", "with_interp",
              "likelihood_prefix", "marg_param_prefix")


def _resolve(node, root, cache_dir):
    """relative paths (keys `path`, `*_file`) are taken relative to `root`; fourier-space cache files may be redirected"""
    if isinstance(node, dict):
        out = {}
        for k, v in node.items():
            if isinstance(v, str) and (k == "path" or k.endswith("_file")):
                if cache_dir is not None and k.endswith("fourier_file"):
                    v = os.path.join(cache_dir, os.path.basename(v))
                elif not os.path.isabs(v):
                    v = os.path.normpath(os.path.join(root, v))
                out[k] = v
            elif k == "path" and isinstance(v, list):
                out[k] = [x if os.path.isabs(x) else os.path.normpath(os.path.join(root, x)) for x in v]
            else:
                out[k] = _resolve(v, root, cache_dir)
        return out
    return node


def load_info(info, root=None, cache_dir=None):
    """info: the Cobaya input dictionary or a path to its yaml.  Returns (EFTLSS, {likelihood name: EFTLike}), initialised
    and bound to each other; `root` is the directory relative paths refer to (default: the yaml's directory or cwd)."""
    if isinstance(info, (str, os.PathLike)):
        import yaml

        root = root or os.path.dirname(os.path.abspath(info))
        with open(info) as fh:
            info = yaml.safe_load(fh)
    root = root or os.getcwd()
    block = next(v for k, v in info["theory"].items() if k.split(".")[-1].lower() == "eftlss")
    tracers = _resolve(deepcopy(block["tracers"]), root, cache_dir)
    th = EFTLSS(tracers, cache_dir_path=block.get("cache_dir_path"))
    likes = {}
    for name, cfg in info.get("likelihood", {}).items():
        cfg = cfg or {}
        if str(cfg.get("class", name)).split(".")[-1].lower() != "eftlike":
            continue
        cfg = _resolve(deepcopy(cfg), root, cache_dir)
        likes[name] = EFTLike(**{k: cfg[k] for k in _LIKE_KEYS if k in cfg})
        th.must_provide(likes[name].get_requirements())
    th.initialize()
    for like in likes.values():
        like.initialize_with_provider(th)
    return th, likes


def resolve_params(info, sampled):
    """Evaluate the `params:` block of a Cobaya input for a batch of sampled points: fixed `value: <number>` entries and
    derived-input entries `value: 'lambda a, b: ...'` (e.g. b2, b4 from c2, c4) are computed from `sampled` (dict name ->
    scalar or (B,) array), the way Cobaya feeds them to the theory.  Returns a dict with everything that could be resolved;
    output-only `derived:` lambdas are ignored."""
    import inspect

    import numpy as np

    out = {k: np.asarray(v, float) for k, v in sampled.items()}
    pending = {}
    for name, spec in (info.get("params") or {}).items():
        if name in out:
            continue
        if isinstance(spec, (int, float)):
            out[name] = np.asarray(float(spec))
        elif isinstance(spec, dict) and "value" in spec:
            v = spec["value"]
            if isinstance(v, (int, float)):
                out[name] = np.asarray(float(v))
            elif isinstance(v, str) and v.strip().startswith("lambda"):
                pending[name] = eval(v, {"np": np, "numpy": np})  # noqa: S307 - the user's own input file, as Cobaya does
    progress = True
    while pending and progress:
        progress = False
        for name, fn in list(pending.items()):
            args = list(inspect.signature(fn).parameters)
            if all(a in out for a in args):
                out[name] = np.asarray(fn(*[out[a] for a in args]), float)
                del pending[name]
                progress = True
    return out
