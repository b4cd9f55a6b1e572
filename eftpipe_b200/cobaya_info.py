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
