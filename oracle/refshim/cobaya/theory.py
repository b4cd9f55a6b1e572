"""Mini-Cobaya component protocol (see the package docstring): `CobayaComponent`, `Theory`, `HelperTheory`, `Provider`.
Method names, call order and state layout follow cobaya.theory / cobaya.component 3.x as far as eftpipe relies on them."""
from __future__ import annotations

import importlib
import inspect
import os
from copy import deepcopy

import numpy as np

from .log import HasLogger, LoggedError
from .typing import empty_dict


def _same(a, b):
    """parameter values may be floats or arrays over a batch"""
    if a is b:
        return True
    try:
        if hasattr(a, "detach"):
            a = a.detach().cpu().numpy()
        if hasattr(b, "detach"):
            b = b.detach().cpu().numpy()
        return bool(np.array_equal(np.asarray(a), np.asarray(b)))
    except Exception:
        return False


def _same_dict(a, b):
    return a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)


class CobayaComponent(HasLogger):
    file_base_name = None

    def __init__(self, info=empty_dict, name=None, timing=None, packages_path=None, initialize=True, standalone=True):
        self._name = name or self.__class__.__name__
        self.packages_path = packages_path
        self.set_logger(name=self._name)
        self.timer = None
        for k, v in self.get_defaults().items():  # class defaults (yaml next to the class), then the user's info
            setattr(self, k, deepcopy(v))
        for k, v in (info or {}).items():
            if k not in ("class", "params"):
                setattr(self, k, v)
        if not hasattr(self, "stop_at_error"):
            self.stop_at_error = False
        self._declared_params = dict((info or {}).get("params") or {})
        if initialize:
            self.initialize()

    @classmethod
    def get_defaults(cls):
        """<file_base_name or class name>.yaml in the folder of the class's module, if there is one"""
        out = {}
        for klass in reversed(cls.__mro__):
            base = klass.__dict__.get("file_base_name") or klass.__name__
            try:
                folder = os.path.dirname(inspect.getfile(klass))
            except (TypeError, OSError):
                continue
            path = os.path.join(folder, base + ".yaml")
            if os.path.exists(path):
                import yaml

                with open(path) as fh:
                    out.update(yaml.safe_load(fh) or {})
        return out

    def get_name(self):
        return self._name

    def initialize(self):
        pass

    def close(self):
        pass


class Theory(CobayaComponent):
    def __init__(self, info=empty_dict, name=None, timing=None, packages_path=None, initialize=True, standalone=True):
        self.provider = None
        self.input_params, self.output_params = [], []
        self._current_state = None
        self._input_params_extra = set()
        super().__init__(info, name=name, timing=timing, packages_path=packages_path, initialize=initialize,
                         standalone=standalone)

    # ---- protocol defaults ----
    def get_requirements(self):
        return {}

    def must_provide(self, **requirements):
        return None

    def calculate(self, state, want_derived=True, **params_values_dict):
        pass

    def initialize_with_params(self):
        pass

    def initialize_with_provider(self, provider):
        self.provider = provider

    def get_can_provide(self):
        return []

    def get_can_provide_params(self):
        return []

    def get_can_support_params(self):
        return []

    def get_helper_theories(self):
        return {}

    # ---- results ----
    @property
    def current_state(self):
        return self._current_state

    @property
    def current_derived(self):
        return self._current_state.get("derived") or {}

    def get_result(self, result_name, **kwargs):
        return self._current_state[result_name]

    def get_param(self, p):
        return self._current_state["derived"][p]

    def check_cache_and_compute(self, params_values_dict, dependency_params=None, want_derived=False, cached=True):
        """one-deep cache: recompute only if the component's own parameters or the input parameters of anything it
        depends on changed (Cobaya's fast / slow blocking)"""
        dependency_params = dependency_params or {}
        st = self._current_state
        if cached and st is not None and _same_dict(st["params"], params_values_dict) and \
                _same_dict(st["dependency_params"], dependency_params) and (st["derived"] is not None or not want_derived):
            self.n_cached = getattr(self, "n_cached", 0) + 1
            return True
        state = {"params": dict(params_values_dict), "dependency_params": dict(dependency_params),
                 "derived": {} if want_derived else None}
        try:
            if self.calculate(state, want_derived, **params_values_dict) is False:
                return False
        except LoggedError:
            raise
        self.n_computed = getattr(self, "n_computed", 0) + 1
        self._current_state = state
        return True


class HelperTheory(Theory):
    pass


class Provider:
    """cobaya.theory.Provider: `get_param`, `get_result`, and `get_X(...)` forwarded to the component providing X"""

    def __init__(self, model, requirement_providers):
        self.model = model
        self.requirement_providers = requirement_providers
        self.params = {}

    def set_current_input_params(self, params):
        self.params = params

    def get_param(self, param):
        if isinstance(param, str):
            if param in self.params:
                return self.params[param]
            return self.requirement_providers[param].get_param(param)
        return [self.get_param(p) for p in param]

    def get_result(self, result_name, **kwargs):
        return self.requirement_providers[result_name].get_result(result_name, **kwargs)

    def __getattr__(self, name):
        if name.startswith("get_"):
            comp = self.__dict__.get("requirement_providers", {}).get(name[4:])
            if comp is not None:
                return getattr(comp, name)
        raise AttributeError(name)


def resolve_class(name, cfg=None):
    """`pkg.Class` -> class (Cobaya: component name or its `class:` entry).  A class object passes through."""
    target = (cfg or {}).get("class", name) if isinstance(cfg, dict) else name
    if inspect.isclass(target):
        return target
    module_name, class_name = str(target).rsplit(".", 1)
    return getattr(importlib.import_module(module_name), class_name)
