"""Mini-Cobaya model: `get_model(info)` builds the components named in `info["theory"]` and `info["likelihood"]`
(+ their helper theories), assigns parameters, resolves requirements into providers and an evaluation order, and
evaluates log-likelihoods.  See the package docstring for scope."""
from __future__ import annotations

import inspect
from dataclasses import dataclass, field

import numpy as np

from .likelihood import Likelihood
from .log import LoggedError
from .theory import Provider, Theory, resolve_class


@dataclass
class LogPosterior:
    logpost: object = None
    logpriors: list = field(default_factory=list)
    loglikes: list = field(default_factory=list)
    derived: dict = field(default_factory=dict)

    @property
    def loglike(self):
        return sum(self.loglikes)


class Parameterization:
    def __init__(self, params):
        self._sampled, self._fixed, self._derived, self._lambdas = {}, {}, {}, {}
        for name, spec in (params or {}).items():
            if isinstance(spec, (int, float)):
                self._fixed[name] = float(spec)
            elif isinstance(spec, str) and spec.strip().startswith("lambda"):
                self._lambdas[name] = eval(spec, {"np": np})  # noqa: S307 - the user's own input, as Cobaya does
            elif isinstance(spec, dict) and "prior" in spec:
                self._sampled[name] = spec
            elif isinstance(spec, dict) and "value" in spec:
                v = spec["value"]
                if isinstance(v, str):
                    self._lambdas[name] = eval(v, {"np": np})  # noqa: S307
                else:
                    self._fixed[name] = float(v)
            else:  # {"latex": ...} / {"derived": True} / None: an output parameter
                self._derived[name] = spec

    def sampled_params(self):
        return dict(self._sampled)

    def input_params(self):
        return list(self._sampled) + list(self._fixed) + list(self._lambdas)

    def derived_params(self):
        return list(self._derived)

    def to_input(self, point):
        """sampled point (dict, or sequence in `sampled_params` order) -> all input parameters"""
        if not isinstance(point, dict):
            point = dict(zip(self._sampled, point))
        missing = set(self._sampled) - set(point)
        if missing:
            raise LoggedError(None, "missing sampled parameters: %r", sorted(missing))
        out = dict(self._fixed)
        out.update(point)
        pending = dict(self._lambdas)
        while pending:
            progress = False
            for name, fn in list(pending.items()):
                args = list(inspect.signature(fn).parameters)
                if all(a in out for a in args):
                    out[name] = fn(*[out[a] for a in args])
                    del pending[name]
                    progress = True
            if not progress:
                raise LoggedError(None, "cannot resolve parameter definitions %r", sorted(pending))
        return out


class Model:
    def __init__(self, info, packages_path=None, stop_at_error=True):
        self.info = info
        self.parameterization = Parameterization(info.get("params"))
        self.theory, self.likelihood = {}, {}
        for name, cfg in (info.get("theory") or {}).items():
            cfg = dict(cfg or {})
            cfg.setdefault("stop_at_error", stop_at_error)
            comp = resolve_class(name, cfg)(info=cfg, name=name, packages_path=packages_path)
            self.theory[name] = comp
            for hname, helper in (comp.get_helper_theories() or {}).items():
                self.theory[hname] = helper
        for name, cfg in (info.get("likelihood") or {}).items():
            cfg = dict(cfg or {})
            cfg.setdefault("stop_at_error", stop_at_error)
            self.likelihood[name] = resolve_class(name, cfg)(info=cfg, name=name, packages_path=packages_path)
        self.components = list(self.theory.values()) + list(self.likelihood.values())
        self._resolve()

    # ------------------------------------------------------------------ dependency resolution
    def _find_provider(self, req, requester):
        found = [c for c in self.components if c is not requester and
                 (req in (c.get_can_provide() or []) or req in (c.get_can_provide_params() or [])
                  or callable(getattr(type(c), "get_" + req, None)))]
        if len(found) > 1:  # Cobaya: the component later in the list wins unless `provides` says otherwise
            found = found[-1:]
        return found[0] if found else None

    def _resolve(self):
        input_params = set(self.parameterization.input_params())
        self._deps = {c: set() for c in self.components}       # component -> components it needs
        self._direct_params = {c: set() for c in self.components}
        self.requirement_providers = {}
        pending = {c: {} for c in self.components}             # provider -> {requirement: options} not yet passed on
        queue = []
        for c in self.components:
            c.initialize_with_params()
            reqs = c.get_requirements() or {}
            if not isinstance(reqs, dict):
                reqs = {r: None for r in reqs}
            queue += [(c, r, o) for r, o in reqs.items()]
            for p in list(c._declared_params) + list(c.get_can_support_params() or []):
                if p in input_params:
                    self._direct_params[c].add(p)
        guard = 0
        while queue:
            guard += 1
            if guard > 10000:
                raise LoggedError(None, "requirement resolution does not terminate")
            requester, req, opts = queue.pop(0)
            if req in input_params:  # a requirement that is an input parameter becomes an input of the requester
                self._direct_params[requester].add(req)
                continue
            prov = self._find_provider(req, requester)
            if prov is None:
                raise LoggedError(None, "requirement %s of %s is not satisfied by any component or input parameter",
                                  req, requester.get_name())
            self.requirement_providers[req] = prov
            self._deps[requester].add(prov)
            more = prov.must_provide(**{req: opts if opts is not None else {}})
            if more:
                if not isinstance(more, dict):
                    more = {r: None for r in more}
                queue += [(prov, r, o) for r, o in more.items()]
        # output (derived) parameters
        for p in self.parameterization.derived_params():
            for c in self.components:
                if p in (c.get_can_provide_params() or []):
                    c.output_params.append(p)
                    self.requirement_providers.setdefault(p, c)
        # evaluation order (dependencies first) and inherited parameter dependencies
        order, seen = [], set()

        def visit(c, stack=()):
            if c in seen:
                return
            if c in stack:
                raise LoggedError(None, "circular dependency through %s", c.get_name())
            for d in self._deps[c]:
                visit(d, stack + (c,))
            seen.add(c)
            order.append(c)

        for c in self.components:
            visit(c)
        self._order = order
        self._all_params = {}
        for c in order:
            deps = set()
            for d in self._deps[c]:
                deps |= self._all_params[d] | self._direct_params[d]
            self._all_params[c] = deps
            c.input_params = sorted(self._direct_params[c])
        self.provider = Provider(self, self.requirement_providers)
        for c in self.components:
            c.initialize_with_provider(self.provider)

    # ------------------------------------------------------------------ evaluation
    def _compute(self, input_values, want_derived=True, cached=True):
        self.provider.set_current_input_params(input_values)
        derived = {}
        for c in self._order:
            own = {p: input_values[p] for p in c.input_params}
            dep = {p: input_values[p] for p in sorted(self._all_params[c] - set(c.input_params))}
            if c.check_cache_and_compute(own, dep, want_derived=want_derived, cached=cached) is False:
                return None
            if want_derived:
                derived.update({k: v for k, v in c.current_derived.items()})
        return derived

    def loglikes(self, point, return_derived=True, cached=True):
        """(array of log-likelihoods in `self.likelihood` order, derived dict) for a dict (or sequence) of sampled values"""
        values = self.parameterization.to_input(point)
        derived = self._compute(values, want_derived=return_derived, cached=cached)
        if derived is None and return_derived:
            return np.full(len(self.likelihood), -np.inf), {}
        ll = [lk.current_logp for lk in self.likelihood.values()]
        return (ll, derived) if return_derived else ll

    def loglike(self, point, return_derived=True, cached=True):
        out = self.loglikes(point, return_derived=return_derived, cached=cached)
        return (sum(out[0]), out[1]) if return_derived else sum(out)

    def logposterior(self, point, cached=True):
        ll, derived = self.loglikes(point, return_derived=True, cached=cached)
        return LogPosterior(logpost=sum(ll), logpriors=[0.0], loglikes=list(ll), derived=derived)

    def logpost(self, point, cached=True):
        return self.logposterior(point, cached=cached).logpost

    def close(self):
        for c in self.components:
            c.close()


def get_model(info, packages_path=None, stop_at_error=True, debug=False, **kwargs):
    return Model(info, packages_path=packages_path, stop_at_error=stop_at_error)
