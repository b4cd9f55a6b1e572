"""Cobaya components - `eftpipe.eftlss` / `eftpipe.eftlike` (eftpipe/__init__.py:2-4) on the B200 path.

    theory:
      eftpipe_b200.eftlss: {tracers: {...}}            # same block as `eftpipe.eftlss`
    likelihood:
      LEX_NGC: {class: eftpipe_b200.eftlike, ...}       # same block as `eftpipe.eftlike`

The component tree is the reference's (theory.py:116-886): `eftlss` fans the configuration out to one cosmology-
dependent `EFTLeafKernel` and one nuisance-dependent `EFTLeaf` helper theory per tracer, so that Cobaya's fast / slow
blocking works as before: a step that changes only bias parameters re-runs no pipeline kernel, only the reduction.
Requirement grammar, product getters and their signatures, derived parameters and error conventions follow the reference
(theory.py:165-267, :497-555, :773-874; likelihood.py:275-615).  Evaluation goes through the batched core
(`theory.EFTLSS`, `likelihood.EFTLike`) and the CUDA library; parameter values may be floats (Cobaya proper: B = 1, results
come back as numpy arrays / floats with the reference's shapes) or arrays of B points (population samplers, importance
re-weighting: results keep a leading batch axis and stay on the device).

This module needs the `cobaya` package; it is imported only on request (`eftpipe_b200.eftlss` resolves lazily).
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np

try:
    from cobaya.likelihood import Likelihood
    from cobaya.log import LoggedError as _CobayaLoggedError
    from cobaya.theory import HelperTheory, Theory
except ImportError as ex:  # pragma: no cover
    raise ImportError("eftpipe_b200.cobaya provides Cobaya components and needs the `cobaya` package "
                      "(the batched API in eftpipe_b200.theory / eftpipe_b200.likelihood works without it)") from ex

from . import likelihood as _like
from . import theory as _theory
from .boltzmann import find_boltzmann_extractor
from .marginal import LoggedError

_GRID_PRODUCTS = ("nonlinear_Plk_grid", "nonlinear_Plk_interpolator", "nonlinear_Plk_gaussian_grid")


def leaf_product_name(tracer):
    return f"eftleaf_{tracer}_results"


def leaf_kernel_product_name(tracer):
    return f"eftleaf_kernel_{tracer}_results"


def _is_batched(params):
    return any(np.ndim(v) > 0 for v in params.values())


def _to_host(x, squeeze):
    """device tensor (B, ...) -> numpy; the leading axis is dropped for a single point (reference shapes)"""
    a = x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)
    return a[0] if squeeze else a


def _recursive_update(base, other):
    for k, v in other.items():
        if isinstance(v, dict) and isinstance(base.get(k), dict):
            _recursive_update(base[k], v)
        else:
            base[k] = v


class eftlss(Theory):
    """Effective Field Theory of Large-scale Structures (theory.py:116-267)"""

    file_base_name = "eftlss"
    cache_dir_path = "cache"
    tracers = None

    def initialize(self):
        try:
            self.core = _theory.EFTLSS(deepcopy(self.tracers or {}), cache_dir_path=self.cache_dir_path)
        except LoggedError as ex:
            raise _CobayaLoggedError(self.log, "%s", str(ex))
        self.tracers = self.core.tracers
        self.tracer_names = list(self.tracers)
        self._built = False

    def get_requirements(self):
        """dummy requirements: make it possible to use eftlss with likelihood one (theory.py:159-164)"""
        return {leaf_product_name(t): {} for t in self.tracer_names}

    def must_provide(self, **requirements):
        """redirect requirements to EFTLeaf and EFTLeafKernel (theory.py:165-194); the batched core records them too"""
        super().must_provide(**requirements)
        redirected = {}
        for product, per_tracer in requirements.items():
            if product.startswith("eftleaf_"):
                continue
            for tracer, config in (per_tracer or {}).items():
                if tracer != "default" and tracer not in self.tracer_names:
                    raise _CobayaLoggedError(self.log, "Unknown tracer name: %s", tracer)
                redirected.setdefault(tracer, {})[product] = config
        try:
            self.core.must_provide({k: v for k, v in requirements.items() if not k.startswith("eftleaf_")})
        except LoggedError as ex:
            raise _CobayaLoggedError(self.log, "%s", str(ex))
        reqs = {}
        for tracer, product_config in redirected.items():
            reqs[leaf_product_name(tracer)] = deepcopy(product_config)
            reqs[leaf_kernel_product_name(tracer)] = deepcopy(product_config)
        if default := reqs.pop(leaf_product_name("default"), None):
            reqs.pop(leaf_kernel_product_name("default"), None)
            for name, config in reqs.items():
                merged = deepcopy(default)
                _recursive_update(merged, config)
                reqs[name] = merged
        return reqs

    def get_helper_theories(self):
        info = {"stop_at_error": getattr(self, "stop_at_error", False)}
        helpers = {}
        for i, tracer in enumerate(self.tracer_names):
            zextra = []
            if i == 0 and len(self.tracer_names) < 4:  # Pk_interpolator requires at least 4 redshifts (theory.py:199-203)
                zeff = self.tracers[tracer]["z"]
                zextra = [zeff + j * 0.1 for j in range(1, 5 - len(helpers))]
            helpers[self.leaf_name(tracer)] = EFTLeaf(info=info, name=self.leaf_name(tracer), timing=getattr(self, "timer", None),
                                                      tracer=tracer, eftlss=self)
            helpers[self.leaf_kernel_name(tracer)] = EFTLeafKernel(info=info, name=self.leaf_kernel_name(tracer),
                                                                   timing=getattr(self, "timer", None), tracer=tracer, eftlss=self,
                                                                   zextra=list(zextra))
        return helpers

    def leaf_name(self, tracer):
        return self.get_name() + "." + tracer

    def leaf_kernel_name(self, tracer):
        return self.leaf_name(tracer) + ".kernel"

    def build(self):
        """device plans of every tracer, once all requirements are in (theory.py:399-495 does this per kernel)"""
        if not self._built:
            try:
                self.core.initialize()
            except LoggedError as ex:
                raise _CobayaLoggedError(self.log, "%s", str(ex))
            self._built = True

    def initialize_with_provider(self, provider):
        super().initialize_with_provider(provider)
        self.build()

    # ---- product getters, reference signatures (theory.py:229-267) ----
    def retrieve_product_from_leaf(self, tracer, product):
        if tracer not in self.tracer_names:
            raise ValueError(f"Tracer {tracer} not in {self.tracer_names}!")
        try:
            return self.provider.get_result(leaf_product_name(tracer))[product]
        except KeyError:
            raise _CobayaLoggedError(self.log, "%s not computed, please check if you have specified it in requirements", product)

    def get_nonlinear_Plk_grid(self, tracer, chained=False, binned=False):
        return self.retrieve_product_from_leaf(tracer, ("nonlinear_Plk_grid", chained, binned))

    def get_nonlinear_Plk_gaussian_grid(self, tracer, chained=False, binned=False):
        return self.retrieve_product_from_leaf(tracer, ("nonlinear_Plk_gaussian_grid", chained, binned))

    def get_nonlinear_Plk_interpolator(self, tracer, chained=False):
        return self.retrieve_product_from_leaf(tracer, ("nonlinear_Plk_interpolator", chained))

    def get_snapshots(self, tracer):
        return self.retrieve_product_from_leaf(tracer, "snapshots")

    def get_eft_params_values_dict(self, tracer):
        return self.retrieve_product_from_leaf(tracer, "eft_params_values_dict")

    def get_bird_component(self, tracer):
        return self.retrieve_product_from_leaf(tracer, "bird_component")


class _LeafShared:
    def _setup(self, tracer, eftlss_):
        self.tracer, self.eftlss = tracer, eftlss_
        self.tracer_config = eftlss_.tracers[tracer]
        self.tracer_prefix = _theory.tracer_prefix(tracer, self.tracer_config)

    def cross_type(self):
        return self.tracer_config.get("cross", False)


class EFTLeafKernel(HelperTheory, _LeafShared):
    """EFT theory for a single tracer, the part that depends on the cosmology only (theory.py:297-721): pulls the linear
    power through the `BoltzmannExtractor` seam and runs the fused device pipeline of the tracer"""

    def __init__(self, info=None, name=None, timing=None, packages_path=None, initialize=True, standalone=True, tracer="",
                 eftlss=None, zextra=()):
        self._setup(tracer, eftlss)
        self.zextra = list(zextra or [])
        super().__init__(info=info or {}, name=name, timing=timing, packages_path=packages_path, initialize=initialize,
                         standalone=standalone)

    def initialize(self):
        super().initialize()
        self.basis = self.eftlss.core.build_basis(self.tracer)
        self.zeff = self.tracer_config["z"]
        self.boltzmann = find_boltzmann_extractor(self.tracer_config.get("provider", "classy"),
                                                  self.tracer_config.get("provider_kwargs", {}))
        self.boltzmann.initialize(zeff=self.zeff, use_cb=self.tracer_config.get("use_cb", False), zextra=self.zextra)
        self.with_APeffect = bool(self.tracer_config.get("with_APeffect", False))
        self._epoch = 0

    def initialize_with_provider(self, provider):
        super().initialize_with_provider(provider)
        self.boltzmann.initialize_with_provider(provider)
        self.eftlss.build()

    def get_requirements(self):
        return self.boltzmann.get_requirements()

    def required_power_spectrum(self):
        return self.tracer in self.eftlss.core.plans

    def calculate(self, state, want_derived=True, **params_values_dict):
        core = self.eftlss.core
        state[self.product_name()] = {}
        boltzmann = self.boltzmann
        if self.required_power_spectrum():
            boltzmann.calculate(**params_values_dict)
            kh = np.logspace(-5, 0, 200)  # theory.py:562
            pkh = boltzmann.Pkh(kh)
            if hasattr(pkh, "detach"):  # a device producer (boltzmann.EisensteinHu): tensors go straight to the pipeline
                B = pkh.shape[0]
                cosmo = {k: v for k, v in dict(pkh=pkh, f=boltzmann.f(), DA=boltzmann.DA(), H=boltzmann.H(), rdrag=boltzmann.rdrag(),
                                               h=boltzmann.h()).items() if v is not None}
            else:
                pkh = np.atleast_2d(np.asarray(pkh, float))
                B = pkh.shape[0]
                vec = lambda v: None if v is None else np.broadcast_to(np.asarray(_np_or_tensor(v), float).reshape(-1), (B,)).copy()
                cosmo = dict(pkh=pkh, f=vec(boltzmann.f()), DA=vec(boltzmann.DA()), H=vec(boltzmann.H()), rdrag=vec(boltzmann.rdrag()),
                             h=vec(boltzmann.h()))
            fs8 = boltzmann.fsigma8_z()
            if not (np.isscalar(fs8) and fs8 == -1):
                cosmo["fsigma8_z"] = fs8
            core.calculate({self.tracer: cosmo}, reset=False)
            self._epoch += 1
            state[self.product_name()] = {"tracer": self.tracer, "epoch": self._epoch, "B": B}
            if want_derived:
                core._derived = None
                d = {k: v for k, v in core.derived.items() if k.startswith(self.tracer_prefix)}
                squeeze = B == 1
                for k, v in d.items():
                    state["derived"][k] = (float(np.asarray(v).reshape(-1)[0]) if squeeze and np.ndim(v) else v)
        elif want_derived:  # no power spectrum requested: only the Boltzmann-side derived parameters (theory.py:616-637)
            state["derived"][self.tracer_prefix + "alperp"] = state["derived"][self.tracer_prefix + "alpara"] = -1
            state["derived"][self.tracer_prefix + "fz"] = boltzmann.f()
            state["derived"][self.tracer_prefix + "fsigma8_z"] = boltzmann.fsigma8_z()

    def get_can_provide(self):
        return [self.product_name()]

    def get_can_provide_params(self):
        return [self.tracer_prefix + item for item in ("fz", "fsigma8_z", "fsigma8_cb_z", "alperp", "alpara")]

    def product_name(self):
        return leaf_kernel_product_name(self.tracer)


def _np_or_tensor(v):
    return v.detach().cpu().numpy() if hasattr(v, "detach") else v


class _LeafProducts(dict):
    """products of one EFTLeaf evaluation, reduced on the device the first time they are read"""

    def __init__(self, leaf, params, squeeze):
        super().__init__()
        self.leaf, self.params, self.squeeze = leaf, params, squeeze

    def __missing__(self, key):
        leaf, core, t = self.leaf, self.leaf.eftlss.core, self.leaf.tracer
        if key == "eft_params_values_dict":
            val = {p: self.params.get(p, 0.0) for p in leaf.basis.gaussian_params() + leaf.basis.non_gaussian_params()}
        elif key == "snapshots":
            val = core.get_snapshots(t)
        elif key == "bird_component":
            if "bird_component" not in leaf._must_provide:
                raise KeyError(key)
            val = core.get_bird_component(t, self.params, chained=False, binned=False)
        elif isinstance(key, tuple) and key[0] in _GRID_PRODUCTS and key[0] in leaf._must_provide and \
                tuple(key[1:]) + ((False,) if key[0] == "nonlinear_Plk_interpolator" else ()) in leaf._must_provide[key[0]]:
            try:
                if key[0] == "nonlinear_Plk_grid":
                    ls, k, plk = core.get_nonlinear_Plk_grid(t, self.params, chained=key[1], binned=key[2])
                    val = (ls, k, _to_host(plk, self.squeeze))
                elif key[0] == "nonlinear_Plk_gaussian_grid":
                    ls, k, tab = core.get_nonlinear_Plk_gaussian_grid(t, self.params, chained=key[1], binned=key[2])
                    val = (ls, k, {n: _to_host(v, self.squeeze) for n, v in tab.items()})
                else:
                    fn = core.get_nonlinear_Plk_interpolator(t, self.params, chained=key[1])
                    fn.Plk = _to_host(fn.Plk, self.squeeze)
                    val = fn
            except LoggedError:
                raise KeyError(key)
        else:
            raise KeyError(key)
        self[key] = val
        return val


class EFTLeaf(HelperTheory, _LeafShared):
    """EFT theory for a single tracer, the part that depends on the EFT parameters (theory.py:723-886)"""

    def __init__(self, info=None, name=None, timing=None, packages_path=None, initialize=True, standalone=True, tracer="", eftlss=None):
        self._setup(tracer, eftlss)
        super().__init__(info=info or {}, name=name, timing=timing, packages_path=packages_path, initialize=initialize,
                         standalone=standalone)

    def initialize(self):
        super().initialize()
        self.basis = self.eftlss.core.build_basis(self.tracer)
        self._must_provide = {p: set() for p in _GRID_PRODUCTS}

    def get_requirements(self):
        requires = {k: None for k in self.basis.non_gaussian_params()}
        requires[self.kernel_product_name()] = {}
        return requires

    def must_provide(self, **requirements):
        """theory.py:773-827"""
        super().must_provide(**requirements)
        for product, config in (requirements.get(self.product_name()) or {}).items():
            if product in _GRID_PRODUCTS:
                chained = _theory._bool_list((config or {}).get("chained", False))
                binned = _theory._bool_list((config or {}).get("binned", False))
                if product == "nonlinear_Plk_interpolator" and True in binned:
                    raise _CobayaLoggedError(self.log, "binned Plk interpolator not supported")
                for c in chained:
                    for b in binned:
                        self._must_provide[product].add((c, b))
            elif product in ("snapshots", "eft_params_values_dict", "bird_component"):
                self._must_provide[product] = set()
            else:
                raise _CobayaLoggedError(self.log, "Unexpected requirement %s, this should not happen, please contact the developers", product)

    def calculate(self, state, want_derived=True, **params_values_dict):
        kernel_product = self.provider.get_result(self.kernel_product_name())
        if kernel_product:
            squeeze = kernel_product["B"] == 1 and not _is_batched(params_values_dict)
            state[self.product_name()] = _LeafProducts(self, dict(params_values_dict), squeeze)
        else:
            state[self.product_name()] = {}

    def get_can_provide(self):
        return [self.product_name()]

    def get_can_support_params(self):
        return self.basis.gaussian_params()

    def product_name(self):
        return leaf_product_name(self.tracer)

    def kernel_product_name(self):
        return leaf_kernel_product_name(self.tracer)


class eftlike(Likelihood):
    """EFT likelihood for an arbitrary number of tracers and cross-correlations (likelihood.py:275-615); yaml keys as in
    eftlike.yaml.  PNG / PG assembly, chi^2 and the analytic marginalisation run in the CUDA likelihood kernels on the
    terms the `eftlss` core left on the device."""

    file_base_name = "eftlike"
    likelihood_prefix = None
    marg_param_prefix = "marg_"
    tracers = None
    data = None
    cov = None
    chained = False
    with_interp = True
    with_binning = False
    binning = None
    marg = None
    jeffreys = False

    def initialize(self):
        super().initialize()
        if self.likelihood_prefix is None:
            self.likelihood_prefix = self.get_name() + "_"
        data = deepcopy(self.data)
        tracers = [self.tracers] if isinstance(self.tracers, str) else list(self.tracers)
        if len(tracers) == 1 and isinstance(data, dict) and next(iter(data)) != tracers[0]:
            data = {tracers[0]: data}
        if isinstance(data, list):
            data = {t: data[i] for i, t in enumerate(tracers)}
        self.core = _like.EFTLike(tracers=tracers, data=data, cov=deepcopy(self.cov), chained=self.chained,
                                  with_binning=self.with_binning, binning=deepcopy(self.binning), marg=deepcopy(self.marg),
                                  jeffreys=self.jeffreys, likelihood_prefix=self.likelihood_prefix,
                                  marg_param_prefix=self.marg_param_prefix, with_interp=self.with_interp)
        self.tracers = tracers
        self.data_vector, self.invcov, self.ndata = self.core.data_vector, self.core.invcov, self.core.ndata
        self.minfodict = self.core.minfodict

    def get_requirements(self):
        """likelihood.py:386-432 (+ `kout` on the interpolator requirement, see theory.EFTLSS.must_provide)"""
        reqs = self.core.get_requirements()
        reqs["eft_params_values_dict"] = {t: None for t in self.tracers}
        return reqs

    def initialize_with_provider(self, provider):
        super().initialize_with_provider(provider)
        comps = [c for c in provider.model.theory.values() if isinstance(c, eftlss)]
        if not comps:
            raise _CobayaLoggedError(self.log, "eftlike needs the eftpipe_b200.eftlss theory")
        self.eftlss = comps[0]
        self.eftlss.build()
        try:
            self.core.initialize_with_provider(self.eftlss.core)
        except LoggedError as ex:
            raise _CobayaLoggedError(self.log, "%s", str(ex))
        self.eft_bases = self.core.eft_bases

    # Marginalizable interface (marginal.py:42-58), for consumers that introspect the likelihood
    def marginalizable_params(self):
        return self.core.marginalizable_params()

    def get_data_vector(self):
        return self.data_vector

    def get_invcov(self):
        return self.invcov

    def _params(self):
        params = {}
        for t in self.tracers:
            params.update(self.provider.get_eft_params_values_dict(t))
        return params

    def PNG(self):
        png, _ = self.core.PNG_PG(self._params())
        return _to_host(png, png.shape[0] == 1)

    def PG(self):
        _, pg = self.core.PNG_PG(self._params())
        return _to_host(pg, pg.shape[0] == 1)

    def required_bGbest_related_derived_params(self):
        return any(p.startswith(self.marg_param_prefix) for p in self.output_params)

    def calculate(self, state, want_derived=True, **params_values_dict):
        """likelihood.py:570-594"""
        params = self._params()
        want_best = self.required_bGbest_related_derived_params()
        res = self.core.calculate(params, want_bestfit=want_best)
        squeeze = self.eftlss.core.B == 1 and not _is_batched(params)
        host = lambda x: (float(_to_host(x, True)) if squeeze else x)
        status = _to_host(res["status"], False)
        if squeeze and status[0] != 0:  # marginal.py:113-116 raises for a single point; a batch flags the point instead
            raise RuntimeError("det of F2ij <= 0")
        state["logp"] = host(res["logp"])
        if want_derived:
            state["derived"][self.likelihood_prefix + "chi2"] = host(res[self.likelihood_prefix + "chi2"])
            full = res.get(self.likelihood_prefix + "fullchi2")
            state["derived"][self.likelihood_prefix + "fullchi2"] = host(full) if full is not None else state["derived"][self.likelihood_prefix + "chi2"]
            if want_best:
                for p in self.output_params:
                    if p.startswith(self.marg_param_prefix):
                        v = res["bestfit"].get(p)
                        state["derived"][p] = host(v) if v is not None else 0.0
        if not squeeze:
            state["status"] = res["status"]

    def get_can_provide_params(self):
        """likelihood.py:596-612: chi2, fullchi2 and the best-fit value of every marginalised parameter"""
        out = [self.likelihood_prefix + p for p in ("chi2", "fullchi2")]
        return out + [self.marg_param_prefix + n for n in self._estimated_marg_names()]

    def _estimated_marg_names(self):
        names = []
        for p, config in (self.marg or {}).items():
            if isinstance(config, dict) and not _like.valid_prior_config(config):
                names += [f"{p}{n}" for n in config]
            else:
                names.append(p)
        return names
