"""Batched theory driver - the numeric content of `eftpipe.theory.EFTLSS / EFTLeafKernel / EFTLeaf`
(theory.py:116-886) without the Cobaya plumbing: same per-tracer configuration keys (theory.py:356-388,
eftlss.yaml), same `default` block deep-merge (theory.py:133-138), same stage order
(`calculate_power_spectrum`, theory.py:557-609), same product getters (theory.py:244-267), for a batch of
points per call.  Cobaya's `Theory` protocol cannot be exercised in this image (no cobaya); INTEGRATION.md
shows the thin adapter a maintainer adds on top of this class.
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np

from . import plan as P
from .binning import Binning
from .engine import DevicePlan
from .icc import IntegralConstraint
from .marginal import LoggedError
from .parambasis import find_param_basis
from .plugins import find_window_constructor, probe_linear_stage
from .pybird import APeffect, Common, DAfunc, FiberCollision, Hubble
from .window import Window

_KNOWN = {"prefix", "z", "nd", "km", "kr", "cross", "provider", "provider_kwargs", "use_cb", "with_IRresum", "with_APeffect",
          "with_window", "with_icc", "with_fiber", "with_NNLO", "with_RSD", "kmax", "IRresum", "APeffect", "window", "icc",
          "fiber", "binning", "basis", "counterform", "chained", "ls", "Nl", "optiresum", "IRcutoff", "kIR"}


def tracer_prefix(name, cfg):
    """theory.py:285-291 (`LeafKernelShared.set_tracer_prefix` / `build_basis`): an absent `prefix` means `<tracer>_`;
    an explicitly empty one is kept for the tracer itself"""
    prefix = cfg.get("prefix")
    return name + "_" if prefix is None else prefix


def related_prefix(name, cfg):
    """theory.py:290-291: the parents of a cross tracer fall back to `<tracer>_` also for an EMPTY prefix"""
    return cfg["prefix"] if cfg.get("prefix") else name + "_"


def _construct(cls, cfg, **fixed):
    """`tools.Initializer` (tools.py:176-205): validate a plugin's yaml sub-dict against the constructor signature before
    building it - unknown keyword / missing positional argument -> LoggedError, as in the reference."""
    import inspect

    params = inspect.signature(cls).parameters
    has_var_kw = any(p.kind is p.VAR_KEYWORD for p in params.values())
    for k in cfg:
        if k not in params and not has_var_kw:
            raise LoggedError(f"{cls!r} does not have keyword {k}")
    for k, p in params.items():
        if p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL) or k == "self":
            continue
        if p.default is p.empty and k not in cfg and k not in fixed:
            raise LoggedError(f"missing positional argument {k}")
    return cls(**{**cfg, **fixed})


def _merge(default, cfg):
    out = deepcopy(default)
    for k, v in cfg.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = deepcopy(v)
    return out


class EFTLSS:
    def __init__(self, tracers: dict, cache_dir_path=None):
        tracers = deepcopy(tracers)
        default = tracers.pop("default", {})
        self.tracers = {name: _merge(default, cfg or {}) for name, cfg in tracers.items()}
        for name, cfg in self.tracers.items():
            unknown = set(cfg) - _KNOWN
            if unknown:
                raise LoggedError(f"tracer {name}: unknown configuration keys {sorted(unknown)}")
        self.cache_dir_path = cache_dir_path
        self.requirements = {}
        self.bases, self.commons, self.plans, self.info = {}, {}, {}, {}
        self._state = {}
        self.B = 0

    # ---- requirements (theory.py:773-799) ----
    def must_provide(self, requirements: dict):
        """`nonlinear_Plk_interpolator` (theory.py:785-790) takes one extra key here, `kout`: the abscissae the
        consumer will evaluate the interpolator at.  The reference hands out a callable; on the batched path the
        interpolation is a fixed operator composed into the tracer's projection (plan.interp_matrices), so the theory
        has to know the points when the plan is built."""
        for product, per_tracer in requirements.items():
            if product not in ("nonlinear_Plk_grid", "nonlinear_Plk_gaussian_grid", "nonlinear_Plk_interpolator"):
                raise LoggedError(f"unsupported requirement {product} on the batched path")
            for tracer, req in per_tracer.items():
                if tracer not in self.tracers:
                    raise LoggedError(f"unknown tracer {tracer}")
                old = self.requirements.get(tracer)
                new = dict(ls=sorted(req["ls"]), chained=bool(req.get("chained", False)),
                           binned=bool(req.get("binned", False)), binning=req.get("binning"), interp=None)
                if product == "nonlinear_Plk_interpolator":
                    if "kout" not in req:
                        raise LoggedError("nonlinear_Plk_interpolator: the batched path needs the evaluation points `kout`")
                    new["binned"], new["interp"] = False, np.asarray(req["kout"], float)
                if old is not None and (old["ls"], old["chained"], old["binned"]) != (new["ls"], new["chained"], new["binned"]):
                    raise LoggedError("does not support multiple different product requirements per tracer")
                if old is not None and new["interp"] is None:
                    new["interp"] = old["interp"]
                self.requirements[tracer] = new
        return self

    # ---- plan construction (theory.py:399-495) ----
    def _scales(self, name):
        """kmA, krA, ndA, kmB, krB, ndB (theory.py:663-699)."""
        cfg = self.tracers[name]
        cross = cfg.get("cross")
        def own(c):
            try:
                km, nd = c["km"], c["nd"]
            except KeyError:
                raise LoggedError("must specify km, kr and nd")
            return km, c.get("kr") or km, nd
        if isinstance(cross, (list, tuple)):
            a, b = (self.tracers[t] for t in cross)
            return own(a) + own(b)
        return own(cfg) + own(cfg)

    def initialize(self):
        for name, cfg in self.tracers.items():
            req = self.requirements.get(name)
            if req is None:
                continue
            ls = req["ls"]
            No = max(ls) // 2 + 1 + (1 if req["chained"] else 0)
            Nl = max(cfg.get("Nl", No), No)
            kmA, krA, ndA, kmB, krB, ndB = self._scales(name)
            basis_cls = find_param_basis(cfg.get("basis", "westcoast"))
            cross = cfg.get("cross")
            cross_prefix = [related_prefix(t, self.tracers[t]) for t in cross] if isinstance(cross, (list, tuple)) else []
            basis = basis_cls(prefix=tracer_prefix(name, cfg), cross_prefix=cross_prefix)
            co = Common(Nl=Nl, No=No, kmax=cfg.get("kmax", 0.3), kmA=kmA, krA=krA, ndA=ndA, kmB=kmB, krB=krB, ndB=ndB,
                        counterform=basis.counterform(), with_NNLO=bool(cfg.get("with_NNLO", False)),
                        optiresum=bool(cfg.get("optiresum", False)), IRcutoff=cfg.get("IRcutoff", False),
                        kIR=cfg.get("kIR"))  # theory.py:421-437
            ap = None
            if cfg.get("with_APeffect"):
                apc = dict(cfg.get("APeffect") or {})
                apc.setdefault("z_AP", cfg["z"])  # theory.py:458: z_AP defaults to the tracer's z
                apo = _construct(APeffect, apc, co=Common(Nl=Nl))
                ap = dict(DA=apo.DA, H=apo.H, nbinsmu=apc.get("nbinsmu", 200), accboost=apc.get("accboost", 1), APst=apo.APst)
                self.info.setdefault(name, {})["ap"] = apo
            window = icc = custom = None
            ww = cfg.get("with_window")
            if ww:
                wc = dict(cfg.get("window") or {})
                if cfg.get("with_icc", False):  # theory.py:384-388, :463-471
                    if cross:
                        raise LoggedError("integral constraint correction (icc) not yet supported for cross power spectrum")
                    icc = _construct(IntegralConstraint, dict(cfg.get("icc") or {}), co=co)
                if isinstance(ww, str) and ww not in ("auto", "default"):
                    # theory.py:62-72, :370-377: a class by dotted path with an in-place `.Window(bird)`; probed once
                    # into a fixed operator (plugins.probe_linear_stage) - ICC and the Picc constant included
                    plugin = _construct(find_window_constructor(ww), wc, co=co, icc=icc, name=f"{name}.window")
                    custom = probe_linear_stage(plugin.Window, co)
                else:
                    window = _construct(Window, wc, co=co, icc=icc)
            fiber = None
            if cfg.get("with_fiber"):  # theory.py:378-385, :482-485
                fiber = _construct(FiberCollision, dict(cfg.get("fiber") or {}), co=co)
            binm = None
            keff = co.k
            if req["binned"]:
                bo = Binning(co=co, **(req["binning"] or cfg.get("binning") or {}))
                binm, keff = bo.matrix, bo.keff
            elif req["interp"] is not None:
                # un-binned interpolated products: rows [0, nkout) = PlkInterpolator (theory.py:75-106), rows
                # [nkout, 2 nkout) = the plain cubic interpolation of the marginalised rows (likelihood.py:510-513)
                keff = req["interp"]
                binm = np.vstack(P.interp_matrices(co.k, keff))
            g = P.GridConfig(Nl=Nl, kmax=cfg.get("kmax", 0.3), with_NNLO=co.with_NNLO, optiresum=co.optiresum)
            proj = None
            if custom is not None:
                proj = P.compose_projection(g, window=custom["matrix"], icc=None, binning=binm, chained=req["chained"],
                                            fiber=None if fiber is None else fiber.matrix(),
                                            fiber_st=False if fiber is None else fiber.fiberst,
                                            window_stoch=custom["matrix_st"], window_picc=custom["picc"])
                proj["kout"] = keff
            elif window is not None or binm is not None or req["chained"] or fiber is not None:
                proj = P.compose_projection(
                    g, window=None if window is None else window_matrix(window),
                    icc=None if icc is None else dict(matrix=icc.effective_matrix(), PSN_times_Pshot=icc.PSN),
                    binning=binm, chained=req["chained"], window_st=True if window is None else window.window_st,
                    fiber=None if fiber is None else fiber.matrix(), fiber_st=False if fiber is None else fiber.fiberst)
                proj["kout"] = keff
            rs = cfg.get("IRresum") or {}
            host = P.build_tracer_plan(Nl=Nl, kmax=cfg.get("kmax", 0.3), with_NNLO=co.with_NNLO,
                                       with_resum=bool(cfg.get("with_IRresum", True)), resum_NFFT=rs.get("NFFT", 192),
                                       ap=ap, projection=proj, optiresum=co.optiresum, ircutoff=co.IRcutoff, kIR=co.kIR,
                                       lambda_ir=rs.get("LambdaIR", P.LAMBDA_IR))
            self.bases[name], self.commons[name] = basis, co
            self.plans[name] = DevicePlan(host)
            nl_out, nk = host.out_shape if proj is not None else (Nl, g.Nk)
            picc = host.picc_out if proj is not None else np.zeros(Nl * g.Nk)
            self.info.setdefault(name, {}).update(nout=nl_out * nk, nterm=g.nterm, nk=nk, picc=picc, kout=keff,
                                                  ls=[2 * i for i in range(nl_out)], No=No,
                                                  interp_nk=None if req["interp"] is None else len(keff))
        return self

    def product_info(self, tracer, chained=False, binned=True):
        return self.info[tracer]

    # ---- per batch (theory.py:557-609) ----
    def calculate(self, cosmo: dict):
        """cosmo[tracer] = dict(pkh=(B, 200) on kh = logspace(-5, 0, 200), f=, DA=, H= (B,) [, rdrag, h])."""
        self._state, self._derive_inputs, self._derived = {}, {}, None
        import os

        import torch

        # tracers are independent pipelines: each runs on its own stream, forked from and joined back into the caller's
        # stream (capturable); the tails and latency-bound kernels of one tracer fill with the work of the others
        concurrent = len(self.plans) > 1 and os.environ.get("EFTB_TRACER_STREAMS", "1") != "0"
        main = torch.cuda.current_stream()
        if concurrent:
            if getattr(self, "_streams", None) is None:
                self._streams = {name: torch.cuda.Stream() for name in self.plans}
            fork = torch.cuda.Event()
            fork.record(main)
        for name, dp in self.plans.items():
          with torch.cuda.stream(self._streams[name]) if concurrent else _nullcontext():
            if concurrent:
                self._streams[name].wait_event(fork)
            c = cosmo[name]
            if hasattr(c, "Pkh"):  # a boltzmann.BoltzmannExtractor (theory.py:559-565)
                c = c.cosmo()
            if not self.tracers[name].get("with_RSD", True):  # theory.py:566-567
                c = dict(c, f=c["f"] * 0.0)  # same container type and device as the input
            pm, bm = dp.eval_terms(c["pkh"], c["f"], c.get("DA"), c.get("H"), want_bm=True, want_pm=False)
            f_bm = dp.to_batch_minor(c["f"])[0]
            self._state[name] = (bm, f_bm)
            self.B = bm.shape[-1] if not hasattr(c["pkh"], "shape") else c["pkh"].shape[0]
            self._derive_inputs[name] = c
            if concurrent:
                done = torch.cuda.Event()
                done.record(self._streams[name])
                main.wait_event(done)
                for ten in (bm, f_bm):
                    ten.record_stream(main)  # allocated on the tracer's stream, consumed on the caller's
        return self

    @property
    def derived(self):
        """derived parameters of theory.py:620-648, `{prefix}alperp, alpara, fz, fsigma8_z` per point - evaluated on first
        access (they read the inputs back to the host, which must not happen while `calculate` is being captured into a
        CUDA graph)"""
        if self._derived is None:
            self._derived = {}
            for name, c in self._derive_inputs.items():
                self._derive(name, c)
        return self._derived

    def _derive(self, name, c):
        prefix = tracer_prefix(name, self.tracers[name])
        apo = self.info.get(name, {}).get("ap")
        if apo is not None and c.get("DA") is not None and c.get("H") is not None:
            qperp, qpar = np.asarray(_host(c["DA"]), float) / apo.DA, apo.H / np.asarray(_host(c["H"]), float)  # pybird.py:1560-1562
            if all(x is not None for x in (apo.rdrag_AP, apo.h_AP, c.get("rdrag"), c.get("h"))):
                ratio = (apo.rdrag_AP * apo.h_AP) / (np.asarray(_host(c["rdrag"]), float) * np.asarray(_host(c["h"]), float))
                qperp, qpar = qperp * ratio, qpar * ratio  # pybird.py:1576-1578
            self._derived[prefix + "alperp"], self._derived[prefix + "alpara"] = qperp, qpar
        else:
            self._derived[prefix + "alperp"] = self._derived[prefix + "alpara"] = -1
        self._derived[prefix + "fz"] = np.asarray(_host(c["f"]), float)
        if c.get("fsigma8_z") is not None:
            self._derived[prefix + "fsigma8_z"] = np.asarray(_host(c["fsigma8_z"]), float)

    def get_bird_component(self, tracer, params, chained=False, binned=True):
        """(ls, k, BirdComponent) - theory.py:265-266, :844-847"""
        info = self.info[tracer]
        return info["ls"], info["kout"], self.bases[tracer].reduce_Plk(self._view(tracer), params)

    def get_eft_params_values_dict(self, tracer, params):
        """the tracer's EFT parameters, absent ones as 0.0 (theory.py:262-263, :839-843)"""
        basis = self.bases[tracer]
        names = list(basis.gaussian_params()) + list(basis.non_gaussian_params())
        return {n: params.get(n, 0.0) for n in names}

    def get_snapshots(self, tracer):
        raise LoggedError("snapshots are taken on the stage-by-stage path (pybird.Bird.create_snapshot); the fused "
                          "batched pipeline keeps no intermediate term arrays")

    def get_nonlinear_Plk_terms(self, tracer, chained=False, binned=True):
        return self._state[tracer]

    def get_nonlinear_Plk_grid(self, tracer, params, chained=False, binned=True):
        """(ls, k, Plk (B, No, nk)) - theory.py:244-252 + EFTLeaf reduction (theory.py:846-860)."""
        bird = self._view(tracer)
        comp = self.bases[tracer].reduce_Plk(bird, params)
        info = self.info[tracer]
        plk = comp.sum()
        if info["interp_nk"] is not None:  # interpolated products: the PlkInterpolator rows
            plk = plk[..., : info["interp_nk"]]
        return info["ls"], info["kout"], plk

    def get_nonlinear_Plk_gaussian_grid(self, tracer, params, chained=False, binned=True):
        bird = self._view(tracer)
        info = self.info[tracer]
        table = self.bases[tracer].reduce_Plk_gaussian_table(bird, params)
        if info["interp_nk"] is not None:  # the rows interpolated without the inserted origin
            table = {name: v[..., info["interp_nk"]:] for name, v in table.items()}
        return info["ls"], info["kout"], table

    def get_nonlinear_Plk_interpolator(self, tracer, params, chained=False):
        """theory.py:254-258: a `PlkInterpolator` over the tracer's un-binned grid (request `nonlinear_Plk_grid` with
        `binned: False`).  A tracer whose plan was built for fixed evaluation points (`nonlinear_Plk_interpolator`
        requirement with `kout`) already holds the interpolated values: use `get_nonlinear_Plk_grid` there."""
        info = self.info[tracer]
        if info["interp_nk"] is not None or self.requirements[tracer]["binned"]:
            raise LoggedError("get_nonlinear_Plk_interpolator needs the un-binned grid of the tracer")
        ls, k, plk = self.get_nonlinear_Plk_grid(tracer, params, chained=chained, binned=False)
        return PlkInterpolator(ls[: plk.shape[-2]], k, plk)

    def _view(self, tracer):
        from .transformer import PlainBird

        bm, f_bm = self._state[tracer]
        info = self.info[tracer]
        nl_out = len(info["ls"])
        T = bm.reshape(nl_out, info["nk"], info["nterm"], bm.shape[-1])
        co = self.commons[tracer]
        view = PlainBird(None, co, T, np.asarray(info["picc"]).reshape(nl_out, info["nk"]), self.B, False, f_bm)
        # the reduction honours co.No: chained products expose one multipole fewer (theory.py:599-602)
        return view


class PlkInterpolator:
    """theory.py:75-106 with a leading batch axis: cubic interpolation of k P_l(k) through the grid plus an inserted
    (0, 0) point, extrapolating.  `Plk`: (..., len(ls), nk) torch tensor or array; `fn(l, k)` returns (..., nk') for
    one multipole or (..., len(l), nk') for a list, like the reference.  The interpolation is the fixed operator
    `plan.interp_matrices`; it is applied with one small matrix product (a convenience product, not the likelihood
    path, which composes the same operator into the projection GEMM)."""

    def __init__(self, ls, kgrid, Plk):
        self.ls = list(ls)
        self.kgrid = np.asarray(kgrid, float)
        self.Plk = Plk

    def fn(self, k):
        S = P.interp_matrices(self.kgrid, np.atleast_1d(np.asarray(k, float)))[0]
        if isinstance(self.Plk, np.ndarray):
            return self.Plk @ S.T
        import torch

        return self.Plk @ torch.as_tensor(S.T, dtype=self.Plk.dtype, device=self.Plk.device)

    def __call__(self, l, k):
        l = [l] if isinstance(l, (int, np.integer)) else list(l)
        try:
            idx = [self.ls.index(ll) for ll in l]
        except ValueError as ex:
            raise ValueError(f"l={l} not in {self.ls}") from ex
        out = self.fn(k)
        return out[..., idx[0], :] if len(idx) == 1 else out[..., idx, :]


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def _host(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else x


def window_matrix(window: Window):
    """effective matrix WITHOUT the ICC part (compose_projection subtracts it)."""
    return P.window_effective_matrix(window.Wal, window.p, window.co.k, windowk=window.windowk, withmask=window.withmask)
