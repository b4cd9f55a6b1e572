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


def _bool_list(x):
    """tools.bool_or_list: a bool or a list of bools -> ordered list without repetitions"""
    vals = [bool(x)] if isinstance(x, (bool, int, np.bool_)) else [bool(v) for v in x]
    return list(dict.fromkeys(vals))


class EFTLSS:
    """The batched core of the theory tree.  `eftpipe_b200.cobaya` wraps it into Cobaya components with the reference's
    class layout (EFTLSS -> EFTLeafKernel -> EFTLeaf); it can also be driven directly:

        th = EFTLSS(tracers).must_provide(requirements).initialize()
        th.calculate({tracer: dict(pkh=(B, 200), f=, DA=, H=) or a BoltzmannExtractor})
        ls, k, Plk = th.get_nonlinear_Plk_grid(tracer, params, chained=False, binned=True)

    Every (chained, binned) product a tracer is asked for (theory.py:590-604), the un-binned interpolation points and the
    window / fibre snapshots are row blocks of ONE stacked projection operator (plan.stack_projections): one pass of the
    fused pipeline per tracer yields them all."""

    def __init__(self, tracers: dict, cache_dir_path=None):
        tracers = deepcopy(tracers)
        default = tracers.pop("default", {})
        self.tracers = {name: _merge(default, cfg or {}) for name, cfg in tracers.items()}
        if not self.tracers:
            raise LoggedError("No tracer specified")
        for name, cfg in self.tracers.items():
            unknown = set(cfg) - _KNOWN
            if unknown:
                raise LoggedError(f"tracer {name}: unknown configuration keys {sorted(unknown)}")
            cross = cfg.get("cross", False)  # theory.py:145-155
            if not isinstance(cross, bool):
                if not isinstance(cross, (list, tuple)) or len(cross) != 2:
                    raise LoggedError(f"tracer {name}: expect a list of 2 elements, but given cross={cross!r}")
                if diff := set(cross).difference(self.tracers):
                    raise LoggedError(f"tracer {name}: cross={cross!r} contains unknown tracer names: {diff!r}")
        self.cache_dir_path = cache_dir_path
        self.requirements = {}
        self.bases, self.commons, self.plans, self.info = {}, {}, {}, {}
        self._state = {}
        self.B = 0

    # ---- requirements (theory.py:165-194, :497-555, :773-827) ----
    def must_provide(self, requirements: dict):
        """Reference grammar: `{product: {tracer: {"ls": [...], "chained": bool | [bool], "binned": bool | [bool], "binning":
        {"kout": ...}}}}` for `nonlinear_Plk_grid`, `nonlinear_Plk_gaussian_grid`, `nonlinear_Plk_interpolator`; `snapshots`,
        `bird_component`, `eft_params_values_dict` take no settings.  A `default` tracer entry applies to every tracer named
        in the same product (theory.py:188-192).

        One extension: `nonlinear_Plk_interpolator` accepts `kout`, the abscissae the consumer will evaluate the
        interpolator at.  The reference hands out a callable; on the batched path an interpolation at fixed points is a
        fixed operator composed into the projection (plan.interp_matrices), which the un-binned likelihood read-out
        (likelihood.py:503-547) uses.  Without `kout` the interpolator is built from the un-binned grid product, as in
        theory.py:862-871."""
        for product, per_tracer in requirements.items():
            if product not in ("nonlinear_Plk_grid", "nonlinear_Plk_gaussian_grid", "nonlinear_Plk_interpolator", "snapshots",
                               "bird_component", "eft_params_values_dict"):
                raise LoggedError(f"Unexpected requirement {product}")
            per_tracer = dict(per_tracer or {})
            default = per_tracer.pop("default", None)
            for tracer, cfg in per_tracer.items():
                if tracer not in self.tracers:
                    raise LoggedError(f"Unknown tracer name: {tracer}")
                cfg = _merge(default or {}, cfg or {}) if default else dict(cfg or {})
                req = self.requirements.setdefault(tracer, dict(Nl=0, No=0, chained=[], binned=[], binning=None, interp={},
                                                                snapshots=False, bird_component=False, order=[]))
                if product in ("snapshots", "bird_component"):
                    req[product] = True
                    continue
                if product == "eft_params_values_dict":
                    continue
                ls = [cfg["ls"]] if isinstance(cfg["ls"], (int, np.integer)) else list(cfg["ls"])  # mandatory key
                chained = _bool_list(cfg.get("chained", False))
                if chained == [True]:  # ls are chained multipoles: the power spectrum needs one more (theory.py:511-512)
                    ls = sorted(set(ls + [l + 2 for l in ls]))
                if any(l % 2 == 1 for l in ls):
                    raise LoggedError(f"Invalid multipoles: {ls}")
                if max(ls) > 4:
                    raise LoggedError(f"Unsupported multipoles: {ls}")
                Nl = max(ls) // 2 + 1
                req["Nl"], req["No"] = max(req["Nl"], Nl), max(req["No"], Nl)
                binned = _bool_list(cfg.get("binned", False))
                if product == "nonlinear_Plk_interpolator":
                    if True in binned:
                        raise LoggedError("binned Plk interpolator not supported")
                    if cfg.get("kout") is not None:
                        for c in chained:
                            kout = np.asarray(cfg["kout"], float)
                            if c in req["interp"] and not np.array_equal(req["interp"][c], kout):
                                raise LoggedError("does not support multiple different interpolation grids per tracer")
                            req["interp"][c] = kout
                            req["order"].append(("interp", c, False))
                        req["chained"] = list(dict.fromkeys(req["chained"] + chained))
                        continue
                if True in binned:
                    binning = cfg.get("binning")
                    if binning is None:
                        raise LoggedError("binned=True but missing binning")
                    if "kout" not in binning and "kedges" not in binning:
                        raise LoggedError("missing kout in binning")
                    old = req["binning"]
                    if old is not None:  # theory.py:536-548: different binning settings are not allowed
                        same = set(old) == set(binning) and all(np.array_equal(np.asarray(old[k]), np.asarray(binning[k])) for k in old)
                        if not same:
                            raise LoggedError("does not support multiple different binning requirements")
                    req["binning"] = deepcopy(binning)
                req["chained"] = list(dict.fromkeys(req["chained"] + chained))
                req["binned"] = list(dict.fromkeys(req["binned"] + binned))
                req["order"] += [("grid", c, b) for c in chained for b in binned]
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

    def build_basis(self, name):
        """theory.py:284-294"""
        cfg = self.tracers[name]
        cross = cfg.get("cross")
        cross_prefix = [related_prefix(t, self.tracers[t]) for t in cross] if isinstance(cross, (list, tuple)) else []
        return find_param_basis(cfg.get("basis", "westcoast"))(prefix=tracer_prefix(name, cfg), cross_prefix=cross_prefix)

    def initialize(self):
        import itertools

        for name, cfg in self.tracers.items():
            self.bases[name] = self.build_basis(name)
            req = self.requirements.get(name)
            if req is None or req["No"] == 0:
                continue
            No = req["No"]
            Nl = max(cfg.get("Nl", 0) or 0, req["Nl"])
            kmA, krA, ndA, kmB, krB, ndB = self._scales(name)
            basis = self.bases[name]
            cross = cfg.get("cross")
            counterform = cfg.get("counterform") or basis.counterform()
            co = Common(Nl=Nl, No=No, kmax=cfg.get("kmax", 0.3), kmA=kmA, krA=krA, ndA=ndA, kmB=kmB, krB=krB, ndB=ndB,
                        counterform=counterform, with_NNLO=bool(cfg.get("with_NNLO", False)),
                        optiresum=bool(cfg.get("optiresum", False)), IRcutoff=cfg.get("IRcutoff", False),
                        kIR=cfg.get("kIR"))  # theory.py:421-437
            snaps = {}
            rs = dict(cfg.get("IRresum") or {})
            with_resum = bool(cfg.get("with_IRresum", True))
            if with_resum and rs.get("snapshot"):
                snaps["IRresum"] = ("stage", 0)
            ap = None
            if cfg.get("with_APeffect"):
                apc = dict(cfg.get("APeffect") or {})
                if apc.get("z_AP") is None:
                    apc["z_AP"] = cfg["z"]  # theory.py:458: z_AP defaults to the tracer's z
                apo = _construct(APeffect, apc, co=Common(Nl=Nl))
                ap = dict(DA=apo.DA, H=apo.H, nbinsmu=apc.get("nbinsmu", 200), accboost=apc.get("accboost", 1), APst=apo.APst)
                self.info.setdefault(name, {})["ap"] = apo
                if apc.get("snapshot"):
                    snaps["APeffect"] = ("stage", 1)
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
            g = P.GridConfig(Nl=Nl, kmax=cfg.get("kmax", 0.3), with_NNLO=co.with_NNLO, optiresum=co.optiresum)
            bo = Binning(co=co, **req["binning"]) if req["binning"] is not None else None

            def project(binm, chained, with_fiber=True):
                fm = fiber.matrix() if (fiber is not None and with_fiber) else None
                fst = fiber.fiberst if (fiber is not None and with_fiber) else False
                if custom is not None:
                    return P.compose_projection(g, window=custom["matrix"], icc=None, binning=binm, chained=chained, fiber=fm, fiber_st=fst,
                                                window_stoch=custom["matrix_st"], window_picc=custom["picc"])
                return P.compose_projection(
                    g, window=None if window is None else window_matrix(window),
                    icc=None if icc is None else dict(matrix=icc.effective_matrix(), PSN_times_Pshot=icc.PSN),
                    binning=binm, chained=chained, window_st=True if window is None else window.window_st, fiber=fm, fiber_st=fst)

            # the products, in the order they were asked for (the first one is the tracer's default product)
            keys = list(dict.fromkeys(req["order"] + [("grid", c, b) for c, b in itertools.product(req["chained"], req["binned"])]))
            blocks, meta = [], {}
            for key in keys:
                kind, chained, binned = key
                if kind == "interp":
                    # rows [0, nkout) of every multipole = PlkInterpolator (theory.py:75-106), rows [nkout, 2 nkout) = the
                    # plain cubic interpolation of the marginalised rows (likelihood.py:510-513)
                    keff = req["interp"][chained]
                    binm = np.vstack(P.interp_matrices(co.k, keff))
                elif binned:
                    binm, keff = bo.matrix, bo.keff
                else:
                    binm, keff = None, co.k
                blk = project(binm, chained)
                nl_out, nk = blk["shape"]
                nmult = (No - 1) if chained else No  # theory.py:599-603
                meta[key] = dict(nl_out=nl_out, nk=nk, nout=nl_out * nk, nterm=g.nterm, kout=np.asarray(keff, float),
                                 ls=[2 * i for i in range(min(nmult, nl_out))], No=No, interp_nk=len(keff) if kind == "interp" else None)
                blocks.append(blk)
            if req["snapshots"]:  # theory.py:260, :576-581: what the plugins were told to keep (`snapshot: True`)
                for plug, wcfg, with_f in (("window", cfg.get("window"), False), ("fiber", cfg.get("fiber"), True)):
                    if (wcfg or {}).get("snapshot") and (window is not None or custom is not None) and (plug == "window" or fiber is not None):
                        key = ("snapshot", plug, False)
                        blk = project(None, False, with_fiber=with_f)
                        meta[key] = dict(nl_out=blk["shape"][0], nk=blk["shape"][1], nout=blk["shape"][0] * blk["shape"][1], nterm=g.nterm,
                                         kout=co.k, ls=[2 * i for i in range(blk["shape"][0])], No=No, interp_nk=None)
                        blocks.append(blk)
                        snaps[plug] = ("block", key)
            trivial = window is None and custom is None and fiber is None and all(k == ("grid", False, False) for k in meta)
            proj = None
            if not trivial:
                proj = P.stack_projections(blocks)
                for key, off0, off1, blk in zip(meta, proj["offsets"][:-1], proj["offsets"][1:], blocks):
                    meta[key].update(row0=int(off0), row1=int(off1), picc=np.asarray(blk["picc"], float).reshape(-1))
            else:
                for key in meta:
                    meta[key].update(row0=0, row1=Nl * g.Nk, picc=np.zeros(Nl * g.Nk))
            host = P.build_tracer_plan(Nl=Nl, kmax=cfg.get("kmax", 0.3), with_NNLO=co.with_NNLO, with_resum=with_resum,
                                       resum_NFFT=rs.get("NFFT", 192), ap=ap, projection=proj, optiresum=co.optiresum,
                                       ircutoff=co.IRcutoff, kIR=co.kIR, lambda_ir=rs.get("LambdaIR", P.LAMBDA_IR))
            self.commons[name] = co
            self.plans[name] = DevicePlan(host)
            first = meta[keys[0]]
            self.info.setdefault(name, {}).update(products=meta, default=keys[0], snapshots=snaps, **first)
        return self

    def _product(self, tracer, chained=None, binned=None, interp=False):
        info = self.info[tracer]
        if chained is None and binned is None and not interp:
            return info["products"][info["default"]]
        key = ("interp", bool(chained), False) if interp else ("grid", bool(chained), bool(binned))
        try:
            return info["products"][key]
        except KeyError:
            what = "nonlinear_Plk_interpolator" if interp else "nonlinear_Plk_grid"
            raise LoggedError(f"{what} (chained={bool(chained)}, binned={bool(binned)}) of tracer {tracer} not computed, please check if "
                              "you have specified it in requirements")

    def product_info(self, tracer, chained=None, binned=None, interp=False):
        return self._product(tracer, chained, binned, interp)

    # ---- per batch (theory.py:557-609) ----
    def calculate(self, cosmo: dict, reset=True):
        """cosmo[tracer] = dict(pkh=(B, 200) on kh = logspace(-5, 0, 200), f=, DA=, H= (B,) [, rdrag, h]), or a
        boltzmann.BoltzmannExtractor.  reset=False keeps the state of the tracers not named in `cosmo` (the Cobaya
        components evaluate one tracer per call)."""
        if reset or not hasattr(self, "_derive_inputs"):
            self._state, self._derive_inputs = {}, {}
        self._derived = None
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
          if name not in cosmo:
              continue
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
        self._derived[prefix + "fsigma8_z"] = np.asarray(_host(c["fsigma8_z"]), float) if c.get("fsigma8_z") is not None else -1

    def get_bird_component(self, tracer, params, chained=None, binned=None):
        """(ls, k, BirdComponent) - theory.py:265-266, :844-847"""
        prod = self._product(tracer, chained, binned)
        return prod["ls"], prod["kout"], self.bases[tracer].reduce_Plk(self._view(tracer, prod), params)

    def get_eft_params_values_dict(self, tracer, params):
        """the tracer's EFT parameters, absent ones as 0.0 (theory.py:262-263, :839-843)"""
        basis = self.bases[tracer]
        names = list(basis.gaussian_params()) + list(basis.non_gaussian_params())
        return {n: params.get(n, 0.0) for n in names}

    def get_snapshots(self, tracer):
        """theory.py:260-261: {name: BirdSnapshot-like view} of what the plugins were configured to keep (`snapshot: True` in
        the IRresum / APeffect / window / fiber blocks; request `snapshots` in must_provide).  "IRresum" and "APeffect" are the
        term arrays the fused pipeline left in its workspace (eftb_workspace_terms), "window" and "fiber" are extra row blocks
        of the projection."""
        from .transformer import PlainBird

        info = self.info[tracer]
        if not self.requirements.get(tracer, {}).get("snapshots"):
            raise LoggedError("snapshots not computed, please check if you have specified it in requirements")
        co, dp = self.commons[tracer], self.plans[tracer]
        bm, f_bm = self._state[tracer]
        out = {}
        for name, (kind, what) in info["snapshots"].items():
            if kind == "stage":
                T = dp.stage_terms(self.B, what)
                snap = PlainBird(None, co, T, np.zeros((co.Nl, co.Nk)), self.B, False, f_bm)
            else:
                snap = self._view(tracer, info["products"][what])
            snap.k, snap.ls = co.k.copy(), [2 * i for i in range(co.Nl)]
            out[name] = snap
        return out

    def get_nonlinear_Plk_terms(self, tracer, chained=None, binned=None, interp=False):
        """(batch-minor terms (rows, nterm, Bp) of the product, growth rate (Bp,)) - what the likelihood kernels read"""
        prod = self._product(tracer, chained, binned, interp)
        bm, f_bm = self._state[tracer]
        return bm[prod["row0"] : prod["row1"]], f_bm

    def get_nonlinear_Plk_grid(self, tracer, params, chained=None, binned=None, interp=False):
        """(ls, k, Plk (B, No, nk)) - theory.py:244-252 + EFTLeaf reduction (theory.py:846-860)."""
        prod = self._product(tracer, chained, binned, interp)
        plk = self.bases[tracer].reduce_Plk(self._view(tracer, prod), params).sum()
        if prod["interp_nk"] is not None:  # interpolated products: the PlkInterpolator rows
            plk = plk[..., : prod["interp_nk"]]
        return prod["ls"], prod["kout"], plk[..., : len(prod["ls"]), :]

    def get_nonlinear_Plk_gaussian_grid(self, tracer, params, chained=None, binned=None, interp=False):
        prod = self._product(tracer, chained, binned, interp)
        table = self.bases[tracer].reduce_Plk_gaussian_table(self._view(tracer, prod), params)
        if prod["interp_nk"] is not None:  # the rows interpolated without the inserted origin
            table = {name: v[..., prod["interp_nk"]:] for name, v in table.items()}
        return prod["ls"], prod["kout"], {name: v[..., : len(prod["ls"]), :] for name, v in table.items()}

    def get_nonlinear_Plk_interpolator(self, tracer, params, chained=False):
        """theory.py:254-258, :862-871: a `PlkInterpolator` over the tracer's un-binned grid product"""
        ls, k, plk = self.get_nonlinear_Plk_grid(tracer, params, chained=chained, binned=False)
        return PlkInterpolator(ls, k, plk)

    def _view(self, tracer, prod=None):
        from .transformer import PlainBird

        prod = prod or self._product(tracer)
        bm, f_bm = self._state[tracer]
        T = bm[prod["row0"] : prod["row1"]].reshape(prod["nl_out"], prod["nk"], prod["nterm"], bm.shape[-1])
        # the reduction honours co.No: chained products expose one multipole fewer (theory.py:599-602)
        return PlainBird(None, self.commons[tracer], T, np.asarray(prod["picc"]).reshape(prod["nl_out"], prod["nk"]), self.B, False, f_bm)


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
