"""Device engine: uploads a host `TracerPlan` into a `libeftb200` plan and drives the CUDA stages
on torch CUDA tensors (torch is only the allocator / stream provider here).

Arrays crossing this API are torch float64 CUDA tensors.  "Batch-minor" tensors have shape
(..., Bp) with Bp = padded_batch(B); public results are returned point-major (B, ...).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .plan import N22, NCH, TracerPlan


def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def capture_graph(fn, warmup=2):
    """Capture `fn()` (a fixed sequence of library calls on static device tensors) into a CUDA graph.  Returns
    (graph, outputs): `graph.replay()` re-runs the ~20 kernel launches of a pipeline step with one launch call,
    `outputs` are the static tensors `fn` returned.  The library's launchers are capture-safe after their first
    call (function attributes and the anti-diagonal schedule are set up once, no synchronisation, no allocation)."""
    torch = _lib.require_cuda()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    lib = _lib.load()
    n0 = lib.eftb_launch_count()
    with torch.cuda.graph(graph):
        out = fn()
    # kernels of this library recorded into the graph = kernels every replay launches (bench.py `gpu_launches`)
    graph.library_launches = int(lib.eftb_launch_count() - n0)
    # the graph holds raw pointers into the plans / workspaces `fn` closes over: keep them alive as long as the graph
    graph.keepalive = fn
    return graph, out


class HostPipeline:
    """Double-buffered host -> device -> host driver for a fixed-shape evaluation `fn(*device_inputs) -> device outputs`
    (what an MCMC / importance-sampling loop that produces its points on the host needs to keep the GPU busy).

    Per slot: ONE packed pinned input buffer (`host_in(slot)` returns the named views to fill), one device copy of it,
    a CUDA graph of `fn` on views of that device buffer, pinned output buffers.  `submit(slot)` enqueues the slot's
    H2D copy on a copy stream, the graph replay on the current stream and the D2H copies on a second copy stream, so
    the copies of step i+1 / i-1 overlap the kernels of step i; `wait(slot)` blocks until the slot's outputs are on the
    host.  Slots share the library workspaces, so replays are serialised on the compute stream by construction."""

    def __init__(self, fn, shapes: dict, nslots=2, use_graph=True):
        torch = _lib.require_cuda()
        self.use_graph = use_graph
        self.torch, self.names = torch, list(shapes)
        sizes = [int(np.prod(shapes[n])) for n in self.names]
        offs = np.concatenate([[0], np.cumsum(sizes)])
        self.copy_in, self.copy_out = torch.cuda.Stream(), torch.cuda.Stream()
        self.h_in, self.d_in, self.graphs, self.d_out, self.h_out = [], [], [], [], []
        self.ev_h2d, self.ev_comp, self.ev_d2h = [], [], []
        view = lambda buf: {n: buf[offs[i] : offs[i + 1]].view(*shapes[n]) for i, n in enumerate(self.names)}
        for _ in range(nslots):
            h = torch.empty(int(offs[-1]), dtype=torch.float64).pin_memory()
            d = torch.empty(int(offs[-1]), dtype=torch.float64, device="cuda")
            self.h_in.append((h, view(h)))
            self.d_in.append((d, view(d)))
        self._fn = fn
        self._captured = False

    def host_in(self, slot):
        """dict name -> pinned host view to fill before `submit(slot)`"""
        return self.h_in[slot][1]

    def _capture(self):
        torch = self.torch
        for slot in range(len(self.h_in)):
            self.d_in[slot][0].copy_(self.h_in[slot][0])
        torch.cuda.synchronize()
        for slot in range(len(self.h_in)):
            dv = self.d_in[slot][1]
            if self.use_graph:
                graph, outs = capture_graph(lambda dv=dv: self._fn(**dv))
            else:  # eager launches (profiling): the outputs are re-produced by every submit
                graph, outs = None, self._fn(**dv)
            outs = list(outs) if isinstance(outs, (tuple, list)) else [outs]
            self.graphs.append(graph)
            self.d_out.append(outs)
            self.h_out.append([torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs])
            self.ev_h2d.append(torch.cuda.Event())
            self.ev_comp.append(torch.cuda.Event())
            self.ev_d2h.append(torch.cuda.Event())
        self._captured = True

    def submit(self, slot):
        torch = self.torch
        if not self._captured:
            self._capture()
        compute = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_in):
            self.copy_in.wait_event(self.ev_comp[slot])  # the previous replay of this slot has consumed its inputs
            self.d_in[slot][0].copy_(self.h_in[slot][0], non_blocking=True)
            self.ev_h2d[slot].record(self.copy_in)
        compute.wait_event(self.ev_h2d[slot])
        compute.wait_event(self.ev_d2h[slot])            # ... and its previous outputs are already on the host
        if self.graphs[slot] is not None:
            self.graphs[slot].replay()
        else:
            outs = self._fn(**self.d_in[slot][1])
            self.d_out[slot] = list(outs) if isinstance(outs, (tuple, list)) else [outs]
        self.ev_comp[slot].record(compute)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(self.ev_comp[slot])
            for h, d in zip(self.h_out[slot], self.d_out[slot]):
                h.copy_(d, non_blocking=True)
            self.ev_d2h[slot].record(self.copy_out)

    def wait(self, slot):
        """block until the outputs of the last `submit(slot)` are in `outputs(slot)`"""
        self.ev_d2h[slot].synchronize()
        return self.h_out[slot]

    def join(self):
        """make the current stream wait for every outstanding read-back (for timing with stream events)"""
        compute = self.torch.cuda.current_stream()
        for ev in self.ev_d2h:
            compute.wait_event(ev)

    @property
    def h2d_bytes(self):
        return self.h_in[0][0].numel() * 8

    @property
    def d2h_bytes(self):
        return sum(h.numel() * h.element_size() for h in self.h_out[0])


class DevicePlan:
    """One tracer's pipeline on the current CUDA device."""

    def __init__(self, plan: TracerPlan):
        self.torch = _lib.require_cuda()
        self.lib = _lib.load()
        self.host = plan
        g, lay = plan.grid, plan.front
        cfg = _lib.EftbConfig()
        cfg.Nl, cfg.Nk, cfg.Ns, cfg.Nmax, cfg.nterm, cfg.with_nnlo = g.Nl, g.Nk, g.Ns, g.NFFT, g.nterm, int(g.with_NNLO)
        cfg.nin, cfg.ntail, cfg.ntailx, cfg.front_rows = lay.nin, lay.ntail, lay.ntailx, plan.Wf.shape[0]
        rows = lay.rows
        cfg.row_cre, cfg.row_cim, cfg.row_p11, cfg.row_p13 = rows["cre"][0], rows["cim"][0], rows["P11"][0], rows["P13raw"][0]
        cfg.row_c11, cfg.row_cct = rows["C11"][0], rows["Cct"][0]
        cfg.row_cctnnlo = rows["CctNNLO"][0] if g.with_NNLO else -1
        cfg.row_x, cfg.row_y = rows["X"][0], rows["Y"][0]
        cfg.row_cre_cf, cfg.row_cim_cf = (rows["cre_cf"][0], rows["cim_cf"][0]) if "cre_cf" in rows else (-1, -1)
        aux = plan.front_aux
        cfg.inv_dlog, cfg.wx_last, cfg.wx_prev = aux["inv_dlog"], aux["wX_last"], aux["wX_prev"]
        cfg.npair = plan.pair_table.shape[0]
        keep = []  # host arrays must outlive the create call

        def ptr(a, ctype=C.c_double, dtype=np.float64):
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return _lib.as_ptr(a, ctype)

        cst = _lib.EftbConstants()
        cst.k, cst.l11, cst.lct, cst.lctnnlo = ptr(g.k), ptr(g.l11), ptr(g.lct), ptr(g.lctNNLO)
        cst.l22, cst.l13 = ptr(g.l22), ptr(g.l13)
        cst.Wf, cst.lr, cst.lrx = ptr(plan.Wf), ptr(aux["lr"]), ptr(aux["lrx"])
        cst.pair_table = ptr(plan.pair_table.view(np.float64))
        cst.pair_offsets = ptr(plan.pair_offsets, C.c_int32, np.int32)
        cst.Ak, cst.As = ptr(plan.Ak), ptr(plan.As)
        if plan.resum is not None:
            rs = plan.resum
            cfg.has_resum, cfg.NIR, cfg.Na, cfg.Nkr, cfg.Nklow, cfg.qdeg = 1, rs["NIR"], rs["Na"], g.Nkr, g.Nklow, rs["q"].shape[-1]
            cst.R, cst.q, cst.kr2 = ptr(rs["R"]), ptr(rs["q"]), ptr(rs["kr2"])
        if plan.ap is not None:
            ap = plan.ap
            cfg.has_ap, cfg.nmu, cfg.nint, cfg.ap_st = 1, ap["mu"].size, ap["nint"], int(plan.ap_st)
            cfg.da_fid, cfg.h_fid = plan.ap_fid
            cst.Cinv, cst.knot_lo, cst.basis, cst.mu, cst.wl = ptr(ap["Cinv"]), ptr(ap["knot_lo"]), ptr(ap["basis"]), ptr(ap["mu"]), ptr(ap["wl"])
        if plan.project is not None:
            cfg.has_project, cfg.nout, cfg.nl_out = 1, plan.project.shape[0], plan.out_shape[0]
            cst.project = ptr(plan.project)
            if plan.project_stoch is not None:
                cst.project_st = ptr(plan.project_stoch)
        handle = C.c_void_p()
        _lib.check(self.lib.eftb_plan_create(C.byref(cfg), C.byref(cst), C.byref(handle)), "eftb_plan_create")
        self.handle, self.cfg = handle, cfg
        self._ws = None
        self._ap_scratch = None
        self._rs_scratch = None
        if plan.project is not None:
            self.out_shape = (plan.out_shape[0], g.nterm, plan.out_shape[1])
        else:
            self.out_shape = (g.Nl, g.nterm, g.Nk)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.eftb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def padded(self, B):
        return self.lib.eftb_padded_batch(int(B))

    def _empty(self, *shape):
        return self.torch.empty(shape, dtype=self.torch.float64, device="cuda")

    def _dev(self, a):
        t = self.torch
        if isinstance(a, t.Tensor):
            return a.to(device="cuda", dtype=t.float64).contiguous()
        return t.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device="cuda")

    def to_batch_minor(self, x):
        """(B, R) point-major -> (R, Bp)."""
        x = self._dev(x)
        if x.dim() == 1:
            x = x[:, None]
        B, R = x.shape
        out = self._empty(R, self.padded(B))
        _lib.check(self.lib.eftb_to_batch_minor(_p(x), B, R, _p(out), _stream_ptr(self.torch)), "to_batch_minor")
        return out

    def to_point_major(self, x, B):
        """(R, Bp) -> (B, R)."""
        R = x.shape[0]
        out = self._empty(B, R)
        _lib.check(self.lib.eftb_to_point_major(_p(x), B, R, C.c_void_p(0), _p(out), _stream_ptr(self.torch)), "to_point_major")
        return out

    # ------------------------------------------------------------------ stages (batch-minor tensors)
    def front(self, plin):
        plin = self._dev(plin)
        B = plin.shape[0]
        Bp = self.padded(B)
        u = self._empty(self.host.front.K, Bp)
        F = self._empty(self.cfg.front_rows, Bp)
        _lib.check(self.lib.eftb_front(self.handle, B, _p(plin), _p(u), _p(F), _stream_ptr(self.torch)), "eftb_front")
        return F

    @property
    def has_cf_set(self):
        """IRcutoff "loop" / "resum": the configuration-space terms use their own FFTLog coefficients"""
        return self.cfg.row_cre_cf >= 0

    def antidiag(self, F, B, cf_set=False):
        D = self._empty(NCH, self.cfg.Nmax + 1, 2, self.padded(B))
        fn = self.lib.eftb_antidiag_cf if cf_set else self.lib.eftb_antidiag
        _lib.check(fn(self.handle, B, _p(F), _p(D), _stream_ptr(self.torch)), "eftb_antidiag")
        return D

    def spectral(self, D, B, Dcf=None):
        Bp = self.padded(B)
        P22 = self._empty(N22, self.cfg.Nk, Bp)
        Cs = self._empty(self.cfg.Nl, NCH, self.cfg.Ns, Bp)
        _lib.check(self.lib.eftb_spectral(self.handle, B, _p(D), _p(Dcf), _p(P22), _p(Cs), _stream_ptr(self.torch)), "eftb_spectral")
        return P22, Cs

    def spectral_grouped(self, D, f_bm, B, Dcf=None):
        """fused-path variant of `spectral`: returns (P22, Cr) with the Cloopl rows of Cr already filled"""
        Bp = self.padded(B)
        c = self.cfg
        P22 = self._empty(N22, c.Nk, Bp)
        Dg = self._empty(c.Nl, 12, c.Nmax + 1, 2, Bp)
        Cr = self._empty(Bp, c.Nl, 14 + c.with_nnlo, c.Ns)  # point-major
        _lib.check(self.lib.eftb_spectral_grouped(self.handle, B, _p(D), _p(Dcf), _p(f_bm), _p(Dg), _p(P22), _p(Cr),
                                                  _stream_ptr(self.torch)), "eftb_spectral_grouped")
        return P22, Cr

    def group(self, F, P22, Cs, f_bm, B, Cr=None):
        """Cs=None with a `Cr` from `spectral_grouped`: only the k-space rows and C11/Cct are assembled"""
        Bp = self.padded(B)
        c = self.cfg
        T = self._empty(c.Nl, c.Nk, c.nterm, Bp)
        if Cr is None:
            Cr = self._empty(Bp, c.Nl, 14 + c.with_nnlo, c.Ns)  # point-major
        _lib.check(self.lib.eftb_group(self.handle, B, _p(F), _p(P22), _p(Cs), _p(f_bm), _p(T), _p(Cr),
                                       _stream_ptr(self.torch)), "eftb_group")
        return T, Cr

    def resum(self, F, Cr, f_bm, T, B):
        need = self.lib.eftb_resum_scratch_bytes(self.handle, int(B))
        if self._rs_scratch is None or self._rs_scratch.numel() * 8 < need:
            self._rs_scratch = self.torch.empty((need + 7) // 8, dtype=self.torch.float64, device="cuda")
        _lib.check(self.lib.eftb_resum(self.handle, B, _p(F), _p(Cr), _p(f_bm), _p(T), _p(self._rs_scratch),
                                       _stream_ptr(self.torch)), "eftb_resum")
        return T

    def ap(self, T, DA_bm, H_bm, B):
        need = self.lib.eftb_ap_scratch_bytes(self.handle, int(B))
        if self._ap_scratch is None or self._ap_scratch.numel() * 8 < need:
            self._ap_scratch = self.torch.empty((need + 7) // 8, dtype=self.torch.float64, device="cuda")
        out = self.torch.empty_like(T)
        _lib.check(self.lib.eftb_ap(self.handle, B, _p(T), _p(DA_bm), _p(H_bm), _p(self._ap_scratch), _p(out),
                                    _stream_ptr(self.torch)), "eftb_ap")
        return out

    def project(self, T, B):
        out = self._empty(self.cfg.nout, self.cfg.nterm, self.padded(B))
        _lib.check(self.lib.eftb_project(self.handle, B, _p(T), _p(out), _stream_ptr(self.torch)), "eftb_project")
        return out

    # ------------------------------------------------------------------ fused pipeline
    def workspace(self, B):
        need = self.lib.eftb_workspace_bytes(self.handle, int(B))
        if self._ws is None or self._ws.numel() * 8 < need:
            self._ws = self.torch.empty((need + 7) // 8, dtype=self.torch.float64, device="cuda")
        return self._ws, need

    def eval_terms(self, plin, f, DA=None, H=None, want_bm=False, want_pm=True, out_pm=None, out_bm=None):
        """theory.py:557-609 for a batch.  plin (B, nin), f/DA/H (B,) - device tensors or arrays.
        Returns (terms_pm (B, Nl_out, nterm, nk_out) or None, terms_bm (rows, nterm, Bp) or None)."""
        t = self.torch
        plin = self._dev(plin)
        B = plin.shape[0]
        f = self._dev(f)
        DA = self._dev(DA) if DA is not None else None
        H = self._dev(H) if H is not None else None
        ws, need = self.workspace(B)
        Bp = self.padded(B)
        if want_pm and out_pm is None:
            out_pm = self._empty(B, *self.out_shape)
        if want_bm and out_bm is None:
            rows = self.cfg.nout if self.cfg.has_project else self.cfg.Nl * self.cfg.Nk
            out_bm = self._empty(rows, self.cfg.nterm, Bp)
        _lib.check(
            self.lib.eftb_eval_terms(self.handle, B, _p(plin), _p(f), _p(DA), _p(H), _p(out_bm if want_bm else None),
                                     _p(out_pm if want_pm else None), _p(ws), need, _stream_ptr(t)),
            "eftb_eval_terms",
        )
        return (out_pm if want_pm else None), (out_bm if want_bm else None)


    def stage_terms(self, B, stage):
        """(Nl, Nk, nterm, Bp) batch-minor term arrays the last `eval_terms` left in the workspace: stage 0 = after the IR
        resummation, 1 = after AP (the reference's "IRresum" / "APeffect" Bird snapshots)"""
        if self._ws is None:
            raise RuntimeError("stage_terms: no evaluation has run on this plan yet")
        need = self.lib.eftb_workspace_bytes(self.handle, int(B))
        out = self._empty(self.cfg.Nl, self.cfg.Nk, self.cfg.nterm, self.padded(B))
        _lib.check(self.lib.eftb_workspace_terms(self.handle, B, _p(self._ws), need, int(stage), _p(out), _stream_ptr(self.torch)),
                   "eftb_workspace_terms")
        return out


class DeviceLikelihood:
    """likelihood.py EFTLike.calculate + marginal.py for a batch of points (see likelihood.py here)."""

    def __init__(self, spec: dict):
        self.torch = _lib.require_cuda()
        self.lib = _lib.load()
        self.spec = spec
        cfg = _lib.EftbLikeConfig()
        cfg.ntracer, cfg.ndata, cfg.ngauss = spec["ntracer"], spec["ndata"], spec["ngauss"]
        cfg.npar, cfg.jeffreys = spec["npar"], int(spec["jeffreys"])
        keep = []

        def ptr(a, ctype, dtype):
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return _lib.as_ptr(a, ctype)

        dp = lambda a: ptr(a, C.c_double, np.float64)
        ip = lambda a: ptr(a, C.c_int32, np.int32)
        cst = _lib.EftbLikeConstants()
        cst.nout, cst.nterm, cst.scales, cst.par_index = ip(spec["nout"]), ip(spec["nterm"]), dp(spec["scales"]), ip(spec["par_index"])
        cst.eastcoast, cst.d_tracer, cst.d_row = ip(spec["eastcoast"]), ip(spec["d_tracer"]), ip(spec["d_row"])
        cst.data, cst.picc, cst.invcov = dp(spec["data"]), dp(spec["picc"]), dp(spec["invcov"])
        ng = max(spec["ngauss"], 1)
        cst.g_count, cst.g_tracer = ip(spec["g_count"] if spec["ngauss"] else np.zeros(1)), ip(spec["g_tracer"] if spec["ngauss"] else np.zeros(2))
        cst.g_term = ip(spec["g_term"] if spec["ngauss"] else np.zeros(6))
        cst.g_var = ip(spec["g_var"] if spec["ngauss"] else np.zeros(6))
        cst.g_coef = dp(spec["g_coef"] if spec["ngauss"] else np.zeros(6))
        cst.sigma_inv = dp(spec["sigma_inv"] if spec["ngauss"] else np.zeros((ng, ng)))
        cst.sigma_inv_mu = dp(spec["sigma_inv_mu"] if spec["ngauss"] else np.zeros(ng))
        cst.mu_sigma_mu = float(spec["mu_sigma_mu"])
        cst.d_row_g = ip(spec.get("d_row_g", spec["d_row"]))
        if spec.get("mode") is not None and np.any(spec["mode"]):  # custom EFT bases: explicit bias columns
            cst.mode, cst.xb_off = ip(spec["mode"]), ip(spec["xb_off"])
            cst.xg_off = ip(spec["xg_off"] if spec["ngauss"] else np.zeros(1))
        handle = C.c_void_p()
        _lib.check(self.lib.eftb_like_create(C.byref(cfg), C.byref(cst), C.byref(handle)), "eftb_like_create")
        self.handle, self.cfg = handle, cfg
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.eftb_like_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _workspace(self, B):
        need = self.lib.eftb_like_workspace_bytes(self.handle, int(B))
        if self._ws is None or self._ws.numel() * 8 < need:
            self._ws = self.torch.empty((need + 7) // 8, dtype=self.torch.float64, device="cuda")
        return self._ws, need

    def _ptr_array(self, tensors):
        arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        return arr

    def eval(self, B, terms_bm, f_bm, nuis_bm, want_bestfit=False, want_fullchi2=False, prior_loc=None, prior_sigma_inv=None):
        """(logp, status, bestfit) or, with want_fullchi2, (logp, status, bestfit, fullchi2) - marginal.py:79-137.
        prior_loc / prior_sigma_inv: (B, nG) device tensors, the per-point Gaussian prior of callable `loc` / `scale`
        (marginal.py:60-77; diagonal of Sigma^-1) in place of the plan constants."""
        t = self.torch
        if (prior_loc is None) != (prior_sigma_inv is None):
            raise ValueError("prior_loc and prior_sigma_inv go together")
        if prior_loc is not None:
            for p in (prior_loc, prior_sigma_inv):
                if tuple(p.shape) != (B, self.cfg.ngauss) or p.dtype != t.float64 or not p.is_cuda or not p.is_contiguous():
                    raise ValueError(f"per-point priors must be contiguous float64 CUDA tensors of shape ({B}, {self.cfg.ngauss})")
        ws, need = self._workspace(B)
        logp = t.empty(B, dtype=t.float64, device="cuda")
        status = t.empty(B, dtype=t.int32, device="cuda")
        best = t.empty((B, self.cfg.ngauss), dtype=t.float64, device="cuda") if want_bestfit else None
        full = t.empty(B, dtype=t.float64, device="cuda") if want_fullchi2 else None
        ta, fa = self._ptr_array(terms_bm), self._ptr_array(f_bm)
        _lib.check(
            self.lib.eftb_like_eval_priors(self.handle, B, ta, fa, _p(nuis_bm), _p(prior_loc), _p(prior_sigma_inv), _p(logp),
                                           _p(best), _p(full), _p(status), _p(ws), need, _stream_ptr(t)),
            "eftb_like_eval",
        )
        return (logp, status, best, full) if want_fullchi2 else (logp, status, best)

    def residuals(self, B):
        """(B, ndata): model minus data, PNG - d, of the last `eval` (read from its workspace, nothing recomputed)"""
        t = self.torch
        out = t.empty((B, self.cfg.ndata), dtype=t.float64, device="cuda")
        _lib.check(self.lib.eftb_like_residuals(self.handle, B, _p(self._ws), _p(out), _stream_ptr(t)), "eftb_like_residuals")
        return out

    def vectors(self, B, terms_bm, f_bm, nuis_bm):
        t = self.torch
        ws, need = self._workspace(B)
        vec = t.empty((B, self.cfg.ndata, self.cfg.ngauss + 1), dtype=t.float64, device="cuda")
        ta, fa = self._ptr_array(terms_bm), self._ptr_array(f_bm)
        _lib.check(self.lib.eftb_like_vectors(self.handle, B, ta, fa, _p(nuis_bm), _p(vec), _p(ws), need, _stream_ptr(t)),
                   "eftb_like_vectors")
        return vec
