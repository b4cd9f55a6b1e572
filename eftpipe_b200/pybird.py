"""Batched, device-resident mirror of `eftpipe.pybird.pybird` (reference file pybird/pybird.py):
`Common`, `Bird`, `NonLinear`, `Resum`, `APeffect` with the reference's constructor keywords, method
names, stage order and error behaviour, so that `theory.py:557-609`-style driver code runs unchanged:

    co = Common(Nl=3, kmax=0.3, kmA=0.7, krA=0.25, ndA=4.5e-5)
    nl, rs = NonLinear(load=False, save=False, co=co), Resum(co=co)
    bird = Bird(kin, Plin, f, DA, H, z, co=co)        # Plin: (200,) or (B, 200); f, DA, H scalars or (B,)
    nl.PsCf(bird); bird.setPsCfl(); rs.Ps(bird); ap.AP(bird); win.Window(bird)

Differences, all additive: a leading batch axis (B cosmologies per call), arrays are torch float64 CUDA
tensors, and the arithmetic runs in libeftb200's sm_100a kernels through the C ABI - there is no CPU
path.  `optiresum`, `IRcutoff`/`kIR` and `LambdaIR` are plan constants: they change the fixed operators the
plan builder composes (eftpipe_b200/plan.py), not the kernels.  With `optiresum` the configuration-space
arrays (`C11`, `Cct`, `C22`, `C13`, `Cloopl`) hold the extracted BAO peak on `co.sr` (what `Resum.extractBAO`,
pybird.py:1382-1400, feeds the resummation) instead of the raw correlation function on `co.s`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, plan as P, synthetic
from .fftlog import FFTLog


def Hubble(Om, z):
    """pybird.py:34-36"""
    return ((Om) * (1 + z) ** 3.0 + (1 - Om)) ** 0.5


def DAfunc(Om, z):
    """pybird.py:39-42"""
    return synthetic.angular_distance(Om, z)


def fN(Om, z):
    """pybird.py:28-31"""
    return synthetic.growth_rate(Om, z)


class Common:
    """pybird.py:486-582 - same keywords, same validation."""

    def __init__(self, Nl=None, No=None, kmax=0.3, optiresum=False, kmA=0.7, krA=0.25, ndA=3e-4, kmB=None, krB=None,
                 ndB=None, counterform="westcoast", with_NNLO=False, kIR=None, IRcutoff=False):
        if IRcutoff and kIR is None:
            raise ValueError("kIR must be specified when doing IRcutoff")
        if IRcutoff is True:  # pybird.py:530-531
            IRcutoff = "all"
        if IRcutoff not in (False, "all", "loop", "resum"):
            raise ValueError(f"unexpected IRcutoff option: {IRcutoff}")  # pybird.py:1160 raises at the first PsCf
        self.optiresum, self.IRcutoff, self.kIR = bool(optiresum), IRcutoff, kIR
        self.kmA, self.krA, self.ndA = kmA, krA, ndA
        self.kmB = kmA if kmB is None else kmB
        self.krB = krA if krB is None else krB
        self.ndB = ndA if ndB is None else ndB
        self.counterform, self.with_NNLO = counterform, with_NNLO
        if Nl is None and No is None:
            self.Nl = self.No = 2
        elif not (Nl is None or No is None):
            self.Nl, self.No = Nl, No
        else:
            self.Nl = self.No = Nl or No
        if self.No > self.Nl:
            raise ValueError("No should always be smaller than Nl")
        self.N11, self.Nct, self.NctNNLO, self.N22, self.N13, self.Nloop = 3, 6, 3, 28, 10, 12
        self.kmax = kmax
        g = P.GridConfig(Nl=self.Nl, kmax=kmax, with_NNLO=with_NNLO, optiresum=self.optiresum)
        self.k, self.s, self.sr, self.kr = g.k, g.s, g.sr, g.kr  # sr: the resummation grid (= s unless optiresum)
        self.Nk, self.Ns, self.Nkr, self.Nklow = g.Nk, g.Ns_full, g.Nkr, g.Nklow
        self.Nsr = g.Ns
        self.l11, self.lct, self.lctNNLO, self.l22, self.l13 = g.l11, g.lct, g.lctNNLO, g.l22, g.l13
        self.nterm = g.nterm
        # stage registry -> one device plan per Common (built lazily, rebuilt when a stage is added)
        self._stages = {}
        self._devices = {}  # FFTLog taper of the loop coefficients -> (version, DevicePlan)
        self._version = 0

    def _register(self, kind, obj):
        self._stages[kind] = obj
        self._version += 1

    def device_plan(self, window=None):
        """the device plan of the stages registered on this Common; `window`: the FFTLog taper `NonLinear.PsCf(bird, window=)`
        was called with (pybird.py:1143; default: the NonLinear object's own) - one plan per distinct value, since the taper
        is part of the front operator"""
        from .engine import DevicePlan

        nl = self._stages.get("nonlinear")
        if nl is None:
            raise RuntimeError("a NonLinear object must be constructed for this Common before evaluation")
        window = nl.window if window is None else window
        dev = self._devices.get(window)
        if dev is None or dev[0] != self._version:
            rs, ap = self._stages.get("resum"), self._stages.get("ap")
            host = P.build_tracer_plan(
                Nl=self.Nl, kmax=self.kmax, NFFT=nl.NFFT, with_NNLO=self.with_NNLO, kin=nl.kin, window=window,
                with_resum=rs is not None, resum_NFFT=rs.NFFT if rs is not None else 192,
                ap=None if ap is None else dict(DA=ap.DA, H=ap.H, nbinsmu=ap._nbinsmu, accboost=ap._accboost, APst=ap.APst),
                loop_cache=nl._loop_cache, optiresum=self.optiresum, ircutoff=self.IRcutoff, kIR=self.kIR,
                lambda_ir=rs.LambdaIR if rs is not None else P.LAMBDA_IR)
            dev = self._devices[window] = (self._version, DevicePlan(host))
        return dev[1]


common = Common()


class _TermsView:
    """Read access to the term arrays with the reference's names and axis order (+ leading batch)."""

    co: Common
    B: int
    _T = None  # (Nl_out, nk_out, nterm, Bp) batch-minor
    _squeeze = False
    _picc = None  # numpy (Nl_out, nk_out), cosmology independent

    def _terms(self, a, b):
        if self._T is None:
            raise AttributeError("term arrays are not available before setPsCfl()")
        v = self._T[:, :, a:b, : self.B].permute(3, 0, 2, 1)
        return v[0] if self._squeeze else v

    def _set_terms(self, a, b, value):
        """the reference's stages rebind (`bird.P11l = ...`, pybird.py:1613) or update in place (`bird.Ploopl += ...`,
        :1445) the term arrays: the getters hand out views of the device array, so in-place updates already land in it;
        rebinding copies the new values in.  `value`: (B, Nl, n, nk) [or (Nl, n, nk) for an unbatched bird], tensor or array."""
        import torch

        if self._T is None:
            raise AttributeError("term arrays are not available before setPsCfl()")
        v = value if isinstance(value, torch.Tensor) else torch.as_tensor(np.asarray(value, float))
        v = v.to(self._T.device, self._T.dtype)
        if self._squeeze:
            v = v[None]
        self._T[:, :, a:b, : self.B] = v.permute(1, 3, 2, 0)

    P11l = property(lambda self: self._terms(0, 3), lambda self, v: self._set_terms(0, 3, v))
    Pctl = property(lambda self: self._terms(3, 9), lambda self, v: self._set_terms(3, 9, v))
    Ploopl = property(lambda self: self._terms(9, 21), lambda self, v: self._set_terms(9, 21, v))
    Pstl = property(lambda self: self._terms(21, 24), lambda self, v: self._set_terms(21, 24, v))

    @property
    def PctNNLOl(self):
        if not self.co.with_NNLO:
            shape = (self._T.shape[0], 3, self._T.shape[1])
            z = self._T.new_zeros((self.B,) + shape)
            return z[0] if self._squeeze else z
        return self._terms(24, 27)

    @PctNNLOl.setter
    def PctNNLOl(self, value):
        if not self.co.with_NNLO:
            raise AttributeError("this bird carries no NNLO counterterm rows (co.with_NNLO is False)")
        self._set_terms(24, 27, value)

    @property
    def Picc(self):
        import torch

        p = self._picc if self._picc is not None else np.zeros((self._T.shape[0], self._T.shape[1]))
        t = torch.as_tensor(p, device="cuda")
        return t if self._squeeze else t.expand(self.B, *t.shape)

    @Picc.setter
    def Picc(self, value):
        """`bird.Picc = bird.Picc - x` (window.py:405): Picc is one constant array per plan here (cosmology independent)"""
        v = value.detach().cpu().numpy() if hasattr(value, "detach") else np.asarray(value, float)
        if v.ndim == 3:
            if not np.allclose(v, v[:1]):
                raise ValueError("Picc must not depend on the cosmology")
            v = v[0]
        self._picc = np.array(v, float)

    def add_Picc(self, delta):
        base = self._picc if self._picc is not None else np.zeros((self._T.shape[0], self._T.shape[1]))
        self._picc = base + np.asarray(delta, float)

    def terms_point_major(self):
        """(B, Nl_out, nterm, nk_out): P11l | Pctl | Ploopl | Pstl [| PctNNLOl] concatenated."""
        return self._T[:, :, :, : self.B].permute(3, 0, 2, 1)


class BirdSnapshot(_TermsView):
    """pybird.py:616-632"""

    def __init__(self, bird):
        self.co, self.f, self.B, self._squeeze = bird.co, bird.f, bird.B, bird._squeeze
        self.k = bird.co.k.copy()
        self.ls = [2 * i for i in range(bird.co.Nl)]
        self._T = bird._T.clone()
        self._picc = None if bird._picc is None else bird._picc.copy()


class Bird(_TermsView):
    """pybird.py:635-866.  `Plin` may carry a leading batch axis; `f`, `DA`, `H` broadcast against it."""

    def __init__(self, kin, Plin, f, DA=None, H=None, z=None, co=common, rdrag=None, h=None):
        self.torch = _lib.require_cuda()
        t = self.torch
        self.co = co
        self.kin = np.asarray(kin, float)
        dev = lambda x: (x if isinstance(x, t.Tensor) else t.as_tensor(np.asarray(x, float))).to("cuda", t.float64)
        Pin = dev(Plin)
        self._squeeze = Pin.dim() == 1
        self.Pin = Pin.reshape(-1, Pin.shape[-1]).contiguous()
        self.B = self.Pin.shape[0]
        bc = lambda x: None if x is None else dev(x).reshape(-1).expand(self.B).contiguous()
        self._f, self._DA, self._H = bc(f), bc(DA), bc(H)
        self.f = f
        self.DA, self.H, self.z, self.rdrag, self.h = DA, H, z, rdrag, h
        self._F = self._D = self._P22 = self._Cs = self._Cr = self._T = None
        self._bm = {}
        self._window = None  # FFTLog taper chosen by NonLinear.PsCf(bird, window=) (None: the NonLinear object's)
        self.snapshots = {}

    # ---- lazily produced device state ----
    def _bm_scalar(self, name):
        if name not in self._bm:
            src = getattr(self, "_" + name)
            if src is None:
                raise ValueError(f"Bird was constructed without {name}")
            self._bm[name] = self._plan().to_batch_minor(src)[0]
        return self._bm[name]

    def _plan(self):
        return self.co.device_plan(self._window)

    def _front(self):
        if self._F is None:
            self._F = self._plan().front(self.Pin)
        return self._F

    def _rows(self, name):
        a, n = self._plan().host.front.rows[name]
        return self._front()[a : a + n, : self.B]

    def _out(self, v):
        return v[0] if self._squeeze else v

    @property
    def P11(self):  # pybird.py:694-695
        return self._out(self._rows("P11").T)

    @property
    def P22(self):
        return self._out(self._P22[:, :, : self.B].permute(2, 0, 1))

    @property
    def P13(self):  # pybird.py:1080-1086
        t = self.torch
        k3 = t.as_tensor(self.co.k**3, device="cuda")
        raw = self._rows("P13raw").reshape(10, self.co.Nk, self.B)
        return self._out((raw * (k3[:, None] * self._rows("P11"))[None]).permute(2, 0, 1))

    @property
    def C11(self):
        return self._out(self._rows("C11").reshape(self.co.Nl, self.co.Nsr, self.B).permute(2, 0, 1))

    @property
    def Cct(self):
        return self._out(self._rows("Cct").reshape(self.co.Nl, self.co.Nsr, self.B).permute(2, 0, 1))

    @property
    def C22(self):
        return self._out(self._Cs[:, :28, :, : self.B].permute(3, 0, 1, 2))

    @property
    def C13(self):
        return self._out(self._Cs[:, 28:, :, : self.B].permute(3, 0, 1, 2))

    @property
    def Cloopl(self):
        return self._out(self._Cr[: self.B, :, 2:14, :])

    # ---- reference methods ----
    def setPsCfl(self):
        """Legendre weighting, f-grouping, stochastic basis, shot-noise subtraction (pybird.py:737-866)."""
        if self._P22 is None:
            raise RuntimeError("NonLinear.PsCf(bird) must run before setPsCfl()")
        dp = self._plan()
        self._T, self._Cr = dp.group(self._front(), self._P22, self._Cs, self._bm_scalar("f"), self.B)

    def create_snapshot(self, name):
        if name not in self.snapshots:
            self.snapshots[name] = BirdSnapshot(self)


class NonLinear:
    """pybird.py:870-1171.  Builds the loop kernels at construction (host, once); `PsCf` runs the
    front-end GEMM, the anti-diagonal kernel and the spectral GEMMs on the device."""

    def __init__(self, load=True, save=True, path="./", NFFT=256, co=common, name="pybird.nonlinear", kin=None,
                 window=0.2):
        import os

        self.co, self.NFFT, self.window = co, NFFT, window
        self.kin = np.logspace(-5, 0, 200) if kin is None else np.asarray(kin, float)
        self.fftsettings = dict(Nmax=NFFT, xmin=1.5e-5, xmax=1000.0, bias=-1.6)
        self.fft = FFTLog(**self.fftsettings)
        suffix = "_NNLO" if co.with_NNLO else ""
        eggpath = os.path.join(path, f"pyegg{NFFT}_Nl{co.Nl}{suffix}.npz")
        self._loop_cache = None
        if load:
            try:  # the reference's cache format (pybird.py:923-956): reuse M22/M13 when Pow matches
                L = np.load(eggpath)
                if not (self.fft.Pow - L["Pow"]).any():
                    self._loop_cache = (L["M22"], L["M13"])
                    save = False
            except Exception:
                pass
        if self._loop_cache is None:
            self._loop_cache = P.loop_matrices(self.fft)
        self.M22, self.M13 = self._loop_cache
        if save:
            try:
                np.savez(eggpath, Pow=self.fft.Pow, M22=self.M22, M13=self.M13)
            except Exception:
                pass
        co._register("nonlinear", self)

    def PsCf(self, bird: Bird, window=None):
        """pybird.py:1143-1171.  `window`: the FFTLog taper of this call (reference default 0.2; here the default is the value
        given to NonLinear(window=...)).  The taper is part of the front operator: every distinct value gets its own device
        plan, built the first time it is used."""
        window = self.window if window is None else window
        if window != (bird._window if bird._window is not None else self.window):
            bird._window, bird._F, bird._bm = window, None, {}
        dp = bird._plan()
        F = bird._front()
        bird._D = dp.antidiag(F, bird.B)
        # IRcutoff "loop"/"resum": coef_cf differs from coef_pk (pybird.py:1151-1160) -> a second anti-diagonal pass
        bird._Dcf = dp.antidiag(F, bird.B, cf_set=True) if dp.has_cf_set else None
        bird._P22, bird._Cs = dp.spectral(bird._D, bird.B, bird._Dcf)


class Resum:
    """pybird.py:1174-1464: full resummation, or the BAO-peak-only "optiresum" variant when `co.optiresum`."""

    def __init__(self, LambdaIR=0.2, NFFT=192, co=common, name="pybird.IRresum", snapshot=False):
        self.co, self.LambdaIR, self.NFFT, self.snapshot = co, float(LambdaIR), NFFT, snapshot
        self.sr = co.sr
        self.NIR = 16 if co.Nl == 3 else 8
        self.Na = 3 if self.NIR == 16 else 2
        self.Nn = 2 * self.NIR * self.Na
        co._register("resum", self)

    def Ps(self, bird: Bird, window=None):
        if bird._T is None:
            raise RuntimeError("bird.setPsCfl() must run before Resum.Ps")
        dp = bird._plan()
        dp.resum(bird._front(), bird._Cr, bird._bm_scalar("f"), bird._T, bird.B)
        if self.snapshot:
            bird.create_snapshot("IRresum")


class APeffect:
    """pybird.py:1467-1628."""

    def __init__(self, Om_AP=None, z_AP=None, DA=None, H=None, rdrag_AP=None, h_AP=None, nbinsmu=200, accboost=1,
                 Nlmax=None, APst=False, co=common, name="pybird.apeffect", snapshot=False):
        self.co, self.APst, self.snapshot = co, APst, snapshot
        if DA is not None and H is not None:
            self.DA, self.H = DA, H
        elif Om_AP is not None and z_AP is not None:
            self.DA, self.H = DAfunc(Om_AP, z_AP), Hubble(Om_AP, z_AP)
        else:
            raise ValueError("expect input params: Om_AP and z_AP, or DA and H")
        self.rdrag_AP, self.h_AP = rdrag_AP, h_AP
        self._nbinsmu, self._accboost = nbinsmu, accboost
        self.nbinsmu = accboost * nbinsmu
        self.Nlmax = Nlmax if Nlmax else co.Nl
        if self.Nlmax > co.Nl:
            raise ValueError(f"request Nlmax={self.Nlmax}, while bird only compute Nl up to {co.Nl}")
        if self.Nlmax != co.Nl:  # the reference itself cannot run this (pybird.py:1541 reads self.Nlmax before it is set)
            raise NotImplementedError("Nlmax < Nl is not supported (nor functional in the reference)")
        co._register("ap", self)

    def get_AP_param(self, bird):
        return bird._DA / self.DA, self.H / bird._H  # pybird.py:1560-1562

    def get_alperp_alpara(self, bird):
        if any(x is None for x in (self.rdrag_AP, self.h_AP, bird.rdrag, bird.h)):
            return self.get_AP_param(bird)
        t = bird.torch
        ratio = (self.rdrag_AP * self.h_AP) / (t.as_tensor(bird.rdrag, device="cuda") * t.as_tensor(bird.h, device="cuda"))
        return bird._DA / self.DA * ratio, self.H / bird._H * ratio  # pybird.py:1576-1578

    def AP(self, bird: Bird, q=None):
        dp = bird._plan()
        if q is not None:  # pybird.py:1603-1606: explicit (qperp, qpar), scalars or one pair per cosmology
            t = bird.torch
            dev = lambda x: (x if isinstance(x, t.Tensor) else t.as_tensor(np.asarray(x, float))).to("cuda", t.float64)
            qperp, qpar = (dev(v).reshape(-1).expand(bird.B).contiguous() for v in q)
            DA_bm, H_bm = dp.to_batch_minor(qperp * self.DA)[0], dp.to_batch_minor(self.H / qpar)[0]
        else:
            DA_bm, H_bm = bird._bm_scalar("DA"), bird._bm_scalar("H")
        bird._T = dp.ap(bird._T, DA_bm, H_bm, bird.B)
        if self.snapshot:
            bird.create_snapshot("APeffect")


# --------------------------------------------------------------------------------------------------
class FiberCollision:
    """pybird.py:1631-1809 - effective-window fibre-collision correction, same keywords.  `fibcolWindow(bird)` adds
    the correlated correction dPcorr to P11l, Pctl, Ploopl[, PctNNLOl] and, with `fiberst`, to Pstl (not to Picc);
    dPcorr is linear in the spectrum, so it is one fixed operator (plan.fiber_matrix) applied on the DMMA GEMM."""

    def __init__(self, fs, Dfc, ktrust=0.25, fiberst=False, co=None, name="pybird.fiber", snapshot=False):
        self.co = common if co is None else co
        self.fs, self.Dfc, self.ktrust, self.fiberst = fs, Dfc, ktrust, fiberst
        self.name, self.snapshot = name, snapshot
        self._matrix = None

    def matrix(self):
        """(Nl, Nk, Nl, Nk) operator F with dPcorr = F.P"""
        if self._matrix is None:
            self._matrix = P.fiber_matrix(self.co.k, self.co.Nl, self.fs, self.Dfc, self.ktrust)
        return self._matrix

    def fibcolWindow(self, bird):
        n = self.co.Nl * self.co.Nk
        op = np.eye(n) + self.matrix().reshape(n, n)
        apply_node_operator(bird, op, self.co.Nl, stochastic=self.fiberst, cache_owner=self)
        if self.snapshot:
            bird.create_snapshot("fiber")


class NodeOperator:
    """A fixed matrix on the (multipole, k-node) axis applied to every term row of a batch
    (`eftb_operator_*`): window, integral constraint, binning, chained mixing."""

    def __init__(self, matrix):
        self.lib = _lib.load()
        m = np.ascontiguousarray(matrix, dtype=np.float64)
        self.M, self.K = m.shape
        h = C.c_void_p()
        _lib.check(self.lib.eftb_operator_create(self.M, self.K, _lib.as_ptr(m), C.byref(h)), "eftb_operator_create")
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.eftb_operator_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def apply(self, T):
        """T: (Nl, Nk, nterm, Bp) -> (M, nterm, Bp) rows."""
        import torch

        Nl, Nk, nterm, Bp = T.shape
        assert Nl * Nk == self.K
        out = torch.empty((self.M, nterm, Bp), dtype=torch.float64, device="cuda")
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(self.lib.eftb_operator_apply(self.handle, C.c_void_p(T.data_ptr()), C.c_void_p(out.data_ptr()),
                                                nterm * Bp, s), "eftb_operator_apply")
        return out


def apply_node_operator(view: _TermsView, matrix, n_out_l, stochastic=True, cache_owner=None, in_place=True):
    """new_T[(a,k')] = matrix @ T[(l,k)] for every term row; returns the new (Na, nk', nterm, Bp) tensor."""
    op = getattr(cache_owner, "_node_operator", None) if cache_owner is not None else None
    if op is None:
        op = NodeOperator(matrix)
        if cache_owner is not None:
            cache_owner._node_operator = op
    T = view._T.contiguous()
    out = op.apply(T).reshape(n_out_l, -1, T.shape[2], T.shape[3])
    if not stochastic:
        if out.shape != T.shape:
            raise ValueError("stochastic terms can only be exempted when the operator preserves the node grid")
        out[:, :, 21:24] = T[:, :, 21:24]
    if in_place:
        view._T = out
    return out
