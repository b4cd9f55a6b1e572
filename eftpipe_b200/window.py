"""Survey window convolution - mirror of `eftpipe.window.Window` (window.py:40-415).

Precompute (host, once per plan; same cache files as the reference: `<window_fourier_file>.npy` + a
`.json` meta sidecar with strict equality check, window.py:204-260, :361-369):
    Wal[a, l, k, p] = p^2 Re sum_n Coef[a,l,k,n] p^{-eta_n - 3} M[l,n]      (window.py:262-346)
from the configuration-space multipoles Q_q(s) through a 4096-point FFTLog with a spherical-Bessel
kernel.  Per evaluation the reference resamples every term onto the p grid with a cubic spline and
contracts with `Waldk` (window.py:371-387); for the fixed internal nodes that is one fixed real matrix
on the Nl*Nk nodes (`effective_matrix`), which the CUDA path applies as a DMMA GEMM over the batch.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
from scipy.special import spherical_jn

from . import tables
from .fftlog import FFTLog
from .plan import window_effective_matrix, window_pgrid

window_kgrid = window_pgrid  # reference name (window.py:27)


class MetaInfoError(Exception):
    pass


# (2a+1) (a l q; 0 0 0)^2 coupling of window multipoles (window.py:286-303), indices a, l, q in units of 2
_CALQ = np.array([
    [[1, 0, 0, 0], [0, 1 / 5, 0, 0], [0, 0, 1 / 9, 0], [0, 0, 0, 1 / 13]],
    [[0, 1, 0, 0], [1, 2 / 7, 2 / 7, 0], [0, 2 / 7, 100 / 693, 25 / 143], [0, 0, 25 / 143, 14 / 143]],
    [[0, 0, 1, 0], [0, 18 / 35, 20 / 77, 45 / 143], [1, 20 / 77, 162 / 1001, 20 / 143],
     [0, 45 / 143, 20 / 143, 252 / 2431]],
    [[0, 0, 0, 1], [0, 0, 5 / 11, 14 / 55], [0, 5 / 11, 20 / 99, 28 / 187], [1, 14 / 55, 28 / 187, 400 / 3553]],
])


def _device_real_products(A_list, X):
    """[A @ X for A in A_list] on the GPU through the library's fixed-operator DMMA GEMM (`eftb_operator_*`):
    A (M, K) host arrays, X (K, N) host array; returns host arrays (M, N)."""
    import ctypes as C

    from . import _lib

    torch = _lib.require_cuda()
    lib = _lib.load()
    K, N = X.shape
    Np = (N + 31) // 32 * 32
    Xd = torch.zeros((K, Np), dtype=torch.float64, device="cuda")
    Xd[:, :N] = torch.as_tensor(X, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = []
    for A in A_list:
        A = np.ascontiguousarray(A, dtype=np.float64)
        h = C.c_void_p()
        _lib.check(lib.eftb_operator_create(A.shape[0], K, _lib.as_ptr(A), C.byref(h)), "eftb_operator_create")
        Cd = torch.empty((A.shape[0], Np), dtype=torch.float64, device="cuda")
        try:
            _lib.check(lib.eftb_operator_apply(h, C.c_void_p(Xd.data_ptr()), C.c_void_p(Cd.data_ptr()), Np, stream),
                       "eftb_operator_apply")
            out.append(Cd[:, :N].cpu().numpy())
        finally:
            lib.eftb_operator_destroy(h)
    return out


def compute_Wal(s_Q, k, Na, Nl, Nq=3, pmax=None, accboost=1, Nmax=4096, xmin_factor=1.0, xmax_factor=100.0,
                bias=-1.6, window_param=1, device=False):
    """Fourier-space window matrix (window.py:262-346).  s_Q: (ns, 1+nq) columns s, Q0, Q2, ...

    The dominant cost is the sum over the 4097 FFTLog frequencies, Re sum_n coef[l,k,n] M[l,n] p^{-eta_n-3}: a real
    product [Re(coef M) | -Im(coef M)] (Nl*Nk x 2N) times [Re pPow; Im pPow] (2N x Np).  With `device=True` it runs
    on the GPU through the library's DMMA GEMM, otherwise through the host BLAS."""
    k = np.asarray(k, float)
    pmax = float(k.max()) if pmax is None else pmax
    p = window_pgrid(pmax, accboost)
    tab = np.asarray(s_Q, float)
    while tab[0, 0] == 0.0:  # remove s = 0 (window.py:275-276)
        tab = tab[1:]
    tab = tab[:, : 1 + Nq]
    sw, Qq = tab[:, 0], tab[:, 1:].T
    Qal = np.einsum("alq,qs->als", _CALQ[..., :Nq], Qq)[:Na, :Nl]
    fft = FFTLog(Nmax=Nmax, xmin=sw[0] * xmin_factor, xmax=sw[-1] * xmax_factor, bias=bias)
    pPow = np.exp(np.outer(-fft.Pow - 3.0, np.log(p)))
    M = np.array([4 * np.pi * tables.bessel_power(2 * l, -0.5 * fft.Pow) for l in range(Nl)])
    X = np.concatenate([pPow.real, pPow.imag], axis=0)  # (2N, Np), shared by every (a, l)
    lhs = []
    for a in range(Na):  # one output multipole at a time keeps the (k, n) work arrays small
        kern = lambda x, a=a: spherical_jn(2 * a, x[None, None, :] * k[None, :, None])
        coef = fft.coef_host(sw, Qal[a][:, None, :], extrap="padding", window=window_param, kernel=kern)
        phase = ((-1j) ** (2 * a)) * ((1j) ** (2 * np.arange(Nl)))[:, None, None]
        cm = (phase * coef * M[:, None, :]).reshape(Nl * k.size, -1)  # (Nl*Nk, N): coef[l,k,n] M[l,n]
        lhs.append(np.concatenate([cm.real, -cm.imag], axis=1))
    prods = _device_real_products(lhs, X) if device else [A @ X for A in lhs]
    Wal = np.stack([pr.reshape(Nl, k.size, p.size) for pr in prods]) * p**2
    return Wal, p


class Window:
    """Same constructor keywords, cache files and errors as the reference class."""

    def __init__(self, window_fourier_file=None, window_configspace_file=None, co=None, load=True, save=True,
                 check_meta=True, Na=None, Nl=None, Nq=3, pmax=None, accboost=1, withmask=True, windowk=0.05,
                 Nmax=4096, xmin_factor=1.0, xmax_factor=100.0, bias=-1.6, window_param=1, window_st=True,
                 icc=None, name="pybird.window", snapshot=False, window_configspace_array=None, device=None):
        """`device`: build the Fourier-space matrix on the GPU (True), on the host (False), or on the GPU when one is
        present (None, default)."""
        from .pybird import common

        if device is None:
            try:
                import torch

                device = bool(torch.cuda.is_available())
            except Exception:
                device = False
        self._device_build = bool(device)

        self.co = co if co is not None else common
        if window_fourier_file is None and window_configspace_file is None and window_configspace_array is None:
            raise ValueError("Window requires window_fourier_file or window_configspace_file or both")
        self.window_fourier_file = Path(window_fourier_file).resolve() if window_fourier_file else None
        self.window_configspace_file = Path(window_configspace_file).resolve() if window_configspace_file else None
        self._array = window_configspace_array
        self._load = load
        self._save = save if self.window_fourier_file else False
        self.check_meta = check_meta
        self._create_meta = True
        self.window_st, self.withmask, self.windowk = window_st, withmask, windowk
        Na = Na if Na else self.co.Nl
        Nl = Nl if Nl else self.co.Nl
        if Na > self.co.Nl or Nl > self.co.Nl:
            raise ValueError(f"request Na={Na}, Nl={Nl} while bird only compute Nl up to {self.co.Nl}")
        if Na > Nl:
            raise ValueError(f"dangerous settings Na={Na}, Nl={Nl}")
        if pmax is None:
            pmax = float(self.co.k.max())
        self.p = window_pgrid(kmax=pmax, accboost=accboost)
        self.meta = dict(Na=Na, Nl=Nl, Nq=Nq, pmax=pmax, accboost=accboost, Nmax=Nmax, xmin_factor=xmin_factor,
                         xmax_factor=xmax_factor, bias=bias, window_param=window_param,
                         window_configspace_file=str(self.window_configspace_file) if self.window_configspace_file else None,
                         k=self.co.k.tolist())
        self.Wal = self._load_Wal()
        if self.Wal is None:
            self.Wal = self._compute_Wal()
        if self._save:
            self._save_Wal()
        self.icc = icc
        self.snapshot = snapshot
        self._op = None

    # ---- cache handling (window.py:204-260, :361-369) ----
    def _load_Wal(self):
        Wal, path = None, self.window_fourier_file
        if self._load and path is not None:
            try:
                Wal = np.load(path)
            except (OSError, TypeError):
                Wal = None
            else:
                if Wal.shape[1] != self.meta["Nl"]:
                    Wal = None
                    self.window_fourier_file = path = path.with_name(path.stem + f'_Nl{self.meta["Nl"]}.npy')
                    try:
                        Wal = np.load(path)
                    except OSError:
                        Wal = None
            if Wal is not None and self.check_meta:
                meta_file = path.with_suffix(".json")
                if not meta_file.exists():
                    self._create_meta = False
                else:
                    with meta_file.open("r") as fh:
                        meta = json.load(fh)
                    if self.meta["window_configspace_file"] is None:
                        self.meta["window_configspace_file"] = meta["window_configspace_file"]
                    if meta != self.meta:
                        raise MetaInfoError(f"inconsistent meta info\nloaded matrix's meta:\n{meta}\nexpect:\n{self.meta}")
        if Wal is not None:
            self._save = False
        return Wal

    def _compute_Wal(self):
        if self._array is not None:
            tab = np.asarray(self._array, float)
        else:
            if self.window_configspace_file is None:
                raise ValueError("please specify a configuration space mask file")
            try:
                tab = np.loadtxt(self.window_configspace_file)
            except OSError:
                raise OSError(f"Error: can't load mask file: {self.window_configspace_file}")
        m = self.meta
        Wal, _ = compute_Wal(tab, self.co.k, m["Na"], m["Nl"], Nq=m["Nq"], pmax=m["pmax"], accboost=m["accboost"],
                             Nmax=m["Nmax"], xmin_factor=m["xmin_factor"], xmax_factor=m["xmax_factor"], bias=m["bias"],
                             window_param=m["window_param"], device=self._device_build)
        return Wal

    def _save_Wal(self):
        # the reference saves on the MPI root only (window.py:361-369); here several ranks may build the same
        # window concurrently, so every file is written under a private name and renamed into place atomically
        import os

        final = self.window_fourier_file
        try:  # cache IO failures are not fatal (the reference swallows them with a warning, pybird.py:954-956)
            final.parent.mkdir(parents=True, exist_ok=True)
            tmp = final.with_name(f".{final.stem}.{os.getpid()}.tmp.npy")
            np.save(tmp, self.Wal)
            os.replace(tmp, final)
            if self._create_meta:
                mtmp = final.with_name(f".{final.stem}.{os.getpid()}.tmp.json")
                with mtmp.open("w") as fh:
                    json.dump(self.meta, fh, indent=2)
                os.replace(mtmp, final.with_suffix(".json"))
        except OSError as exc:
            import warnings

            warnings.warn(f"could not cache the window matrix at {final}: {exc}")

    # ---- operators ----
    @property
    def Waldk(self):
        """Masked, dp-weighted matrix (window.py:348-359)."""
        W = self.Wal
        if self.withmask:
            keep = (self.p[None, :] < self.co.k[:, None] + self.windowk) & (self.p[None, :] > self.co.k[:, None] - self.windowk)
            W = W * keep[None, None]
        return W * np.concatenate([[0.0], self.p[1:] - self.p[:-1]])

    def effective_matrix(self):
        """(Na, Nk, Nl, Nk): `integrWindow` on the internal nodes; ICC subtracted when attached
        (window.py:393-404)."""
        op = window_effective_matrix(self.Wal, self.p, self.co.k, windowk=self.windowk, withmask=self.withmask)
        if self.icc is not None:
            op = op - self.icc.effective_matrix()
        return op

    def Window(self, bird):
        """Apply the window (and ICC) to a batched Bird in place (window.py:389-415)."""
        from .pybird import apply_node_operator

        if self._op is None:
            op = self.effective_matrix()
            self._op = (np.ascontiguousarray(op.reshape(op.shape[0] * op.shape[1], -1)), op.shape[0])
        apply_node_operator(bird, self._op[0], self._op[1], stochastic=self.window_st, cache_owner=self)
        if self.icc is not None:
            bird.add_Picc(-self.icc.PSN)  # window.py:405
        if self.snapshot:
            bird.create_snapshot("window")


# --------------------------------------------------------------------------------------------------------------------
# Window *matrix* stage (window.py:418-577): a pre-tabulated mixing matrix between band powers instead of the
# configuration-space multipoles.  It interpolates every term onto a fixed band grid and multiplies by the matrix - on
# the fixed internal nodes again one fixed operator, here on (Nl * Nk) -> (n_ell_out * n_bin_out).
class PInfo:
    """multipoles / k range / number of bins of one side of a tabulated window matrix (window.py:418-423)"""

    def __init__(self, ells, kmin, kmax, nbins):
        self.ells, self.kmin, self.kmax, self.nbins = tuple(ells), float(kmin), float(kmax), int(nbins)

    def centres(self):
        edges = np.linspace(self.kmin, self.kmax, self.nbins + 1)
        return 0.5 * (edges[1:] + edges[:-1])


def to_window_matrix(matrix, inpoles: PInfo, outpoles: PInfo, ells_in, kmax_in, ells_out, kmin_out, kmax_out):
    """Cut the (multipole, band) blocks `ells_out` x `ells_in`, bands below `kmax_in` on the input side and in
    [kmin_out, kmax_out) on the output side, out of a flat (out, in) matrix and return them as
    (len(ells_out), len(ells_in), nk_out, nk_in) (window.py:426-469; like the reference it keeps the order of the
    multipoles of the file)."""
    matrix = np.asarray(matrix, float)

    def select(info, ells, lo, hi):
        c = info.centres()
        i0, i1 = np.searchsorted(c, lo) if lo is not None else 0, np.searchsorted(c, hi)
        keep = np.zeros(info.nbins * len(info.ells), dtype=bool)
        for j, ell in enumerate(info.ells):
            if ell in ells:
                keep[j * info.nbins + i0 : j * info.nbins + i1] = True
        return keep

    sub = matrix[np.ix_(select(outpoles, tuple(ells_out), kmin_out, kmax_out), select(inpoles, tuple(ells_in), None, kmax_in))]
    nout, nin = sub.shape[0] // len(ells_out), sub.shape[1] // len(ells_in)
    return sub.reshape(len(ells_out), nout, len(ells_in), nin).transpose(0, 2, 1, 3).copy()


class PolesInfo(tuple):
    """(nells, kstart, kend, nbin) (window.py:472-476)"""

    def __new__(cls, nells, kstart, kend, nbin):
        return super().__new__(cls, (nells, kstart, kend, nbin))

    nells = property(lambda self: self[0])
    kstart = property(lambda self: self[1])
    kend = property(lambda self: self[2])
    nbin = property(lambda self: self[3])


class WindowMatrix:
    """`eftpipe.window.WindowMatrix` (window.py:479-577): `matrix` (n_ell_out, n_ell_in, nbin_out, nbin_in); every term is
    interpolated (cubic, extrapolating) from `co.k` onto `kavg` and contracted with it.  Same keywords, the same
    validation errors, and the reference's hard-wired band grid (`kavg`, window.py:541-544)."""

    def __init__(self, matrix, inpoles, outpoles, co=None, window_st=False, icc=None, name="pybird.WindowMatrix", snapshot=False):
        from .pybird import common

        self.matrix = np.asarray(matrix, float)
        self.inpoles, self.outpoles = PolesInfo(*inpoles), PolesInfo(*outpoles)
        self.co = co if co is not None else common
        self.window_st, self.icc, self.name, self.snapshot = window_st, icc, name, snapshot
        if self.icc:
            raise NotImplementedError("ICC not implemented for WindowMatrix")
        if self.matrix.shape != (self.outpoles.nells, self.inpoles.nells, self.outpoles.nbin, self.inpoles.nbin):
            raise ValueError("matrix shape does not match meta information")
        if self.inpoles.nells != self.co.Nl:
            raise ValueError("input poles do not match self.co.Nl")
        self._op = None

    @classmethod
    def load(cls, path, ells, kmin, kmax, co=None, window_st=False, icc=None, name="pybird.WindowMatrix", snapshot=False):
        """window.py:507-538: a text matrix in the layout the reference hard-codes (in: l = 0, 2, 4 on 400 bands of
        [0, 0.4]; out: l = 0..4 on 40 bands of [0, 0.4])"""
        from .pybird import common

        co = co if co is not None else common
        m = to_window_matrix(np.loadtxt(path), PInfo((0, 2, 4), 0, 0.4, 400), PInfo((0, 1, 2, 3, 4), 0, 0.4, 40),
                             ells_in=tuple(2 * i for i in range(co.Nl)), kmax_in=co.k.max(), ells_out=tuple(ells),
                             kmin_out=kmin, kmax_out=kmax)
        return cls(m, PolesInfo(co.Nl, 0, co.k.max(), m.shape[3]), PolesInfo(len(ells), kmin, kmax, m.shape[2]), co=co,
                   window_st=window_st, icc=icc, name=name, snapshot=snapshot)

    @property
    def kavg(self):
        return np.linspace(0, 0.4, 400)[:300]  # window.py:541-544 (the reference's own hard-wired grid)

    def effective_matrix(self):
        """(n_ell_out, nbin_out, Nl, Nk): interpolation onto `kavg` composed with the band matrix"""
        from .plan import cubic_matrix

        if self.matrix.shape[3] != self.kavg.size:
            raise ValueError(f"matrix has {self.matrix.shape[3]} input bands, the band grid {self.kavg.size}")
        return np.einsum("alkp,pn->akln", self.matrix, cubic_matrix(self.co.k, self.kavg), optimize=True)

    def convolve(self, Plk):
        """host arrays (l, ..., Nk) -> (a, ..., nbin_out) (window.py:550-558)"""
        return np.einsum("akln,l...n->a...k", self.effective_matrix(), np.asarray(Plk, float), optimize=True)

    def Window(self, bird):
        """in place on a batched device Bird (one DMMA GEMM over the batch) or on a numpy BirdLike (window.py:560-577)"""
        from .pybird import _TermsView, apply_node_operator

        if isinstance(bird, _TermsView):
            if self._op is None:
                op = self.effective_matrix()
                self._op = (np.ascontiguousarray(op.reshape(op.shape[0] * op.shape[1], -1)), op.shape[0])
            apply_node_operator(bird, self._op[0], self._op[1], stochastic=self.window_st, cache_owner=self)
            bird.Picc = np.zeros((self.matrix.shape[0], self.matrix.shape[2]))  # window.py:573-575: convolved, then zeroed
        else:
            bird.P11l, bird.Pctl, bird.Ploopl = self.convolve(bird.P11l), self.convolve(bird.Pctl), self.convolve(bird.Ploopl)
            if bird.co.with_NNLO:
                bird.PctNNLOl = self.convolve(bird.PctNNLOl)
            if self.window_st:
                bird.Pstl = self.convolve(bird.Pstl)
            bird.Picc = np.zeros((self.matrix.shape[0], self.matrix.shape[2]))
        if self.snapshot:
            bird.create_snapshot("window")
