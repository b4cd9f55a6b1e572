"""Multi-tracer EFT likelihood - mirror of `eftpipe.likelihood.EFTLike` (likelihood.py:275-615) with its
helpers `parse_kmask`, `mask_covariance`, `hartlap`, `flatten`, `MultipoleInfo` (likelihood.py:78-272),
batched over points.  Data handling (files, masks, inverse covariance) stays on the host as in the
reference; the per-point work - bias reduction to PNG / PG (likelihood.py:483-549), chi^2 and the analytic
marginalisation (marginal.py) - runs in the CUDA kernels of csrc/like.cu through `engine.DeviceLikelihood`.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .marginal import LoggedError, Marginalizable, valid_prior_config
from .parambasis import NPAR, BirdComponent, TCT, TLOOP, TST, T11
from .transformer import f_batch_minor


# ------------------------------------------------------------------------------------ host helpers
def regularize_float_bound(x, n, default=None):
    if x is None:
        if default is None:
            raise TypeError("empty bound is not allowed if default is not provided")
        return [default] * n
    if isinstance(x, (float, int)):
        return [x] * n
    if len(x) != n:
        raise ValueError(f"expect len(x) = {n}, obtained {len(x)}")
    return list(x)


def parse_kmask(kall, ells, kmin, kmax):
    """likelihood.py:78-113"""
    try:
        kmins = regularize_float_bound(kmin, len(ells), -1)
        kmaxs = regularize_float_bound(kmax, len(ells), 1e10)
    except ValueError as ex:
        raise ValueError("length of kmin/kmax does not match ells") from ex
    ret = {}
    for ell, lo, hi in zip(sorted(ells), kmins, kmaxs):
        ret[ell] = slice(int(np.searchsorted(kall, lo)), int(np.searchsorted(kall, hi, side="right")))
    return ret


def slice_union(slices):
    slices = list(slices)
    return slice(min(s.start for s in slices), max(s.stop for s in slices))


def mask_covariance(cov, *args):
    """likelihood.py:122-160; *args = repeated (ls, ls_tot, kall, kmin, kmax)."""
    mask1d = np.array([], dtype=bool)
    rest = list(args)
    while rest:
        ls, ls_tot, kall, kmin, kmax, *rest = rest
        kmask = parse_kmask(kall, ls, kmin, kmax)
        mask = np.zeros((len(ls_tot), len(kall)), dtype=bool)
        for i, ell in enumerate(ls_tot):
            if ell in kmask:
                mask[i, kmask[ell]] = True
        mask1d = np.hstack((mask1d, mask.flatten()))
    if cov.shape[0] != mask1d.size:
        raise ValueError(f"covariance matrix's shape {cov.shape} does not match input data, "
                         f"expected {(mask1d.size, mask1d.size)}")
    return cov[np.outer(mask1d, mask1d)].reshape(mask1d.sum(), -1)


def hartlap(Nreal: int, ndata: int) -> float:
    return (Nreal - ndata - 2) / (Nreal - 1)


def flatten_rows(ls, nk, mask=None):
    """Row indices (into a (Nl_out*nk) multipole grid) selected by `flatten` (likelihood.py:167-195)."""
    rows = []
    for ell in ls:
        i = ell // 2
        sl = mask[ell] if mask else slice(0, nk)
        rows.extend(range(i * nk + sl.start, i * nk + sl.stop))
    return np.array(rows, dtype=np.int32)


def extract_multipole_info(names):
    """likelihood.py:48-62: ("P0", "P2", "k", ...) -> ("P", [0, 2])"""
    import re

    pattern = re.compile(r"^([A-Za-z]+)(\d+)$")
    symbols, ells = set(), []
    for x in names:
        if match := pattern.match(x):
            s, e = match.groups()
            symbols.add(s)
            ells.append(int(e))
    if len(symbols) == 0:
        raise ValueError("no valid multipole name found")
    if len(symbols) != 1:
        raise ValueError(f"ambiguous multipole names: {symbols}")
    return symbols.pop(), sorted(ells)


def find_reader_else_default(name, default, **kwargs):
    """reader.py:43-62: "default" / None, or a dotted path to `fn(path, **kwargs)`"""
    if (name or "default") == "default":
        return default
    import importlib

    module_name, callable_name = name.rsplit(".", 1)
    fn = getattr(importlib.import_module(module_name), callable_name)
    return lambda path: fn(path, **kwargs)


def read_pkl(path, **kwargs):
    """reader.py:29-40: a commented-header text table ("# k P0 P2 P4"); without a header the names are inferred.
    Returns (column names, 2-d array); column 0 is k."""
    names = None
    with open(path) as fh:
        for line in fh:
            if not line.startswith("#"):
                break
            toks = line.lstrip("#").split()
            if names is None and len(toks) >= 2 and all("=" not in t for t in toks):
                names = toks
    arr = np.atleast_2d(np.loadtxt(path, **kwargs))
    if names is None or len(names) != arr.shape[1]:
        names = ["k"] + [f"P{2 * i}" for i in range(arr.shape[1] - 1)]
    return names, arr


def _as_table(obj):
    """(names, array) from what a data reader returns: that pair, a pandas DataFrame (reader.py convention: first column
    is k), or a bare array"""
    if isinstance(obj, tuple):
        return list(obj[0]), np.asarray(obj[1], float)
    if hasattr(obj, "columns") and hasattr(obj, "to_numpy"):
        return [str(c) for c in obj.columns], obj.to_numpy(dtype=float)
    arr = np.asarray(obj, float)
    return ["k"] + [f"P{2 * i}" for i in range(arr.shape[1] - 1)], arr


@dataclass
class MultipoleInfo:
    """likelihood.py:225-272 for an in-memory table (k, then one column per multipole in `ls_tot`)."""

    symbol: str
    ls: list
    ls_tot: list
    kall: np.ndarray
    kmin: object
    kmax: object
    kmask: dict = field(repr=False, default=None)
    kout: np.ndarray = field(repr=False, default=None)
    kout_mask: dict = field(repr=False, default=None)
    data_vector: np.ndarray = field(repr=False, default=None)

    @classmethod
    def load(cls, table=None, ls=None, ls_tot=None, kmin=None, kmax=None, symbol=None, path=None, reader="default",
             reader_kwargs=None):
        """`path` / `reader` / `reader_kwargs` are the reference's yaml keys (likelihood.py:241-252): the multipole symbol
        and the available ells come from the column names; `table` takes an in-memory (k, multipoles...) array."""
        if path is not None or isinstance(table, (str, bytes)):
            names, table = _as_table(find_reader_else_default(reader, read_pkl, **(reader_kwargs or {}))(path or table))
            fsym, ftot = extract_multipole_info(names)
            cols = [names.index(f"{fsym}{ell}") for ell in ftot]
            table = np.column_stack([table[:, 0]] + [table[:, c] for c in cols])
            symbol, ls_tot = symbol or fsym, ls_tot or ftot
        table = np.asarray(table, float)
        symbol = symbol or "P"
        ls = [ls] if isinstance(ls, int) else list(ls)
        ls_tot = list(ls_tot) if ls_tot is not None else [2 * i for i in range(table.shape[1] - 1)]
        if missing := set(ls).difference(ls_tot):
            raise ValueError(f"ls {missing} not found in data")
        kall = table[:, 0]
        kmask = parse_kmask(kall, ls, kmin, kmax)
        data = np.hstack([table[:, 1 + ls_tot.index(ell)][kmask[ell]] for ell in ls])
        kout = kall[slice_union(kmask.values())]
        return cls(symbol=symbol, ls=ls, ls_tot=ls_tot, kall=kall, kmin=kmin, kmax=kmax, kmask=kmask, kout=kout,
                   kout_mask=parse_kmask(kout, ls, kmin, kmax), data_vector=data)


# ------------------------------------------------------------------------------------ device spec
def build_spec(tracers, data, invcov, gaussian=(), sigma_inv=None, mu=None, jeffreys=False):
    """Assemble the constant tables of `eftb_like_create`.

    tracers : list of dict(basis=, co=, nout=, nterm=, rows=int array of selected output rows,
              picc=(nout,) constant, parents=(iA, iB) or None)
    gaussian: ordered names of the marginalised parameters
    """
    nt = len(tracers)
    d_tracer = np.concatenate([np.full(len(t["rows"]), i, dtype=np.int32) for i, t in enumerate(tracers)])
    d_row = np.concatenate([np.asarray(t["rows"], dtype=np.int32) for t in tracers])
    # rows the marginalised-parameter derivatives read (un-binned interpolated products: a different operator)
    d_row_g = np.concatenate([np.asarray(t.get("rows_g", t["rows"]), dtype=np.int32) for t in tracers])
    picc = np.concatenate([np.asarray(t["picc"], float)[np.asarray(t["rows"])] for t in tracers])
    ndata = d_row.size
    data = np.asarray(data, float)
    if data.size != ndata or invcov.shape != (ndata, ndata):
        raise ValueError(f"data ({data.size}) / invcov {invcov.shape} do not match the {ndata} selected rows")
    ng = len(gaussian)
    g_count = np.zeros(max(ng, 1), dtype=np.int32)
    g_tracer = np.zeros((max(ng, 1), 2), dtype=np.int32)
    g_term = np.zeros((max(ng, 1), 2, 3), dtype=np.int32)
    g_var = np.zeros((max(ng, 1), 2, 3), dtype=np.int32)
    g_coef = np.zeros((max(ng, 1), 2, 3))
    mode = np.array([int(getattr(t["basis"], "kernel_mode", "") == "explicit") for t in tracers], dtype=np.int32)
    xb_off = np.zeros(nt, dtype=np.int32)
    xg_off = np.full((max(ng, 1), nt), -1, dtype=np.int32)
    owned = np.zeros(max(ng, 1), dtype=bool)
    ncol = NPAR * nt  # explicit blocks follow the built-in columns
    for it, t in enumerate(tracers):
        if mode[it]:  # custom basis: nterm bias columns + nterm columns per Gaussian parameter it knows
            xb_off[it] = ncol
            ncol += t["nterm"]
            for g, name in enumerate(gaussian):
                if name in t["basis"].gaussian_params():
                    xg_off[g, it] = ncol
                    ncol += t["nterm"]
                    owned[g] = True
            continue
        desc = t["basis"].gaussian_descriptors(t["co"])
        for name, entries in desc.items():
            if name not in gaussian:
                continue
            g = list(gaussian).index(name)
            e = g_count[g]
            if e >= 2:
                raise NotImplementedError(f"gaussian parameter {name} enters more than two tracers")
            g_tracer[g, e] = it
            for q, (term, var, coef) in enumerate(entries):
                g_term[g, e, q], g_var[g, e, q], g_coef[g, e, q] = term, var, coef
            g_count[g] = e + 1
            owned[g] = True
    if ng and not owned[:ng].all():
        missing = [n for n, c in zip(gaussian, owned) if not c]
        raise LoggedError(f"marginalised parameters {missing} do not belong to any tracer")
    sigma_inv = np.zeros((ng, ng)) if sigma_inv is None else np.asarray(sigma_inv, float)
    mu = np.zeros(ng) if mu is None else np.asarray(mu, float)
    scales = np.array([[t["co"].kmA, t["co"].krA, t["co"].ndA, t["co"].kmB, t["co"].krB, t["co"].ndB] for t in tracers])
    return dict(
        ntracer=nt, ndata=ndata, ngauss=ng, npar=ncol, jeffreys=bool(jeffreys), mode=mode, xb_off=xb_off, xg_off=xg_off,
        gaussian=list(gaussian), tracer_nterm=[t["nterm"] for t in tracers], tracer_co=[t["co"] for t in tracers],
        nout=np.array([t["nout"] for t in tracers], dtype=np.int32),
        nterm=np.array([t["nterm"] for t in tracers], dtype=np.int32), scales=scales,
        par_index=np.arange(NPAR * nt, dtype=np.int32).reshape(nt, NPAR),
        eastcoast=np.array([int(t["basis"].counterform() == "eastcoast") for t in tracers], dtype=np.int32),
        d_tracer=d_tracer, d_row=d_row, d_row_g=d_row_g, data=data, picc=picc, invcov=np.ascontiguousarray(invcov, float),
        g_count=g_count, g_tracer=g_tracer, g_term=g_term, g_var=g_var, g_coef=g_coef,
        sigma_inv=sigma_inv, sigma_inv_mu=sigma_inv @ mu, mu_sigma_mu=float(mu @ sigma_inv @ mu),
    )


def pack_nuisance(torch, bases, params, f_list, B, Bp, spec=None):
    """(npar, Bp) batch-minor nuisance array from a parameter dictionary (scalars or (B,) arrays).
    Parameters missing from `params` are 0, as in the reference (`basis.default()`, parambasis.py:234-236).
    `spec` (build_spec): needed when a tracer uses a custom basis (explicit bias columns after the built-in ones)."""
    npar = NPAR * len(bases) if spec is None else int(spec["npar"])
    nuis = torch.zeros((npar, Bp), dtype=torch.float64, device="cuda")
    conv = {}

    def dev(v):
        if isinstance(v, (int, float)):
            return float(v)
        key = id(v)
        if key not in conv:
            t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v, float))
            conv[key] = t.to("cuda", torch.float64).reshape(-1)
        return conv[key]

    pdev = {k: dev(v) for k, v in params.items()}
    for it, basis in enumerate(bases):
        if getattr(basis, "kernel_mode", "") == "explicit":
            if spec is None:
                raise ValueError("a custom basis needs the likelihood spec to place its columns")
            bias, table = basis.explicit_columns(params, f_list[it][:B], B, spec["tracer_co"][it], spec["tracer_nterm"][it], spec["gaussian"])
            put = lambda off, arr: nuis[off : off + arr.shape[1]].__setitem__((slice(None), slice(0, B)), torch.as_tensor(arr.T.copy(), device="cuda"))
            put(int(spec["xb_off"][it]), bias)
            for g, name in enumerate(spec["gaussian"]):
                off = int(spec["xg_off"][g, it])
                if off >= 0:
                    put(off, table[name])
            if Bp > B:
                nuis[:, B:] = nuis[:, B - 1 : B]
            continue
        cols = basis.kernel_columns(pdev, f_list[it][:B] if f_list[it] is not None else None)
        for i, v in enumerate(cols):
            if isinstance(v, float):
                if v != 0.0:
                    nuis[NPAR * it + i].fill_(v)
            else:
                nuis[NPAR * it + i, :B] = v
                if Bp > B:
                    nuis[NPAR * it + i, B:] = v[-1]
    return nuis


def _probe_co(co):
    from types import SimpleNamespace

    return SimpleNamespace(**{k: getattr(co, k) for k in ("kmA", "krA", "ndA", "kmB", "krB", "ndB", "counterform", "with_NNLO")}, No=1)


def reduce_on_device(basis, bird, params, want_table=False):
    """`basis.reduce_Plk(bird, params)` and `reduce_Plk_gaussian_table` for one birdlike
    (parambasis.py:231-316): returns (BirdComponent-like sum as (B, No, nk) tensor, table dict)."""
    from .engine import DeviceLikelihood

    import torch

    co = bird.co
    Nl_out, nk, nterm, Bp = bird._T.shape
    No = min(co.No, Nl_out)
    rows = np.arange(No * nk, dtype=np.int32)
    picc = np.zeros(Nl_out * nk) if bird._picc is None else np.asarray(bird._picc, float).reshape(-1)
    explicit = getattr(basis, "kernel_mode", "") == "explicit"
    if not want_table:
        gaussian = []
    elif explicit:  # what the basis' own table holds for this configuration (probed once on a unit bird)
        from .parambasis import _UnitBird

        probe = basis.inner.reduce_Plk_gaussian_table(_UnitBird(_probe_co(co), 0.5, nterm), {n: 1.0 for n in
                                                      list(basis.inner.gaussian_params()) + list(basis.inner.non_gaussian_params())})
        gaussian = list(probe)
    else:
        gaussian = [n for n in basis.gaussian_descriptors(co)]
    key = (type(basis).__name__, basis.prefix, tuple(basis.cross_prefix), Nl_out, nk, want_table, id(co))
    cache = bird.__dict__.setdefault("_reduce_cache", {})
    if key not in cache:
        spec = build_spec([dict(basis=basis, co=co, nout=Nl_out * nk, nterm=nterm, rows=rows, picc=picc)],
                          data=np.zeros(rows.size), invcov=np.eye(rows.size), gaussian=gaussian)
        cache[key] = DeviceLikelihood(spec)
    like = cache[key]
    f_bm = f_batch_minor(bird)
    nuis = pack_nuisance(torch, [basis], params, [f_bm], bird.B, Bp, spec=like.spec)
    vec = like.vectors(bird.B, [bird._T.contiguous()], [f_bm], nuis)  # (B, ndata, 1+ng)
    total = vec[:, :, 0].reshape(bird.B, No, nk)
    table = {name: vec[:, :, 1 + i].reshape(bird.B, No, nk) for i, name in enumerate(gaussian)}
    if bird._squeeze:
        total = total[0]
        table = {k: v[0] for k, v in table.items()}
    zero = torch.zeros_like(total)
    comp = BirdComponent(Plin=total, Ploop=zero, Pct=zero, Pst=zero, Picc=zero)  # kernel returns the sum
    return comp, table


# ------------------------------------------------------------------------------------ EFTLike
class EFTLike(Marginalizable):
    """Batched counterpart of likelihood.py:275-615.  Construction keywords follow the reference's yaml
    keys (`tracers, data, cov, chained, with_binning, binning, marg, jeffreys`); instead of a Cobaya provider
    the theory object (`theory.EFTLSS`) is passed to `initialize_with_provider` and must offer
    `get_nonlinear_Plk_terms(tracer, chained, binned)` (batch-minor terms) and `.bases`.

    data[tracer] = dict(table=<array or path: k, P_l columns>, ls=[...], kmin=, kmax=[, ls_tot=])
    cov          = dict(matrix=<array or path>, Nreal=None, rescale=1)
    """

    def __init__(self, tracers, data, cov, chained=False, with_binning=True, binning=None, marg=None, jeffreys=False,
                 likelihood_prefix=None, marg_param_prefix="marg_", with_interp=True):
        self.tracers = [tracers] if isinstance(tracers, str) else list(tracers)
        n = len(self.tracers)
        as_dict = lambda x: x if isinstance(x, dict) and set(x) == set(self.tracers) else (
            {t: x[i] for i, t in enumerate(self.tracers)} if isinstance(x, list) else {t: x for t in self.tracers})
        self.data = data if set(data) == set(self.tracers) else {self.tracers[0]: data}
        self.chained, self.with_binning = as_dict(chained), as_dict(with_binning)
        self.with_interp = as_dict(with_interp)
        self.binning = as_dict(binning or {})
        self.cov = cov if isinstance(cov, dict) else ({"path": cov} if isinstance(cov, (str, bytes, list)) else {"matrix": cov})
        self.marg, self.jeffreys = marg or {}, jeffreys
        self.likelihood_prefix = likelihood_prefix or "eftlike_"
        self.marg_param_prefix = marg_param_prefix
        self.initialize()

    def initialize(self):
        """likelihood.py:293-307, :337-363"""
        self.minfodict = {t: MultipoleInfo.load(**self.data[t]) for t in self.tracers}
        self.data_vector = np.hstack([m.data_vector for m in self.minfodict.values()])
        self.ndata = self.data_vector.size
        for t, m in self.minfodict.items():
            self.binning[t] = dict(self.binning.get(t) or {}, kout=m.kout)
        if "path" in self.cov:  # likelihood.py:337-347: reader by dotted path, a list of paths is block-diagonal
            import scipy.linalg

            rd = find_reader_else_default(self.cov.get("reader"), lambda q: np.loadtxt(q, **self.cov.get("reader_kwargs", {})),
                                          **self.cov.get("reader_kwargs", {}))
            path = self.cov["path"]
            cov = scipy.linalg.block_diag(*[np.asarray(rd(q), float) for q in path]) if isinstance(path, list) else np.array(rd(path), float)
        else:
            mat = self.cov["matrix"]
            cov = np.loadtxt(mat) if isinstance(mat, (str, bytes)) else np.array(mat, float)
        cov = cov / self.cov.get("rescale", 1)
        self.hartlap = None
        if (Nreal := self.cov.get("Nreal")) is not None:
            self.hartlap = hartlap(Nreal, self.ndata)
            cov = cov / self.hartlap
        self.full_covmat = cov
        args = ()
        for m in self.minfodict.values():
            args += (m.ls, m.ls_tot, m.kall, m.kmin, m.kmax)
        self.invcov = np.linalg.inv(mask_covariance(cov, *args))

    def get_requirements(self):
        """likelihood.py:386-432: binned grid, interpolator (with the evaluation points, see EFTLSS.must_provide) or
        the raw un-binned grid per tracer."""
        reqs = {"nonlinear_Plk_grid": {}, "nonlinear_Plk_interpolator": {}, "nonlinear_Plk_gaussian_grid": {}}
        for t, m in self.minfodict.items():
            if self.with_binning[t]:
                req = {"ls": m.ls, "chained": self.chained[t], "binned": True, "binning": self.binning[t]}
                reqs["nonlinear_Plk_grid"][t] = req
            elif self.with_interp[t]:
                req = {"ls": m.ls, "chained": self.chained[t]}
                reqs["nonlinear_Plk_interpolator"][t] = dict(req, kout=m.kout)
                req = dict(req, binned=False)
            else:
                req = {"ls": m.ls, "chained": self.chained[t], "binned": False}
                reqs["nonlinear_Plk_grid"][t] = req
            if self.marg:
                reqs["nonlinear_Plk_gaussian_grid"][t] = req
        return {k: v for k, v in reqs.items() if v}

    def _product_key(self, t):
        """which of the tracer's products this likelihood reads (likelihood.py:503-516, :535-547)"""
        binned = bool(self.with_binning[t])
        return dict(chained=bool(self.chained[t]), binned=binned, interp=not binned and bool(self.with_interp[t]))

    def marginalizable_params(self):
        params = []
        for b in self.eft_bases:
            params += b.gaussian_params()
        return list(dict.fromkeys(params))

    def update_prior(self, prior):
        flat = {}
        for p, config in prior.items():  # prefix form, likelihood.py:198-222
            if valid_prior_config(config):
                flat[p] = config
            elif isinstance(config, dict):
                for name, sub in config.items():
                    flat[f"{p}{name}"] = sub
            else:
                raise ValueError(f"invalid prior config: {config}")
        return super().update_prior(flat)

    def initialize_with_provider(self, theory):
        """likelihood.py:434-473: bases, priors, and here also the device tables."""
        from .engine import DeviceLikelihood

        self.provider = theory
        self.eft_bases = [theory.bases[t] for t in self.tracers]
        gaussian = []
        if self.marg:
            self.setup_prior(self.marg)
            gaussian = list(self.valid_prior)
        elif self.jeffreys:
            raise NotImplementedError
        specs = []
        for t in self.tracers:
            m = self.minfodict[t]
            info = theory.product_info(t, **self._product_key(t))
            nk, rows_g = info["nk"], None
            if self.with_binning[t]:
                if nk != m.kout.size or info.get("interp_nk") is not None:
                    raise ValueError(f"{t}: theory bins ({nk}) do not match the data k-range ({m.kout.size})")
                rows = flatten_rows(m.ls, nk, m.kout_mask)
            elif self.with_interp[t]:  # likelihood.py:510-513, :541-544
                if info.get("interp_nk") != m.kout.size:
                    raise ValueError(f"{t}: the theory was not asked for the interpolator at this tracer's {m.kout.size} data points")
                rows = flatten_rows(m.ls, nk, m.kout_mask)  # first half of each multipole's rows: PlkInterpolator
                rows_g = rows + m.kout.size                 # second half: plain cubic interpolation
            else:  # likelihood.py:515-516, :546-547: every node of the internal grid, no mask
                if info.get("interp_nk") is not None:
                    raise ValueError(f"{t}: raw-grid likelihood on a plan built for interpolated products")
                rows = flatten_rows(m.ls, nk, None)
            spec = dict(basis=theory.bases[t], co=theory.commons[t], nout=info["nout"], nterm=info["nterm"],
                        rows=rows, picc=info["picc"])
            if rows_g is not None:
                spec["rows_g"] = rows_g
            specs.append(spec)
        # callable loc / scale (marginal.py:13-20): the prior is per point, evaluated in `calculate`; the plan holds none
        self.callable_prior = bool(gaussian) and self.has_callable_prior()
        sig = self.sigma_inv if gaussian and not self.callable_prior else None
        mu = self.mu_G if gaussian and not self.callable_prior else None
        self.spec = build_spec(specs, self.data_vector, self.invcov, gaussian=gaussian, sigma_inv=sig, mu=mu,
                               jeffreys=self.jeffreys)
        self.device = DeviceLikelihood(self.spec)
        self.gaussian_names = gaussian

    # ---- per batch ----
    def _inputs(self, params):
        import torch

        th = self.provider
        terms, fs = [], []
        for t in self.tracers:
            bm, f_bm = th.get_nonlinear_Plk_terms(t, **self._product_key(t))
            terms.append(bm)
            fs.append(f_bm)
        B = th.B
        Bp = terms[0].shape[-1]
        nuis = pack_nuisance(torch, self.eft_bases, params, fs, B, Bp, spec=self.spec)
        return B, terms, fs, nuis

    def PNG_PG(self, params):
        """(PNG (B, ndata), PG (B, nG, ndata)) - likelihood.py:483-549."""
        B, terms, fs, nuis = self._inputs(params)
        vec = self.device.vectors(B, terms, fs, nuis)
        import torch

        png = vec[:, :, 0] + torch.as_tensor(self.data_vector, device="cuda")
        return png, vec[:, :, 1:].permute(0, 2, 1)

    def env(self, params):
        """likelihood.py:560-564: numpy plus every tracer's EFT parameters - here arrays over the batch"""
        out = {"np": np}
        for t in self.tracers:
            for name, val in self.provider.get_eft_params_values_dict(t, params).items():
                out[name] = val.detach().cpu().numpy() if hasattr(val, "detach") else np.asarray(val, dtype=np.float64)
        return out

    def calculate(self, params, want_bestfit=False):
        """likelihood.py:570-594 for a batch: returns dict(logp=(B,), chi2=(B,), status=(B,) [, bestfit])."""
        B, terms, fs, nuis = self._inputs(params)
        loc = sinv = None
        if self.callable_prior:
            import torch

            loc, sinv = (torch.as_tensor(a, device="cuda") for a in self.point_priors(self.env(params), B))
        logp, status, best, full = self.device.eval(B, terms, fs, nuis, want_bestfit=want_bestfit, want_fullchi2=True,
                                                    prior_loc=loc, prior_sigma_inv=sinv)
        out = {"logp": logp, self.likelihood_prefix + "chi2": -2.0 * logp, "status": status}
        if self.gaussian_names:  # likelihood.py:583-590: chi^2 at the best-fit marginalised parameters
            out[self.likelihood_prefix + "fullchi2"] = full
        if want_bestfit:
            out["bestfit"] = {self.marg_param_prefix + n: best[:, i] for i, n in enumerate(self.gaussian_names)}
        return out

    def logp(self, params):
        return self.calculate(params)["logp"]
