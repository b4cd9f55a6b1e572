#!/usr/bin/env python
"""Benchmark of the PyBird one-loop multipole + likelihood hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): single LRG tracer, Nl=3, NFFT=256: FFTLog -> one-loop P22/P13 + config-space
images -> IR resummation -> AP -> DR16 NGC LRG window (accboost 4, windowk 0.1) + synthetic integral constraint
-> 18 k-bins -> bias reduction -> analytically marginalised likelihood (6 Gaussian parameters, Jeffreys) against
the DR16 NGC LRG data vector (54 points); batch of 1024 synthetic linear spectra PER GPU (weak scaling: the
batch shards with no data-path collective, only the per-point log-likelihood is gathered).

One "step" = one pass of the whole path over one batch.  `value` = evaluations/s with inputs resident in HBM;
`e2e` = the same through the host-facing call with pinned HOST buffers (H2D of P_lin/f/DA/H/nuisance and D2H
of logp + multipoles inside the timed region).  The `--impl reference` arm times the CPU restatement of the
reference path (oracle/, one process per host core) on the same workload; it is also what `cpu_baseline`
reports (rank 0, N=1, bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "P_l(k) one-loop multipole+likelihood evals/sec"
UNIT = "evaluations/s"
Z_EFF, Z_AP = 0.7, 0.696
GAUSS = ["b3", "cct", "cr1", "cr2", "ce0", "cequad"]
WORKLOAD = ("config2: single LRG Nl=3 NFFT=256 + IRresum + AP(APst) + DR16 NGC LRG window(accboost4,windowk0.1) + synthetic ICC"
            " + 18 bins + marginalised likelihood (6 Gaussian params, Jeffreys, 54 data points)")


# ------------------------------------------------------------------------------------------ setup (host)
def load_fixture():
    return dict(np.load(os.path.join(ROOT, "eftpipe_b200", "data", "dr16_ngc.npz")))


def window_cache_path():
    d = os.path.join(ROOT, "gpurun_out", "cache")
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, "win_NGC_LRG_acc4.npy")


def host_setup(B, seed=20261018 + 2, device=None):
    """Everything cosmology independent + the synthetic inputs (excluded from all timings)."""
    from eftpipe_b200 import likelihood, pybird, synthetic, window

    fx = load_fixture()
    co = pybird.Common(Nl=3, No=3, kmax=0.3, kmA=0.7, krA=0.25, ndA=4.5e-5)
    t0 = time.time()
    win = window.Window(window_fourier_file=window_cache_path(), window_configspace_array=fx["win_LRG"], co=co,
                        accboost=4, windowk=0.1, device=device)  # device=False: the reference arm never touches the GPU
    t_window = time.time() - t0
    Pshot = 1.0 / 4.5e-5
    PSN = 1e-3 / co.k[None, :] * np.array([1.0, 0.3, 0.1])[:, None]  # synthetic ICC
    minfo = likelihood.MultipoleInfo.load(fx["NGC_LRG_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20)
    cov = fx["cov_NGC_L024_P"] / likelihood.hartlap(1000, minfo.data_vector.size)
    invcov = np.linalg.inv(likelihood.mask_covariance(cov, [0, 2, 4], [0, 2, 4], minfo.kall, 0.02, 0.20))
    batch = synthetic.make_batch(B, Z_EFF, seed=seed, unique=min(B, 32))
    nuis = synthetic.draw_nuisance(B, seed=seed)
    return dict(fx=fx, co=co, win=win, Pshot=Pshot, PSN=PSN, minfo=minfo, invcov=invcov, batch=batch, nuis=nuis,
                t_window=t_window)


def kernel_columns(nuis):
    """(B, 17) west-coast kernel inputs; Gaussian (marginalised) parameters are zero in PNG."""
    from eftpipe_b200 import synthetic

    b1, c2, _, c4 = nuis[:, 0], nuis[:, 1], nuis[:, 2], nuis[:, 3]
    b2, b4 = synthetic.c2c4_to_b2b4(c2, c4)
    cols = np.zeros((nuis.shape[0], 17))
    cols[:, 0], cols[:, 1], cols[:, 3] = b1, b2, b4
    cols[:, 7:14] = cols[:, 0:7]
    return cols


# ------------------------------------------------------------------------------------------ CPU arm
_W = {}


def _cpu_worker_limit():
    """pool initializer: one evaluation per core, BLAS threading off inside the worker"""
    from threadpoolctl import threadpool_limits

    _W["_limit"] = threadpool_limits(limits=1)


def _cpu_prepare(wal_path, common_kw, kout, invcov, data, PSN_Pshot):
    """oracle objects, built once in the parent (forked workers share the pages)"""
    import warnings

    warnings.filterwarnings("ignore")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pybird_oracle as orc

    co = orc.Common(**common_kw)
    Wal = np.load(wal_path)
    p = orc.window_pgrid(0.3, 4)
    _W.update(orc=orc, co=co, nl=orc.NonLinear(co), rs=orc.Resum(co),
              ap=orc.APeffect(co, Om_AP=0.307115, z_AP=Z_AP, APst=True),
              Waldk=orc.mask_and_measure(Wal, p, co.k, 0.1), Waldk_ic=orc.mask_and_measure(0.05 * Wal, p, co.k, 0.1), p=p,
              binning=orc.Binning(kout, co), invcov=invcov, data=data, PSN_Pshot=PSN_Pshot)


def _cpu_eval(args):
    """One full evaluation of the reference formulation (oracle/pybird_oracle.py)."""
    kin, plin, f, DA, H, cols = args
    W = _W
    orc, co = W["orc"], W["co"]
    b = orc.Bird(co, kin, plin, f, DA, H, Z_EFF)
    W["nl"].PsCf(b)
    orc.set_PsCfl(b)
    W["rs"].Ps(b)
    W["ap"].AP(b)
    orc.apply_window(b, W["Waldk"], W["p"], window_st=True, icc=(W["Waldk_ic"], W["PSN_Pshot"]))
    terms = W["binning"].transform(orc.bird_terms(b))
    bsA = list(cols[0:7])
    PNG = orc.reduce_Plk(co, f, terms, bsA).reshape(-1)
    tab = orc.gaussian_table_west(co, f, terms, bsA[0])
    PG = np.array([tab[n].reshape(-1) for n in GAUSS])
    return orc.marginalized_logp(PNG, PG, W["data"], W["invcov"], jeffreys=True)


def cpu_arm(S, nproc, npoints, repeats=1):
    """Time `npoints` evaluations `repeats` times on `nproc` worker processes; returns (evals/s list, logp)."""
    import multiprocessing as mp

    b = S["batch"]
    cols = kernel_columns(S["nuis"])
    work = [(b.kin, b.plin[i % len(b)], b.f[i % len(b)], b.DA[i % len(b)], b.H[i % len(b)], cols[i % len(b)]) for i in range(npoints)]
    init = (window_cache_path(), dict(Nl=3, No=3, kmax=0.3, kmA=0.7, krA=0.25, ndA=4.5e-5), S["minfo"].kout, S["invcov"],
            S["minfo"].data_vector, S["PSN"] * S["Pshot"])
    rates, out = [], None
    _cpu_prepare(*init)
    if nproc == 1:
        for _ in range(repeats):
            t0 = time.perf_counter()
            out = [_cpu_eval(w) for w in work]
            rates.append(npoints / (time.perf_counter() - t0))
        return rates, np.array(out)
    ctx = mp.get_context("fork")
    with ctx.Pool(nproc, initializer=_cpu_worker_limit) as pool:
        pool.map(_cpu_eval, work[:nproc])  # touch every worker once (imports, caches)
        for _ in range(repeats):
            t0 = time.perf_counter()
            out = pool.map(_cpu_eval, work, chunksize=max(1, npoints // (nproc * 2)))
            rates.append(npoints / (time.perf_counter() - t0))
    return rates, np.array(out)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    S = host_setup(args.batch, device=False)
    cores = os.cpu_count() or 1
    nproc = max(1, min(cores, 96))
    per_step = nproc * 2
    rates, _ = cpu_arm(S, nproc, per_step, repeats=args.warmup + args.steps)
    timed = rates[args.warmup:]
    value = float(len(timed) / sum(1.0 / r for r in timed))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * per_step / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch, "sample_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": "port",
                         "sample": f"{per_step} evaluations per step on {nproc} single-threaded worker processes "
                                   f"(oracle/pybird_oracle.py restatement of the reference numpy path; host has {cores} cores)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(p[0]))
                smax = float(p[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from eftpipe_b200 import _lib, likelihood, parambasis, plan as P
    from eftpipe_b200.engine import DeviceLikelihood, DevicePlan

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    if world > 1 and rank != 0:
        dist.barrier()  # rank 0 builds (and caches) the window matrix first; the others load it
    S = host_setup(B, seed=20261018 + 2 + 1000 * rank)  # every rank owns a different shard of the point set
    if world > 1 and rank == 0:
        dist.barrier()
    co, win = S["co"], S["win"]
    g = P.GridConfig(Nl=3)
    binm, keff, _, _ = P.binning_matrix(g.k, S["minfo"].kout)
    Weff = P.window_effective_matrix(win.Wal, win.p, g.k, windowk=0.1)
    proj = P.compose_projection(g, window=Weff, icc=dict(matrix=0.05 * Weff, PSN_times_Pshot=S["PSN"] * S["Pshot"]), binning=binm)
    t0 = time.time()
    host_plan = P.build_tracer_plan(Nl=3, ap=dict(DA=P_DA(), H=P_H(), APst=True), projection=proj)
    t_plan = time.time() - t0
    dp = DevicePlan(host_plan)
    basis = parambasis.WestCoastBasis(prefix="")
    nk = S["minfo"].kout.size
    spec = likelihood.build_spec([dict(basis=basis, co=co, nout=3 * nk, nterm=24, rows=np.arange(3 * nk, dtype=np.int32),
                                       picc=host_plan.picc_out)], S["minfo"].data_vector, S["invcov"], gaussian=GAUSS, jeffreys=True)
    like = DeviceLikelihood(spec)
    lib = _lib.load()

    b = S["batch"]
    cols = kernel_columns(S["nuis"])
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    d_plin, d_f, d_DA, d_H, d_cols = dev(b.plin), dev(b.f), dev(b.DA), dev(b.H), dev(cols)
    Bp = dp.padded(B)
    terms_bm = torch.empty((3 * nk, 24, Bp), dtype=torch.float64, device="cuda")
    launches = {"n": 0}

    def step(plin, f, DA, H, cc):
        """one pass of the hot path over one batch; returns (logp, multipoles PNG-data)"""
        dp.eval_terms(plin, f, DA, H, want_bm=True, want_pm=False, out_bm=terms_bm)
        f_bm = dp.to_batch_minor(f)[0]
        nuis_bm = dp.to_batch_minor(cc)
        logp, status, _ = like.eval(B, [terms_bm], [f_bm], nuis_bm)
        return logp, status

    # kernels launched per step (counted from the call graph of csrc/api.cu + like.cu):
    # front: transpose+tails+gemm (3) | f,DA,H to batch-minor (1) | antidiag (1) | spectral: regroup + 2 gemms (3) | group (1)
    # | resum: Q(f) + sweep (2) | ap: gemm + geom + apply (3) | project gemm (1) | f, nuisance transposes (2)
    # | like: vectors+gemm+finish (3)
    launches["n"] = 20

    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")  # > 126 MB L2
    for _ in range(args.warmup):
        step(d_plin, d_f, d_DA, d_H, d_cols)
    torch.cuda.synchronize()
    # one CUDA graph per step: the 22 launches are replayed with a single launch call
    graph = None
    if not args.no_graph:
        from eftpipe_b200.engine import capture_graph

        try:
            graph, (g_logp, g_status) = capture_graph(lambda: step(d_plin, d_f, d_DA, d_H, d_cols))
            for _ in range(args.warmup):
                graph.replay()
            torch.cuda.synchronize()
        except Exception as exc:  # report and fall back to eager launches
            print(f"bench: CUDA graph capture failed ({exc}); timing eager launches", file=sys.stderr)
            graph = None
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    if args.profile_range:  # ncu --profile-from-start off: capture the timed steps only (no plan-build / autotune launches)
        torch.cuda.cudart().cudaProfilerStart()
    w0 = time.time()
    for i in range(args.steps):
        flush.zero_()  # evict L2 between timed iterations (not timed)
        ev[i][0].record()
        if graph is not None:
            graph.replay()
            logp, status = g_logp, g_status
        else:
            logp, status = step(d_plin, d_f, d_DA, d_H, d_cols)
        ev[i][1].record()
    torch.cuda.synchronize()
    w1 = time.time()
    if args.profile_range:
        torch.cuda.cudart().cudaProfilerStop()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop(w0, w1)
    dev_ms = sum(a.elapsed_time(bb) for a, bb in ev)

    # ---- end to end through the package's host driver (engine.HostPipeline): every step copies its inputs from pinned
    # host memory (one packed H2D transfer), evaluates, and reads log-likelihoods + multipoles back to pinned host
    # memory; two slots, so the copies of neighbouring steps overlap the kernels - all of it inside the timed region
    from eftpipe_b200.engine import HostPipeline

    def e2e_device(plin, f, DA, H, cols):
        lp, _ = step(plin, f, DA, H, cols)
        return lp, like.residuals(B)  # the multipoles minus the data, straight from the likelihood's workspace

    host_arrays = dict(plin=b.plin, f=b.f, DA=b.DA, H=b.H, cols=cols)
    pipe = HostPipeline(e2e_device, {n: tuple(np.asarray(a).shape) for n, a in host_arrays.items()}, nslots=2,
                        use_graph=graph is not None)
    for slot in range(2):
        for n, a in host_arrays.items():
            pipe.host_in(slot)[n].copy_(torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)))
    nslots = 2
    for i in range(max(2, args.warmup)):
        pipe.submit(i % nslots)
    pipe.join()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        pipe.submit(i % nslots)
    pipe.join()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    e2e_logp = pipe.wait((args.steps - 1) % nslots)[0].numpy().copy()
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes

    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        from eftpipe_b200.shard import gather_points

        all_logp = gather_points(logp, B * world)  # the only cross-GPU traffic of the path: per-point log-likelihoods
        assert all_logp.shape[0] == B * world
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    total = B * world * args.steps
    value = total / (dev_ms * 1e-3)
    e2e_value = total / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- per-stage device times and the roofline of the dominant kernel (rank 0)
    stage_ms, roof = stage_profile(dp, like, lib, S, d_plin, d_f, d_DA, d_H, d_cols, terms_bm, B, args)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"batch-shard x{world}",
                   "l2": "256 MiB buffer written between timed iterations; per-step working set ~%d MB" % (dp.lib.eftb_workspace_bytes(dp.handle, B) // 2**20),
                   "launch": "one CUDA graph replay per step" if graph is not None else "eager launches",
                   "e2e": "engine.HostPipeline: packed pinned inputs -> H2D -> step -> D2H of logp + multipoles, two slots "
                          "(copies of neighbouring steps overlap the kernels)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches["n"] * args.steps,
        "roofline": roof, "stage_ms": stage_ms,
        "plan_build_s": {"loop_plan": round(t_plan, 2), "window_LRG": round(S["t_window"], 2)},
        "logp_check": {"finite": bool(torch.isfinite(logp).all()), "status_nonzero": int((status != 0).sum()),
                       "e2e_equals_device_path": bool(np.array_equal(e2e_logp, logp.cpu().numpy()))},
    }
    if world == 1 and not args.no_cpu:
        # same arrangement as `--impl reference`: one single-threaded worker process per host core (the fastest way to
        # run the numpy path: ~12x the rate of one process with all BLAS threads), 2 evaluations per worker, 2 repeats
        cores = os.cpu_count() or 1
        nproc = max(1, min(cores, 96))
        ncheck = min(B, 2 * nproc)
        rates, ref_logp = cpu_arm(S, nproc, ncheck, repeats=2)
        got = logp[:ncheck].cpu().numpy()
        line["cpu_baseline"] = {"value": float(max(rates)), "unit": UNIT, "cores": nproc, "kind": "port",
                                "sample": f"first {ncheck} points of the batch, best of 2 repeats, {nproc} single-threaded worker "
                                          f"processes (host has {cores} cores); oracle/pybird_oracle.py restatement of the "
                                          "reference numpy path"}
        line["logp_check"]["max_rel_err_vs_oracle"] = float(np.max(np.abs(got - ref_logp) / np.abs(ref_logp)))
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------ config 3 (informational)
def run_multitracer(args):
    """BASELINE configs[2]/[3]: the DR16 NGC LRG x ELG x cross likelihood (3 tracer pipelines, 142 data points, 14
    analytically marginalised parameters) through the reference-facing API (theory.EFTLSS + likelihood.EFTLike).
    Not the driver's bench line (that is config 2); run with --workload config3 to get the north-star figure."""
    import torch
    import torch.distributed as dist

    from eftpipe_b200 import likelihood, synthetic, theory
    from eftpipe_b200.engine import capture_graph

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fx = load_fixture()
    B = args.batch
    ap = dict(Om_AP=0.307115, rdrag_AP=147.66, h_AP=0.6777, APst=True)
    tracers = {
        "LRG_NGC": dict(prefix="LRG_NGC_", z=0.696, nd=4.5e-5, window=dict(window_configspace_array=fx["win_LRG"])),
        "ELG_NGC": dict(prefix="ELG_NGC_", z=0.849, nd=2.3e-4, window=dict(window_configspace_array=fx["win_ELG"])),
        "X_NGC": dict(prefix="X_NGC_", z=0.763, cross=["LRG_NGC", "ELG_NGC"], window=dict(window_configspace_array=fx["win_X"])),
        "default": dict(km=0.7, kr=0.25, with_IRresum=True, with_APeffect=True, with_window=True, APeffect=ap,
                        window=dict(accboost=4, windowk=0.1)),
    }
    west = {n: {"scale": None} for n in ("b3", "cct", "cr1", "cr2", "ce0", "cequad")}
    marg = {"LRG_NGC_": west, "ELG_NGC_": dict(west), "X_NGC_ce0": {"scale": None}, "X_NGC_cequad": {"scale": None}}
    if world > 1 and rank != 0:
        dist.barrier()
    t0 = time.time()
    like = likelihood.EFTLike(
        tracers=["LRG_NGC", "ELG_NGC", "X_NGC"], chained=[False, True, False],
        data={"LRG_NGC": dict(table=fx["NGC_LRG_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20),
              "ELG_NGC": dict(table=fx["NGC_ELG_Q"], ls=[0, 2], kmin=0.03, kmax=0.20, symbol="Q"),
              "X_NGC": dict(table=fx["NGC_X_P"], ls=[0, 2, 4], kmin=0.02, kmax=0.20)},
        cov=dict(matrix=fx["cov_NGC_L024E02X024_PQP"], Nreal=1000), with_binning=True, jeffreys=True, marg=marg)
    th = theory.EFTLSS(tracers).must_provide(like.get_requirements()).initialize()
    like.initialize_with_provider(th)
    t_setup = time.time() - t0
    if world > 1 and rank == 0:
        dist.barrier()
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    cosmo = {}
    for name, z in (("LRG_NGC", 0.696), ("ELG_NGC", 0.849), ("X_NGC", 0.763)):
        b = synthetic.make_batch(B, z, seed=20261018 + 3 + 1000 * rank, unique=min(B, 32))
        cosmo[name] = dict(pkh=dev(b.plin), f=dev(b.f), DA=dev(b.DA), H=dev(b.H))
    rng = np.random.default_rng(11 + rank)
    params = {}
    for pre, b1 in (("LRG_NGC_", 2.1), ("ELG_NGC_", 1.4)):
        params[pre + "b1"] = dev(b1 + 0.05 * rng.standard_normal(B))
        c2 = 0.7 + 0.1 * rng.standard_normal(B)
        params[pre + "b2"], params[pre + "b4"] = dev(c2 / np.sqrt(2)), dev(c2 / np.sqrt(2))

    def step():
        th.calculate(cosmo)
        res = like.calculate(params)
        return res["logp"], res["status"]

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    graph = None
    if not args.no_graph:
        try:
            graph, (g_logp, g_status) = capture_graph(step)
            graph.replay()
            torch.cuda.synchronize()
        except Exception as exc:
            print(f"bench: CUDA graph capture failed ({exc}); timing eager launches", file=sys.stderr)
            graph = None
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    w0 = time.time()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        if graph is not None:
            graph.replay()
            logp, status = g_logp, g_status
        else:
            logp, status = step()
        ev[i][1].record()
    torch.cuda.synchronize()
    w1 = time.time()
    clocks = sampler.stop(w0, w1)
    t = torch.tensor([sum(a.elapsed_time(b_) for a, b_ in ev)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t[0])
    if rank == 0:
        line = {"metric": METRIC, "value": B * world * args.steps / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "config3: DR16 NGC LRG x ELG x cross, 3 tracer pipelines (IRresum + AP + window + binning, ELG chained),"
                                       " 142 data points, 14 marginalised parameters (Jeffreys), through theory.EFTLSS + likelihood.EFTLike",
                           "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"batch-shard x{world}",
                           "launch": "one CUDA graph replay per step" if graph is not None else "eager launches"},
                "clocks": clocks, "setup_s": round(t_setup, 1),
                "logp_check": {"finite": bool(torch.isfinite(logp).all()), "status_nonzero": int((status != 0).sum())}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def P_DA():
    from eftpipe_b200 import synthetic

    return synthetic.angular_distance(0.307115, Z_AP)


def P_H():
    from eftpipe_b200 import synthetic

    return synthetic.hubble(0.307115, Z_AP)


def stage_profile(dp, like, lib, S, d_plin, d_f, d_DA, d_H, d_cols, terms_bm, B, args):
    """CUDA-event time of every stage entry point (stage-level C ABI) and the roofline of the slowest."""
    import ctypes as C

    import torch

    reps = max(3, min(args.steps, 10))
    F = dp.front(d_plin)
    D = dp.antidiag(F, B)
    f_bm, DA_bm, H_bm = (dp.to_batch_minor(x)[0] for x in (d_f, d_DA, d_H))
    P22, Cr = dp.spectral_grouped(D, f_bm, B)
    T, Cr = dp.group(F, P22, None, f_bm, B, Cr=Cr)
    T0 = T.clone()
    dp.resum(F, Cr, f_bm, T, B)
    Tap = dp.ap(T, DA_bm, H_bm, B)
    nuis_bm = dp.to_batch_minor(d_cols)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    ms = {
        "front": timed(lambda: dp.front(d_plin)),
        "antidiag": timed(lambda: dp.antidiag(F, B)),
        "spectral": timed(lambda: dp.spectral_grouped(D, f_bm, B)),
        "group": timed(lambda: dp.group(F, P22, None, f_bm, B, Cr=Cr)),
        "resum": timed(lambda: dp.resum(F, Cr, f_bm, T0, B)),
        "ap": timed(lambda: dp.ap(T, DA_bm, H_bm, B)),
        "project": timed(lambda: dp.project(Tap, B)),
        "likelihood": timed(lambda: like.eval(B, [terms_bm], [f_bm], nuis_bm)),
    }
    # FP64 ceilings measured live: DFMA probe (library) and cuBLAS DGEMM 8192^3 (torch)
    tf = C.c_double()
    lib.eftb_probe_fp64(20000, C.byref(tf), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    torch.matmul(a, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.matmul(a, a)
    e1.record()
    torch.cuda.synchronize()
    dgemm_tf = 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a
    g = dp.host.grid
    npair = dp.host.pair_table.shape[0]
    Nmax = g.NFFT
    nslots = P.resum_slot_count(dp.host.resum)
    # algorithmic FP64 flops per evaluation of each stage, as implemented
    flops = {
        "front": 2.0 * dp.host.Wf.size,
        "antidiag": 2.0 * npair * (4 + 4 * P.NCH),
        "spectral": 2.0 * (28 * g.Nk + g.Nl * 12 * g.Ns + g.Nl * 38) * 2 * (Nmax + 1),
        # a = 1 half: Horner sweep of the nslots polynomials + slot sums + 13 row dots per (l, l', k, s);
        # a = 0 half: per slot a (Nkr x Ns)(Ns x NIR) product on the tensor pipe + the Q_0 k^{2(p+1)} fold
        "resum": 2.0 * (g.Nl * g.Nkr * g.Ns * (dp.host.resum["NIR"] * nslots + nslots + 13 * g.Nl)
                        + nslots * dp.host.resum["NIR"] * g.Nkr * (g.Ns + 2 * g.Nl)),
        "ap": 2.0 * (g.Nl * g.Nk * g.Nk * g.nterm + g.Nk * dp.host.ap["mu"].size * (4 * g.Nl * g.Nl + g.Nl * g.Nl + 40)
                     + g.Nl * g.Nk * g.nterm * g.Nl * 8),
        "project": 2.0 * dp.host.project.size * g.nterm,
        "likelihood": 2.0 * like.cfg.ndata * (like.cfg.ndata * (like.cfg.ngauss + 1) + (like.cfg.ngauss + 1) * (like.cfg.ngauss + 2) / 2 + 30),
    }
    top = max((k for k in ms if k in flops), key=lambda k: ms[k])
    peak = max(tf.value, dgemm_tf)
    achieved = flops[top] * B / (ms[top] * 1e-3) / 1e12
    roof = {"bound": "tensor", "kernel": top, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, per launch, from the committed
            # `ncu --set full` capture at B = 1024 (profiles/r1_resum_kernel_v9.txt, r1_top_kernels_v8.txt), scaled to this batch
            "traffic": NCU_DRAM_BYTES_PER_POINT.get(top, 0.0) * B or None,
            "peak_source": "FP64 measured live on this GPU: max(DFMA probe %.1f, cuBLAS DGEMM 8192^3 %.1f TFLOP/s); "
                           "MEASURED_PEAKS.json has no FP64 entry" % (tf.value, dgemm_tf),
            "per_stage_tflops": {k: flops[k] * B / (ms[k] * 1e-3) / 1e12 for k in flops},
            "reference_formulation_flops_per_eval": 5.56e9,
            "effective_tflops_reference_formulation": 5.56e9 * B / ((ms["antidiag"] + ms["spectral"]) * 1e-3) / 1e12}
    return {k: round(v, 4) for k, v in ms.items()}, roof


from eftpipe_b200 import plan as P  # noqa: E402  (host-only module; used in stage_profile)

# DRAM bytes (read + write) per evaluation point of each stage's kernels, from profiles/r1_top_kernels_v6.txt and, for the
# final resum_kernel, r1_resum_kernel_v9.txt (same traffic: 60.74 + 0.14 MB)
# (ncu --set full --clock-control none, B = 1024, config 2): resum_kernel 60.75 + 0.10 MB; antidiag_kernel 12.9 + 100.1 MB;
# spectral = regroup 160.0 + 99.8, P22 GEMM 118.2 + 4.7, C(s) GEMM 152.6 + 3.9 MB; ap = Cinv GEMM 29.6, geom 1.2 + 1.3 (its
# operator stays in L2), apply 70.0 + 4.5 MB
NCU_DRAM_BYTES_PER_POINT = {"resum": 60.85e6 / 1024, "antidiag": 113.0e6 / 1024, "spectral": 539.2e6 / 1024, "ap": 106.6e6 / 1024}


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print to file descriptor 1 (NCCL's version banner, for one) goes to stderr from here on; the
    one JSON line of the contract is written to the original stdout by `emit`."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="points per GPU per step")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3"],
                    help="config2 = the bench line (single tracer); config3 = informational multi-tracer likelihood")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (use with `ncu --profile-from-start off`)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "config3":
        return run_multitracer(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
