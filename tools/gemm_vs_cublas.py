#!/usr/bin/env python
"""The fixed-operator FP64 GEMM of libeftb200 (gemm_f64_kernel: DMMA m8n8k4, cp.async double buffering) against cuBLAS
(torch.matmul -> cublasDgemm) at every GEMM shape of the hot path, on the same box.  C[M][N] = A[M][K] X[K][N], X / C
batch-minor.  Usage on a GPU box:  python tools/gemm_vs_cublas.py [B ...]   (default B = 1024 8192)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from eftpipe_b200 import _lib

lib = _lib.load()
PEAK = None


def ours(A, X):
    M, K = A.shape
    h = C.c_void_p()
    a = np.ascontiguousarray(A.cpu().numpy())
    _lib.check(lib.eftb_operator_create(M, K, _lib.as_ptr(a), C.byref(h)), "create")
    out = torch.empty((M, X.shape[1]), dtype=torch.float64, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    run = lambda: _lib.check(lib.eftb_operator_apply(h, C.c_void_p(X.data_ptr()), C.c_void_p(out.data_ptr()), X.shape[1], s), "apply")
    return run, out, (lambda: lib.eftb_operator_destroy(h))


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return best


def main():
    Bs = [int(x) for x in sys.argv[1:]] or [1024, 8192]
    tf = C.c_double()
    lib.eftb_probe_fp64(20000, C.byref(tf), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    print(f"FP64 DFMA probe: {tf.value:.1f} TFLOP/s\n")
    print("| GEMM | M x K | N | ours ms | ours TF/s | frac of peak | cuBLAS ms | cuBLAS TF/s | ours / cuBLAS |")
    print("|---|---|---|---|---|---|---|---|---|")
    for B in Bs:
        shapes = [("front operator Wf", 1288, 247, B), ("D -> P22(k), 28 channels", 50, 514, 28 * B),
                  ("Dg -> Cloopl(s), 12 rows of one l", 80, 514, 12 * B), ("B-spline collocation Cinv, one l", 50, 50, 24 * B),
                  ("projection (window + binning), LRG", 54, 150, 24 * B), ("likelihood factor L^T, 142 data points", 142, 142, 15 * B)]
        for name, M, K, N in shapes:
            A = torch.randn(M, K, dtype=torch.float64, device="cuda")
            X = torch.randn(K, N, dtype=torch.float64, device="cuda")
            run, out, free = ours(A, X)
            t_o = timed(run)
            ref = torch.matmul(A, X)
            err = float((out - ref).abs().max() / ref.abs().max())
            assert err < 1e-12, err
            buf = torch.empty_like(ref)
            t_c = timed(lambda: torch.matmul(A, X, out=buf))
            fl = 2.0 * M * K * N
            print(f"| {name}, B={B} | {M} x {K} | {N} | {t_o:.4f} | {fl / t_o / 1e9:.1f} | {fl / t_o / 1e9 / tf.value:.2f} | {t_c:.4f} | "
                  f"{fl / t_c / 1e9:.1f} | {t_c / t_o:.2f} |")
            free()


if __name__ == "__main__":
    main()
