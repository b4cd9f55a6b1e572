"""ctypes binding of libeftb200.so (the C ABI declared in include/eftb200.h).

There is NO CPU fallback: if the shared library is missing or a CUDA device is absent, the
per-evaluation entry points raise.  (Plan construction, which is host-side precompute, works
without a GPU and is what the CPU test-suite exercises.)
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EFTB200_LIB") or os.path.join(_HERE, "libeftb200.so")  # override: A/B testing of kernel variants

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class EftbConfig(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ("Nl", "Nk", "Ns", "Nmax", "nterm", "with_nnlo",
                                  "nin", "ntail", "ntailx", "front_rows",
                                  "row_cre", "row_cim", "row_p11", "row_p13", "row_c11", "row_cct",
                                  "row_cctnnlo", "row_x", "row_y")]
        + [(n, C.c_double) for n in ("inv_dlog", "wx_last", "wx_prev")]
        + [("npair", C.c_int32)]
        + [(n, C.c_int32) for n in ("has_resum", "NIR", "Na", "Nkr", "Nklow", "qdeg")]
        + [(n, C.c_int32) for n in ("has_ap", "nmu", "nint", "ap_st")]
        + [(n, C.c_double) for n in ("da_fid", "h_fid")]
        + [(n, C.c_int32) for n in ("has_project", "nout", "nl_out")]
        + [(n, C.c_int32) for n in ("row_cre_cf", "row_cim_cf")]
    )


class EftbConstants(C.Structure):
    _fields_ = [
        ("k", c_double_p), ("l11", c_double_p), ("lct", c_double_p), ("lctnnlo", c_double_p),
        ("l22", c_double_p), ("l13", c_double_p), ("Wf", c_double_p), ("lr", c_double_p),
        ("lrx", c_double_p), ("pair_table", c_double_p), ("pair_offsets", c_int32_p),
        ("Ak", c_double_p), ("As", c_double_p), ("R", c_double_p), ("q", c_double_p),
        ("kr2", c_double_p), ("Cinv", c_double_p), ("knot_lo", c_double_p), ("basis", c_double_p),
        ("mu", c_double_p), ("wl", c_double_p), ("project", c_double_p), ("project_st", c_double_p),
    ]


class EftbLikeConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("ntracer", "ndata", "ngauss", "npar", "jeffreys")]


class EftbLikeConstants(C.Structure):
    _fields_ = [
        ("nout", c_int32_p), ("nterm", c_int32_p), ("scales", c_double_p), ("par_index", c_int32_p),
        ("eastcoast", c_int32_p), ("d_tracer", c_int32_p), ("d_row", c_int32_p), ("data", c_double_p),
        ("picc", c_double_p), ("invcov", c_double_p), ("g_count", c_int32_p), ("g_tracer", c_int32_p),
        ("g_term", c_int32_p), ("g_var", c_int32_p), ("g_coef", c_double_p), ("sigma_inv", c_double_p),
        ("sigma_inv_mu", c_double_p), ("mu_sigma_mu", C.c_double), ("d_row_g", c_int32_p),
        ("mode", c_int32_p), ("xb_off", c_int32_p), ("xg_off", c_int32_p),
    ]


# name -> (restype, argtypes); every symbol include/eftb200.h declares
_VP, _I, _SZ = C.c_void_p, C.c_int, C.c_size_t
SIGNATURES = {
    "eftb_abi_version": (C.c_int, []),
    "eftb_last_error": (C.c_char_p, []),
    "eftb_padded_batch": (C.c_int, [_I]),
    "eftb_launch_count": (C.c_ulonglong, []),
    "eftb_probe_fp64": (C.c_int, [_I, c_double_p, _VP]),
    "eftb_plan_create": (C.c_int, [C.POINTER(EftbConfig), C.POINTER(EftbConstants), C.POINTER(_VP)]),
    "eftb_plan_destroy": (None, [_VP]),
    "eftb_workspace_bytes": (_SZ, [_VP, _I]),
    "eftb_to_batch_minor": (C.c_int, [_VP, _I, _I, _VP, _VP]),
    "eftb_to_point_major": (C.c_int, [_VP, _I, _I, _VP, _VP, _VP]),
    "eftb_front": (C.c_int, [_VP, _I, _VP, _VP, _VP, _VP]),
    "eftb_antidiag": (C.c_int, [_VP, _I, _VP, _VP, _VP]),
    "eftb_antidiag_cf": (C.c_int, [_VP, _I, _VP, _VP, _VP]),
    "eftb_has_cf_set": (C.c_int, [_VP]),
    "eftb_spectral": (C.c_int, [_VP, _I, _VP, _VP, _VP, _VP, _VP]),
    "eftb_spectral_grouped": (C.c_int, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "eftb_group": (C.c_int, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "eftb_resum": (C.c_int, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    "eftb_resum_scratch_bytes": (C.c_size_t, [_VP, _I]),
    "eftb_ap": (C.c_int, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    "eftb_ap_scratch_bytes": (C.c_size_t, [_VP, _I]),
    "eftb_project": (C.c_int, [_VP, _I, _VP, _VP, _VP]),
    "eftb_eval_terms": (C.c_int, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "eftb_workspace_terms": (C.c_int, [_VP, _I, _VP, _SZ, _I, _VP, _VP]),
    "eftb_operator_create": (C.c_int, [_I, _I, c_double_p, C.POINTER(_VP)]),
    "eftb_operator_destroy": (None, [_VP]),
    "eftb_operator_apply": (C.c_int, [_VP, _VP, _VP, _I, _VP]),
    "eftb_eh_power": (C.c_int, [_I, _VP, C.c_double, C.c_double, C.c_double, C.c_double, _VP, _I, _VP, _VP, _I, _I, _VP, _VP, _VP, _VP,
                                _VP, _I, _VP]),
    "eftb_like_create": (C.c_int, [C.POINTER(EftbLikeConfig), C.POINTER(EftbLikeConstants), C.POINTER(_VP)]),
    "eftb_like_destroy": (None, [_VP]),
    "eftb_like_workspace_bytes": (_SZ, [_VP, _I]),
    "eftb_like_eval": (C.c_int, [_VP, _I, C.POINTER(_VP), C.POINTER(_VP), _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "eftb_like_eval_full": (C.c_int, [_VP, _I, C.POINTER(_VP), C.POINTER(_VP), _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "eftb_like_eval_priors": (C.c_int, [_VP, _I, C.POINTER(_VP), C.POINTER(_VP), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "eftb_like_residuals": (C.c_int, [_VP, _I, _VP, _VP, _VP]),
    "eftb_like_vectors": (C.c_int, [_VP, _I, C.POINTER(_VP), C.POINTER(_VP), _VP, _VP, _VP, _SZ, _VP]),
}

_lib = None


class EftbError(RuntimeError):
    pass


def load():
    """Load the shared library (no GPU needed to load / resolve symbols)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EftbError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(eftpipe_b200/csrc/build.sh).  There is no CPU fallback for the evaluation path."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().eftb_last_error().decode(errors="replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise EftbError(f"{what}: status {rc}: {msg}")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise EftbError("eftpipe_b200 evaluates on a CUDA device only (no CPU fallback); none is available")
    return torch


def as_ptr(arr, ctype=C.c_double):
    """Pointer to a C-contiguous numpy array (caller keeps `arr` alive)."""
    return arr.ctypes.data_as(C.POINTER(ctype))
