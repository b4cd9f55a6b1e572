"""CPU: the C-ABI shared library loads and exports every symbol include/eftb200.h declares."""
import ctypes
import os
import re

import pytest

from eftpipe_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "eftb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eftb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header_symbols():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/eftb200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"


def test_abi_version_and_padding():
    lib = _lib.load()
    assert lib.eftb_abi_version() == 5
    assert lib.eftb_padded_batch(1) == 32 and lib.eftb_padded_batch(33) == 64 and lib.eftb_padded_batch(0) == 0


def test_null_arguments_are_rejected_without_a_gpu():
    lib = _lib.load()
    assert lib.eftb_plan_create(None, None, None) == -1
    assert b"NULL" in lib.eftb_last_error()
    assert lib.eftb_front(None, 4, None, None, None, None) == -1
    assert lib.eftb_workspace_bytes(None, 4) == 0


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.EftbError):
        _lib.require_cuda()


def test_library_is_not_older_than_its_sources():
    """guards against shipping a stale in-tree .so to the GPU box"""
    import glob

    src = glob.glob(os.path.join(ROOT, "eftpipe_b200", "csrc", "*.cu*")) + [os.path.join(ROOT, "include", "eftb200.h")]
    newest = max(os.path.getmtime(p) for p in src)
    assert os.path.getmtime(_lib.LIB_PATH) >= newest, "libeftb200.so is stale: run eftpipe_b200/csrc/build.sh"
