"""CPU: the C-ABI library loads, exports every symbol include/ccj_b200.h declares, and refuses to
run without a GPU (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "ccj_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ccj_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    syms = declared_symbols()
    for must in ["ccj_ctx_create", "ccj_ctx_destroy", "ccj_model_load", "ccj_fold_batch", "ccj_batch_prepare",
                 "ccj_batch_fill", "ccj_batch_traceback", "ccj_batch_fetch", "ccj_export_table4", "ccj_last_error"]:
        assert must in syms


def test_library_exports_every_declared_symbol(library):
    for name in declared_symbols():
        assert hasattr(library, name), f"libccj_b200.so does not export {name}"


def test_version(library):
    assert b"sm_100a" in library.ccj_version()


def test_no_cpu_fallback(library):
    """Without a CUDA device the context cannot be created -- the product never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    h = ctypes.c_void_p()
    assert library.ccj_ctx_create(0, ctypes.byref(h)) == -1
    assert not h.value
    import ccj_b200
    with pytest.raises(ccj_b200.CCJError):
        ccj_b200.Context()


def test_missing_library_is_loud(tmp_path):
    import ccj_b200
    with pytest.raises(ccj_b200.CCJError):
        ccj_b200.load_library(tmp_path / "libccj_b200.so")


def test_sources_target_sm_100a_only():
    from ccj_b200 import build
    flags = " ".join(build.NVCC_FLAGS)
    assert "arch=compute_100a,code=sm_100a" in flags and "-lineinfo" in flags
