"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/omx_attn.h declares, and refuses to compute without a B200 (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, load_pkg

omx = load_pkg()


def _declared():
    src = open(os.path.join(ROOT, "include", "omx_attn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(omx_[a-z0-9_]+)\s*\(", src))
    names.discard("omx_error_handler_func")
    return sorted(names)


def test_header_declares_the_reference_entry_points():
    names = _declared()
    for must in ("omx_fast_rope", "omx_fast_rope_dynamic", "omx_fast_scaled_dot_product_attention",
                 "omx_kv_cache_update_and_fetch", "omx_kv_cache_reset", "omx_kv_cache_offset",
                 "omx_concat_kv_cache_update_and_fetch", "omx_attn_decode_fused", "omx_set_error_handler"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(omx.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    assert set(_declared()) == set(omx.EXPORTED_SYMBOLS)
    assert lib.omx_version() == 100


def test_header_compiles_as_plain_c(tmp_path):
    # the boundary is a C ABI: no C++ or torch types in the signatures
    import subprocess
    c = tmp_path / "t.c"
    c.write_text('#include "omx_attn.h"\nint main(void){omx_array a; a.ndim = 0; return (int)sizeof(a) == 0;}\n')
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(c), "-o", str(tmp_path / "t.o")])


def test_struct_layout_matches_ctypes():
    from importlib import import_module
    L = import_module("ominix-mlx_b200._lib")
    assert ctypes.sizeof(L.OmxArray) == 8 + 4 + 4 + 8 * 8 * 2
    assert ctypes.sizeof(L.OmxOptionalFloat) == 8
    assert (L.OMX_BOOL, L.OMX_INT32, L.OMX_FLOAT16, L.OMX_FLOAT32, L.OMX_BFLOAT16) == (0, 7, 9, 10, 12)  # mlx_dtype


def test_cpu_tensors_are_rejected_loudly():
    x = torch.zeros(1, 2, 3, 8)
    with pytest.raises(omx.Exception, match="no CPU fallback"):
        omx.fast.rope(x, 8, False, 10000.0, 1.0, 0)
    with pytest.raises(omx.Exception, match="no CPU fallback"):
        omx.fast.scaled_dot_product_attention(x, x, x, 1.0)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_gpu_means_error_not_fallback():
    sm = ctypes.c_int(0)
    assert omx.lib().omx_device_check(ctypes.byref(sm)) == 1
    assert omx.lib().omx_last_error()


def test_host_side_helpers_without_gpu():
    r = omx.initialize_rope(128, 1e6, False, None, 0)
    assert (r.dimensions, r.traditional, r.base, r.scale) == (128, False, 1e6, 1.0)
    r = omx.initialize_rope(64, 10000.0, True, {"type": "linear", "factor": 4.0}, 0)
    assert r.scale == 0.25 and r.traditional
    r = omx.initialize_rope(64, 10000.0, True, {"rope_type": "linear", "factor": "2"}, 0)
    assert r.scale == 0.5
    with pytest.raises(omx.Exception, match="Unsupported RoPE type"):
        omx.initialize_rope(64, 10000.0, False, {"type": "yarn", "factor": 2.0}, 0)
    with pytest.raises(omx.Exception, match='"factor" is not found'):
        omx.initialize_rope(64, 10000.0, False, {"type": "linear"}, 0)
    assert omx.create_attention_mask(torch.zeros(2, 1, 8)) is None
    assert omx.create_attention_mask(torch.zeros(2, 5, 8)) == omx.AttentionMask.Causal
    m = omx.create_causal_mask(3, 2, device="cpu")
    assert m.shape == (3, 5) and m.dtype == torch.bool
    from oracle import oracle as orc
    assert (m.numpy() == orc.create_causal_mask(3, 2)).all()
    w = omx.create_causal_mask(4, 1, window_size=2, device="cpu")
    assert (w.numpy() == orc.create_causal_mask(4, 1, 2)).all()


def test_rust_ffi_declares_the_same_symbols():
    # the Rust crate cannot be compiled in this image (no cargo); at least its extern block must
    # name exactly the functions the header exports
    src = open(os.path.join(ROOT, "ominix-mlx_b200", "rust", "src", "ffi.rs")).read()
    rust = set(re.findall(r"pub fn (omx_[a-z0-9_]+)\s*\(", src))
    assert rust == set(_declared()), sorted(rust ^ set(_declared()))
    for mod in ("fast", "cache", "utils", "array", "error", "ffi"):
        assert os.path.exists(os.path.join(ROOT, "ominix-mlx_b200", "rust", "src", mod + ".rs")), mod
