"""pytest configuration: `gpu` marker, package/oracle loaders, tensor<->numpy helpers."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_pkg():
    """The product package (directory name has a hyphen, hence importlib)."""
    return importlib.import_module("ominix-mlx_b200")


def load_oracle():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="session")
def omx():
    return load_pkg()


@pytest.fixture(scope="session")
def oracle():
    return load_oracle()


# ---- torch <-> oracle array convention (bf16 travels as uint16 bit patterns) ----

def t2n(t, dtype):
    import torch
    t = t.detach().contiguous().cpu()
    if dtype == "bf16":
        return t.view(torch.int16).numpy().view(np.uint16).copy()
    return t.numpy().copy()


def n2t(a, dtype, device="cpu"):
    import torch
    if dtype == "bf16":
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(torch.bfloat16).to(device)
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def n2f(a, dtype):
    """oracle array -> float32 numpy."""
    if dtype == "bf16":
        return (np.ascontiguousarray(a, np.uint16).astype(np.uint32) << 16).view(np.float32)
    return np.asarray(a, np.float32)


TORCH_DT = {"f32": "float32", "bf16": "bfloat16", "f16": "float16"}


def tdt(dtype):
    import torch
    return getattr(torch, TORCH_DT[dtype])


def randn(shape, dtype, seed, device="cpu"):
    """N(0,1) in `dtype`, generated on CPU for reproducibility."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g, dtype=torch.float32).to(tdt(dtype)).to(device)


# ---- parity bars (north_star): fp32 1e-4 relative, bf16 2e-2 max-abs, KV/rope bit-exact ----

def assert_close(got, want, dtype, what=""):
    """got/want: float32 numpy arrays."""
    got = np.asarray(got, np.float32)
    want = np.asarray(want, np.float32)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.isfinite(got).all(), f"{what}: non-finite output"
    err = float(np.abs(got - want).max()) if got.size else 0.0
    if dtype == "f32":
        bound = 1e-4 * max(1.0, float(np.abs(want).max()) if want.size else 1.0)
        assert err <= bound, f"{what}: fp32 max-abs err {err:.3e} > {bound:.3e}"
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5, err_msg=what)
    else:
        assert err <= 2e-2, f"{what}: {dtype} max-abs err {err:.3e} > 2e-2"


def assert_bits_equal(got_t, want_np, dtype, what=""):
    """torch tensor vs oracle array, bit for bit."""
    g = t2n(got_t, dtype)
    w = np.ascontiguousarray(want_np)
    if dtype == "f32":
        g, w = g.view(np.uint32), w.view(np.uint32)
    elif dtype == "f16":
        g, w = g.view(np.uint16), w.view(np.uint16)
    assert g.shape == w.shape, (g.shape, w.shape)
    bad = int((g != w).sum())
    assert bad == 0, f"{what}: {bad} of {g.size} elements differ bitwise"
