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
