"""The compiled-language host layer (ominix-mlx_b200/host/omx_attn.hpp, C++ because the image has no
Rust toolchain): builds on CPU, and on a B200 its self-checking binary must pass -- rope / rms_norm
bit-exact vs the oracle, the KVCache worked examples of SURVEY Appendix A, sdpa like the reference's
own test shapes, and Attention::forward prefill + decode through the fused composites."""
import os
import subprocess

import pytest

from conftest import ROOT

HOST = os.path.join(ROOT, "ominix-mlx_b200", "host")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HOST], stdout=subprocess.DEVNULL)
    return os.path.join(HOST, "test_host")


def test_cpp_host_layer_builds_and_refuses_to_run_without_a_gpu():
    import torch
    exe = _build()
    assert os.path.exists(exe)
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert r.returncode == 2 and "no sm_100a device" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_host_layer_checks_pass_on_gpu():
    exe = os.path.join(HOST, "test_host")
    if not os.path.exists(exe):
        exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "all host-layer checks passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
