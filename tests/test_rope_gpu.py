"""fast::rope on the B200 vs the CPU oracle: BIT-EXACT for every dtype / mode / layout."""
import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, load_oracle, load_pkg, n2t, randn, t2n, tdt

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
from oracle import mlx_random  # noqa: E402

DEV = "cuda"


def _check(x_cpu, dtype, dims, trad, base, scale, offset, freqs=None, view=None):
    x = x_cpu.to(DEV)
    if view is not None:
        x = view(x)
    xin = x.clone()
    f = None if freqs is None else torch.tensor(freqs, dtype=torch.float32, device=DEV)
    got = omx.fast.rope(x, dims, trad, base, scale, offset, f)
    assert got.shape == x.shape and got.dtype == x.dtype
    assert torch.equal(x, xin), "input modified"
    want = orc.rope(t2n(x, dtype), dims, trad, base, scale, offset, freqs=freqs, dtype=dtype)
    assert_bits_equal(got, want, dtype, f"rope {dtype} dims={dims} trad={trad} off={offset}")
    return got


def test_reference_golden_vector_through_cuda():
    # mlx-rs/src/fast.rs:231-251 / nn/positional_encoding.rs:432-463
    a = torch.from_numpy(mlx_random.uniform_f32(mlx_random.RandomState(71), (2, 8, 16)))
    out = omx.fast.rope(a.to(DEV), 8, False, 10000.0, 1.0, 0, None)
    assert out.shape == (2, 8, 16) and out.dtype == torch.float32
    assert abs(out.double().mean().item() - 0.4562537670135498) <= 0.009125075340270997
    assert abs(out.double().sum().item() - 116.80096435546875) <= 2.3360192871093752
    # and far tighter than the reference's own 2 % tolerance:
    assert abs(out.double().sum().item() - 116.80096435546875) < 1e-4
    r = omx.nn.Rope(8).forward(a.to(DEV))  # nn::Rope path of the same test
    assert torch.equal(r, out)


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("trad", [False, True])
@pytest.mark.parametrize("dims,D", [(128, 128), (64, 128), (8, 16), (32, 80)])
def test_rope_modes_bit_exact(dtype, trad, dims, D):
    x = randn((2, 4, 7, D), dtype, seed=dims + D)
    _check(x, dtype, dims, trad, 10000.0, 1.0, 5)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("offset", [0, 255, 2047, 8191, 32767])
def test_rope_long_positions_bit_exact(dtype, offset):
    # fp32 angle at position ~32k must match the libm-based reference bit for bit
    x = randn((1, 8, 3, 128), dtype, seed=offset)
    _check(x, dtype, 128, False, 1e6, 1.0, offset)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_rope_caller_layout_transposed_view(dtype):
    # callers pass reshape([B,L,H,D]).transpose([0,2,1,3]) views (qwen3-mlx/src/model.rs:172-184)
    x = randn((2, 9, 4, 128), dtype, seed=3)
    _check(x, dtype, 128, False, 1e6, 1.0, 17, view=lambda t: t.transpose(1, 2))
    _check(x, dtype, 64, True, 1e4, 1.0, 17, view=lambda t: t.transpose(1, 2))


def test_rope_scale_and_freqs():
    x = randn((1, 2, 6, 64), "f32", seed=5)
    _check(x, "f32", 64, False, 10000.0, 0.25, 11)  # linear scaling (utils.rs:69-96)
    freqs = (500000.0 ** (np.arange(32) / 32)).astype(np.float32)
    _check(x, "f32", 64, False, None, 1.0, 11, freqs=freqs)
    _check(randn((1, 2, 6, 64), "bf16", seed=6), "bf16", 64, True, None, 0.5, 3, freqs=freqs)


def test_rope_ndim3_and_ndim5():
    _check(randn((2, 6, 16), "f32", seed=7), "f32", 8, False, 10000.0, 1.0, 2)
    _check(randn((2, 2, 3, 6, 16), "bf16", seed=8), "bf16", 16, False, 10000.0, 1.0, 2)


def test_rope_same_position_for_every_batch_row():
    # B > 1, L == 1 must NOT reproduce the Metal quirk
    x = randn((1, 8, 1, 128), "bf16", seed=9).repeat(64, 1, 1, 1)
    got = _check(x, "bf16", 128, False, 1e6, 1.0, 8191)
    assert all(torch.equal(got[b], got[0]) for b in range(1, 64))


def test_rope_dynamic_offset_matches_static():
    x = randn((2, 4, 3, 128), "bf16", seed=10).to(DEV)
    off = torch.tensor([777], dtype=torch.int32, device=DEV)
    a = omx.fast.rope_dynamic(x, 128, False, 1e6, 1.0, off, max_position=4096)
    b = omx.fast.rope(x, 128, False, 1e6, 1.0, 777)
    assert torch.equal(a, b)


def test_rope_config_shapes():
    # C1 / C2 / C5 operands
    _check(randn((1, 16, 1, 128), "f32", seed=11), "f32", 128, False, 1e6, 1.0, 2047)
    _check(randn((64, 32, 1, 128), "bf16", seed=12), "bf16", 128, False, 1e6, 1.0, 8191)
    _check(randn((1, 8, 1, 128), "bf16", seed=13), "bf16", 128, False, 1e6, 1.0, 32767)


def test_rope_prefill_sized_k_c3():
    # C3 keys: [8, 8, 8192, 128] bf16 (134 MB), offset 0, full compare against the oracle
    x = randn((8, 8, 8192, 128), "bf16", seed=14)
    _check(x, "bf16", 128, False, 1e6, 1.0, 0)


def test_rope_errors():
    x = torch.zeros(2, 4, 8, device=DEV)
    with pytest.raises(omx.Exception, match="at least 3 dimensions"):
        omx.fast.rope(torch.zeros(4, 8, device=DEV), 8, False, 10000.0, 1.0, 0)
    with pytest.raises(omx.Exception, match="Only one of base or freqs"):
        omx.fast.rope(x, 8, False, None, 1.0, 0)
    with pytest.raises(omx.Exception, match="Only one of base or freqs"):
        omx.fast.rope(x, 8, False, 10000.0, 1.0, 0, torch.ones(4, device=DEV))
    with pytest.raises(omx.Exception, match="floating"):
        omx.fast.rope(torch.zeros(2, 4, 8, dtype=torch.int32, device=DEV), 8, False, 10000.0, 1.0, 0)
    with pytest.raises(omx.Exception, match="dims"):
        omx.fast.rope(x, 10, False, 10000.0, 1.0, 0)


def test_rope_empty():
    out = omx.fast.rope(torch.zeros(2, 0, 4, 8, device=DEV), 8, False, 10000.0, 1.0, 0)
    assert out.shape == (2, 0, 4, 8)
