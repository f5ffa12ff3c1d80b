"""The CUDA path (through the C ABI) against the committed fixtures of tests/golden/ -- nothing under
oracle/ is executed here; the expected bytes were frozen by tests/golden/make_golden.py.
Bars: rope / rms_norm / KV rows bit-exact; attention fp32 1e-4 relative, bf16 2e-2 max-abs."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, assert_close, load_pkg, n2f, n2t
from test_golden_cpu import SDPA_MASKS, run_appendix_case, sdpa_mask

pytestmark = pytest.mark.gpu
omx = load_pkg()
DEV = "cuda"
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def dev(a, dt):
    return n2t(a, dt, DEV)


def test_reference_rope_vector_on_device():
    f = load("ref_rope_seed71.npz")
    got = omx.fast.rope(dev(f["x"], "f32"), 8, False, 10000.0, 1.0, 0)
    assert_bits_equal(got, f["out"], "f32", "test_rope vector")
    assert abs(float(got.double().mean()) - float(f["ref_mean"])) <= 0.02 * float(f["ref_mean"])
    assert abs(float(got.double().sum()) - float(f["ref_sum"])) <= 0.02 * float(f["ref_sum"])


def test_reference_rms_norm_vector_on_device():
    f = load("ref_rms_norm_seed103.npz")
    got = omx.fast.rms_norm(dev(f["x"], "f32"), torch.ones(16, device=DEV), 1e-5)
    assert_bits_equal(got, f["out"], "f32", "test_rms_norm vector")
    assert abs(float(got.double().sum()) - float(f["ref_sum"])) <= 0.02 * float(f["ref_sum"])


def test_kvcache_appendix_a_table_on_device():
    table = json.load(open(os.path.join(G, "kvcache_appendix_a.json")))

    def append(c, n, seed):
        g = torch.Generator().manual_seed(seed)
        k, v = torch.randn((1, 2, n, 8), generator=g), torch.randn((1, 2, n, 8), generator=g)
        kk, vv = c.update_and_fetch(k.to(DEV), v.to(DEV))
        assert torch.equal(kk[:, :, -n:].cpu(), k) and torch.equal(vv[:, :, -n:].cpu(), v)
    for case in table["cases"].values():
        run_appendix_case(lambda step: omx.KVCache(step), append, lambda c: c.state()[0].shape[2], case, table["step"])


def test_rope_cases_bit_exact():
    f = load("rope_cases.npz")
    for name, dims, trad, base, scale, offset, use_f in json.loads(str(f["meta"])):
        fr = dev(f[f"{name}.freqs"], "f32") if use_f else None
        for dt in ("f32", "bf16", "f16"):
            got = omx.fast.rope(dev(f[f"{name}.{dt}.x"], dt), dims, trad, base, scale, offset, fr)
            assert_bits_equal(got, f[f"{name}.{dt}.out"], dt, f"rope {name} {dt}")


def test_rms_norm_cases_bit_exact():
    f = load("rms_norm_cases.npz")
    for D in (64, 128):
        for dt in ("f32", "bf16"):
            got = omx.fast.rms_norm(dev(f[f"d{D}.{dt}.x"], dt), dev(f[f"d{D}.{dt}.w"], dt), 1e-6)
            assert_bits_equal(got, f[f"d{D}.{dt}.out"], dt, f"rms_norm D={D} {dt}")


@pytest.mark.parametrize("mk", SDPA_MASKS)
def test_sdpa_cases(mk):
    f = load("sdpa_cases.npz")
    for name, B, Hq, Hkv, Lq, Lk, D in json.loads(str(f["meta"])):
        for dt in ("f32", "bf16"):
            m = sdpa_mask(f, name, dt, mk)
            if isinstance(m, str):
                gm = omx.fast.ScaledDotProductAttentionMask.Causal
            elif m is None:
                gm = None
            else:
                gm = torch.from_numpy(m).to(DEV) if m.dtype == np.bool_ else dev(m, dt)
            q, k, v = (dev(f[f"{name}.{dt}.{t}"], dt) for t in "qkv")
            got = omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, gm)
            assert_close(got.float().cpu().numpy(), n2f(f[f"{name}.{dt}.out_{mk}"], dt), dt, f"sdpa {name} {dt} {mk}")


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("tag", ["plain", "norm"])
def test_decode_step_one_launch(dt, tag):
    f = load("decode_step.npz")
    D = 128
    S = f[f"{dt}.k0"].shape[2]
    c = omx.KVCache()
    c.update_and_fetch(dev(f[f"{dt}.k0"], dt), dev(f[f"{dt}.v0"], dt))
    rope = omx.nn.Rope(D, False, 1e6, 1.0)
    qn = kn = None
    if tag == "norm":
        qn, kn = omx.nn.RmsNorm(dev(f[f"{dt}.q_w"], dt), 1e-6), omx.nn.RmsNorm(dev(f[f"{dt}.k_w"], dt), 1e-6)
    omx.launch_count(reset=True)
    got = omx.attn_decode_fused(dev(f[f"{dt}.q"], dt), dev(f[f"{dt}.k_new"], dt), dev(f[f"{dt}.v_new"], dt), c, rope,
                                D ** -0.5, q_norm=qn, k_norm=kn)
    torch.cuda.synchronize()
    assert c.offset() == S + 1
    assert_close(got.float().cpu().numpy(), n2f(f[f"{dt}.{tag}.out"], dt), dt, f"decode step {dt} {tag}")
    sk, sv = c.state()
    assert sk.shape[2] == int(f[f"{dt}.{tag}.cap"])
    assert_bits_equal(sk[:, :, S], f[f"{dt}.{tag}.k_row"], dt, "appended key row")
    assert_bits_equal(sv[:, :, S], f[f"{dt}.{tag}.v_row"], dt, "appended value row")
    assert_bits_equal(sk[:, :, :S], f[f"{dt}.k0"], dt, "earlier key rows")
    assert not sk[:, :, S + 1:].any() and not sv[:, :, S + 1:].any()


@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_dit_joint_fixture(dt):
    f = load("dit_joint.npz")
    q, k, v, cos, sin = (dev(f[f"{dt}.{t}"], dt) for t in ("q", "k", "v", "cos", "sin"))
    D = q.shape[-1]
    qr, kr = omx.dit.apply_rope(q, cos, sin), omx.dit.apply_rope(k, cos, sin)
    assert_bits_equal(qr, f[f"{dt}.q_rope"], dt, "dit rope q")
    assert_bits_equal(kr, f[f"{dt}.k_rope"], dt, "dit rope k")
    got = omx.dit.joint_attention(qr, kr, v, D ** -0.5, out_dtype=torch.float32)
    want = np.swapaxes(f[f"{dt}.out"], 1, 2)
    assert_close(got.cpu().numpy(), want, dt, "dit joint attention")
    fused = omx.dit.attn_fused(q, k, v, D ** -0.5, cos=cos, sin=sin, out_dtype=torch.float32)
    assert_close(fused.cpu().numpy(), want, dt, "dit fused block attention")
