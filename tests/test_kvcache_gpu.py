"""KVCache / ConcatKeyValueCache on the device vs the literal port of cache.rs: offsets,
capacities and the WHOLE backing buffer (zero tail, stale rows) bit for bit."""
import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, load_oracle, load_pkg, randn, t2n

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
DEV = "cuda"


class Pair:
    """Drives the device cache and the oracle with the same calls and compares everything."""

    def __init__(self, dtype="bf16", step=256, B=2, H=2, Dk=8, Dv=8, reserve=None, concat=False):
        self.g = omx.ConcatKeyValueCache() if concat else omx.KVCache(step)
        self.o = orc.ConcatKeyValueCache() if concat else orc.KVCache(step)
        if reserve:
            self.g.reserve(reserve)
        self.dtype, self.B, self.H, self.Dk, self.Dv = dtype, B, H, Dk, Dv
        self.seed = 0
        self.concat = concat

    def append(self, n, strided=False):
        self.seed += 1
        if strided:  # [B, n, H, D] storage viewed [B, H, n, D] as the callers do
            k = randn((self.B, n, self.H, self.Dk), self.dtype, 1000 + self.seed).to(DEV).transpose(1, 2)
            v = randn((self.B, n, self.H, self.Dv), self.dtype, 2000 + self.seed).to(DEV).transpose(1, 2)
        else:
            k = randn((self.B, self.H, n, self.Dk), self.dtype, 1000 + self.seed).to(DEV)
            v = randn((self.B, self.H, n, self.Dv), self.dtype, 2000 + self.seed).to(DEV)
        gk, gv = self.g.update_and_fetch(k, v)
        ok, ov = self.o.update_and_fetch(t2n(k, self.dtype), t2n(v, self.dtype))
        assert self.g.offset() == self.o.offset()
        assert tuple(gk.shape) == ok.shape and tuple(gv.shape) == ov.shape
        assert_bits_equal(gk, ok, self.dtype, "fetched keys")
        assert_bits_equal(gv, ov, self.dtype, "fetched values")
        if not self.concat:
            sk, sv = self.g.state()
            assert tuple(sk.shape) == self.o.keys.shape, (sk.shape, self.o.keys.shape)
            assert_bits_equal(sk, self.o.keys, self.dtype, "whole key buffer")
            assert_bits_equal(sv, self.o.values, self.dtype, "whole value buffer")
        return gk, gv

    def cap(self):
        return self.g.state()[0].shape[2]

    def reset(self):
        self.g.reset()
        self.o.reset()


@pytest.mark.parametrize("dtype", ["bf16", "f32", "f16"])
def test_appendix_a_cases(dtype):
    p = Pair(dtype)
    p.append(5)
    assert (p.g.offset(), p.cap()) == (5, 256)          # A1
    for _ in range(251):
        p.append(1)
    assert (p.g.offset(), p.cap()) == (256, 256)        # A2
    p.append(1)
    assert (p.g.offset(), p.cap()) == (257, 512)        # A3

    p = Pair(dtype)
    p.append(300)
    assert (p.g.offset(), p.cap()) == (300, 512)        # A4
    p.append(300)
    assert (p.g.offset(), p.cap()) == (600, 812)        # A5

    p = Pair(dtype)
    p.append(300)
    p.reset()
    p.append(10)
    assert (p.g.offset(), p.cap()) == (10, 512)         # A6 (stale rows compared by Pair.append)

    p = Pair(dtype)
    p.append(256)
    p.reset()
    p.append(300)
    assert (p.g.offset(), p.cap()) == (300, 768)        # A7

    p = Pair(dtype, Dk=16, Dv=8)
    p.append(3)                                         # A8: Dk != Dv
    sk, sv = p.g.state()
    assert sk.shape == (2, 2, 256, 16) and sv.shape == (2, 2, 256, 8)


def test_strided_inputs_and_custom_step():
    p = Pair("bf16", step=64, B=3, H=4, Dk=128, Dv=128)
    for n in (7, 1, 1, 60, 130, 1):
        p.append(n, strided=True)
    assert p.g.offset() == 200


def test_reserve_does_not_change_the_contract():
    p = Pair("bf16", reserve=4096)
    for n in (5, 251, 1, 300, 1):
        p.append(n)
    assert p.g.offset() == 558  # capacities and buffers are compared inside append()


def test_views_are_strided_slices_of_one_buffer():
    c = omx.KVCache()
    k = randn((1, 2, 3, 8), "f32", 1).to(DEV)
    gk, gv = c.update_and_fetch(k, k)
    assert gk.shape == (1, 2, 3, 8) and gk.stride(1) >= 256 * 8 and gk.stride(2) == 8
    assert c.max_size() is None


def test_prefill_then_decode_like_generate_loop():
    # Generate::next: prefill T tokens, then one token at a time (qwen3-mlx/src/model.rs:808-841)
    p = Pair("bf16", B=1, H=8, Dk=128, Dv=128)
    p.append(700)
    for _ in range(100):
        p.append(1)
    assert p.g.offset() == 800


def test_concat_cache():
    p = Pair("f32", concat=True)
    for n in (4, 1, 1, 9):
        p.append(n)
    assert p.g.offset() == 15
    p.reset()  # trait default: no-op (cache.rs:17-19)
    assert p.g.offset() == 15


def test_shape_and_dtype_errors():
    c = omx.KVCache()
    k = randn((1, 2, 3, 8), "f32", 1).to(DEV)
    c.update_and_fetch(k, k)
    with pytest.raises(omx.Exception, match="does not match the cache"):
        c.update_and_fetch(randn((1, 3, 3, 8), "f32", 1).to(DEV), randn((1, 3, 3, 8), "f32", 1).to(DEV))
    with pytest.raises(omx.Exception, match="dtype"):
        c.update_and_fetch(k.half(), k.half())
    with pytest.raises(omx.Exception, match="4-dimensional"):
        c.update_and_fetch(k[0], k[0])
    assert c.offset() == 3


def test_c2_sized_cache_fill_bit_exact():
    # C2-like geometry at reduced batch: [4, 8, 8192, 128] bf16 in one append (cap 8192)
    B, H, S, D = 4, 8, 8192, 128
    k = randn((B, H, S, D), "bf16", 5).to(DEV)
    v = randn((B, H, S, D), "bf16", 6).to(DEV)
    c = omx.KVCache()
    gk, gv = c.update_and_fetch(k, v)
    assert c.offset() == S and c.state()[0].shape[2] == S
    assert torch.equal(gk, k) and torch.equal(gv, v)


def test_view_survives_one_growth_step():
    # the reference returns refcounted arrays; here fetched views are borrowed, and the buffer a growth step
    # replaces stays allocated until the NEXT growth: a view taken before a growth still reads its rows after it
    c = omx.KVCache()
    k = randn((1, 2, 250, 64), "bf16", 5, DEV)
    K0, V0 = c.update_and_fetch(k, k)
    snap = K0.clone()
    k2 = randn((1, 2, 20, 64), "bf16", 6, DEV)
    K1, _ = c.update_and_fetch(k2, k2)  # 250 + 20 > 256: grows, rows move to a new buffer
    assert K1.data_ptr() != K0.data_ptr()
    filler = [torch.empty(1 << 20, device=DEV) for _ in range(8)]  # would reuse the block had it been freed
    for f in filler:
        f.fill_(1.0)
    torch.cuda.synchronize()
    assert torch.equal(K0, snap), "the pre-growth view was clobbered"
    assert torch.equal(K1[:, :, :250], snap)
