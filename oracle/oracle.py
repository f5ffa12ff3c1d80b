"""CPU oracle for the attention hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module, and only as the checker or the timed CPU
baseline.  The product (ominix-mlx_b200/) never imports it and has no CPU path.

Contents
  * ctypes front-end of oracle/omx_oracle.c (rope, sdpa, DiT manual attention,
    DiT table rope) -- the op-by-op restatement of the MLX v0.30.1 CPU fallback
    graphs behind mlx-rs/src/fast.rs:15-46 and :121-151.
  * KVCache / ConcatKeyValueCache: literal numpy port of
    mlx-rs-core/src/cache.rs:45-195.
  * create_causal_mask / create_attention_mask: mlx-rs-core/src/utils.rs:134-188.
  * sdpa_numpy: independent numpy twin (float64) used to cross-check the C code.

Array convention: numpy arrays; float32 as np.float32, float16 as np.float16,
bfloat16 as np.uint16 BIT PATTERNS together with dtype="bf16".

Parity pin status: rope PINNED by the reference golden vector (see
tests/test_oracle_golden.py); sdpa / KV cache UNPINNED by the reference (it has
no value-level tests for them) -- cross-checked against torch fp64 and the
hand-derived cases of SURVEY.md Appendix A.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DT = {"f16": 9, "f32": 10, "bf16": 12}
_NP = {"f16": np.float16, "f32": np.float32, "bf16": np.uint16}
MASK_NONE, MASK_CAUSAL, MASK_BOOL, MASK_ADD = 0, 1, 2, 3


def build(force=False):
    so = os.path.join(_HERE, "libomx_oracle.so")
    src = os.path.join(_HERE, "omx_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "clean", "all"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.omx_oracle_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads():
    return lib().omx_oracle_num_threads()


def set_threads(n):
    lib().omx_oracle_set_threads(int(n))


# ---------------------------------------------------------------- bf16 helpers

def f32_to_bf16_bits(x):
    """Round-to-nearest-even float32 -> bfloat16 bit patterns (uint16)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = (u + (0x7FFF + ((u >> 16) & 1))) >> 16
    r = np.where(nan, (u >> 16) | 0x40, r)
    return r.astype(np.uint16)


def bf16_bits_to_f32(b):
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def to_f32(x, dtype):
    if dtype == "bf16":
        return bf16_bits_to_f32(x)
    return np.asarray(x, dtype=np.float32)


def from_f32(x, dtype):
    if dtype == "bf16":
        return f32_to_bf16_bits(x)
    return np.asarray(x, dtype=np.float32).astype(_NP[dtype])


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    a = np.ascontiguousarray(a)
    assert a.dtype == _NP[dtype], (a.dtype, dtype)
    return a


# ------------------------------------------------------------------------ rope

def rope_table(T, dims, base, scale, offset, freqs=None):
    """[T, dims/2] f32 cos and sin (before the cast to x.dtype)."""
    half = dims // 2
    c = np.empty((T, half), np.float32)
    s = np.empty((T, half), np.float32)
    f = None if freqs is None else np.ascontiguousarray(freqs, np.float32)
    lib().omx_oracle_rope_table(
        _ptr(c), _ptr(s), ctypes.c_int(T), ctypes.c_int(dims),
        ctypes.c_int(0 if base is None else 1),
        ctypes.c_float(0.0 if base is None else base), ctypes.c_float(scale),
        ctypes.c_int(offset), None if f is None else _ptr(f))
    return c, s


def rope(x, dims, traditional, base, scale, offset, freqs=None, dtype="f32"):
    """mlx_rs::fast::rope (mlx-rs/src/fast.rs:15-46); x: [..., T, D], ndim >= 3."""
    if x.ndim < 3:
        raise ValueError("[rope] Input must have at least 3 dimensions")
    if (base is None) == (freqs is None):
        raise ValueError("[rope] exactly one of base / freqs must be given")
    x = _c(x, dtype)
    B, T, D = x.shape[0], x.shape[-2], x.shape[-1]
    N = int(np.prod(x.shape[1:-2])) if x.ndim > 3 else 1
    out = np.empty_like(x)
    f = None if freqs is None else np.ascontiguousarray(freqs, np.float32)
    lib().omx_oracle_rope(
        _ptr(x), _ptr(out), ctypes.c_int(DT[dtype]), ctypes.c_int(B),
        ctypes.c_int(N), ctypes.c_int(T), ctypes.c_int(D), ctypes.c_int(dims),
        ctypes.c_int(1 if traditional else 0),
        ctypes.c_int(0 if base is None else 1),
        ctypes.c_float(0.0 if base is None else base), ctypes.c_float(scale),
        ctypes.c_int(offset), None if f is None else _ptr(f))
    return out


# -------------------------------------------------------------------- rms_norm

def rms_norm(x, weight, eps, dtype="f32"):
    """mlx_rs::fast::rms_norm (mlx-rs/src/fast.rs:163-180) over the last axis; weight [D] or None."""
    x = _c(x, dtype)
    D = x.shape[-1]
    out = np.empty_like(x)
    w = None if weight is None else _c(weight, dtype)
    if w is not None and w.shape != (D,):
        raise ValueError("[rms_norm] weight must be one-dimensional with the size of the last axis of x")
    lib().omx_oracle_rms_norm(_ptr(x), None if w is None else _ptr(w), _ptr(out), ctypes.c_int(DT[dtype]),
                              ctypes.c_long(x.size // D if D else 0), ctypes.c_int(D), ctypes.c_float(eps))
    return out


# ------------------------------------------------------------------------ sdpa

def _mask_args(mask, B, Hq, Lq, Lk, dtype):
    if mask is None:
        return MASK_NONE, None, 0, (ctypes.c_int64 * 4)(0, 0, 0, 0), None
    if isinstance(mask, str):
        assert mask == "causal"
        return MASK_CAUSAL, None, 0, (ctypes.c_int64 * 4)(0, 0, 0, 0), None
    m = np.asarray(mask)
    if m.dtype == np.bool_:
        mode, mdt, keep = MASK_BOOL, 0, np.ascontiguousarray(m).view(np.uint8)
    else:
        mode, mdt, keep = MASK_ADD, DT[dtype], _c(m, dtype)
    while keep.ndim < 4:
        keep = keep[None]
    full = (B, Hq, Lq, Lk)
    strides = []
    for ax in range(4):
        if keep.shape[ax] == full[ax]:
            strides.append(keep.strides[ax] // keep.itemsize)
        elif keep.shape[ax] == 1:
            strides.append(0)
        else:
            raise ValueError("[scaled_dot_product_attention] mask not broadcastable")
    return mode, _ptr(keep), mdt, (ctypes.c_int64 * 4)(*strides), keep


def sdpa(q, k, v, scale, mask=None, dtype="f32", bool_fill_neg_inf=False):
    """mlx_rs::fast::scaled_dot_product_attention (mlx-rs/src/fast.rs:121-151).

    q [B,Hq,Lq,D], k [B,Hkv,Lk,D], v [B,Hkv,Lk,Dv] -> [B,Hq,Lq,Dv]
    mask: None | "causal" | bool ndarray | float ndarray (in `dtype`).
    """
    q, k, v = _c(q, dtype), _c(k, dtype), _c(v, dtype)
    B, Hq, Lq, D = q.shape
    _, Hkv, Lk, _ = k.shape
    Dv = v.shape[-1]
    assert k.shape[0] == B and v.shape[:3] == k.shape[:3] and k.shape[3] == D
    assert Hq % Hkv == 0
    out = np.empty((B, Hq, Lq, Dv), _NP[dtype])
    mode, mptr, mdt, ms, keep = _mask_args(mask, B, Hq, Lq, Lk, dtype)
    lib().omx_oracle_sdpa(
        _ptr(q), _ptr(k), _ptr(v), _ptr(out), ctypes.c_int(DT[dtype]),
        ctypes.c_int(B), ctypes.c_int(Hq), ctypes.c_int(Hkv), ctypes.c_int(Lq),
        ctypes.c_int(Lk), ctypes.c_int(D), ctypes.c_int(Dv),
        ctypes.c_float(scale), ctypes.c_int(mode), mptr, ctypes.c_int(mdt), ms,
        ctypes.c_int(1 if bool_fill_neg_inf else 0))
    del keep
    return out


def sdpa_numpy(q, k, v, scale, mask=None):
    """Independent float64 twin (exact math, no per-op rounding)."""
    q, k, v = (np.asarray(a, np.float64) for a in (q, k, v))
    B, Hq, Lq, D = q.shape
    Hkv, Lk = k.shape[1], k.shape[2]
    G = Hq // Hkv
    k = np.repeat(k, G, axis=1)
    v = np.repeat(v, G, axis=1)
    s = np.einsum("bhid,bhjd->bhij", q * scale, k)
    if isinstance(mask, str):
        qi = np.arange(max(Lk - Lq, 0), max(Lk - Lq, 0) + Lq)[:, None]
        s = np.where(qi >= np.arange(Lk)[None, :], s, -np.inf)
    elif mask is not None:
        m = np.asarray(mask)
        s = np.where(m, s, -np.inf) if m.dtype == np.bool_ else s + m.astype(np.float64)
    s = s - s.max(-1, keepdims=True)
    p = np.exp(s)
    p /= p.sum(-1, keepdims=True)
    return np.einsum("bhij,bhjd->bhid", p, v)


def dit_attention(q, k, v, dtype, scale_or_div, use_mul=False, add_mask=None):
    """Manual DiT joint attention (klein_model.rs:474-483 / zimage_model.rs:368-384).

    q [B,H,Lq,D], k/v [B,H,Lk,D] -> float32 [B,H,Lq,D] (the f32 scalar array
    promotes the chain to f32 after the first matmul)."""
    q, k, v = _c(q, dtype), _c(k, dtype), _c(v, dtype)
    B, H, Lq, D = q.shape
    Lk = k.shape[2]
    out = np.empty((B, H, Lq, D), np.float32)
    m = None if add_mask is None else np.ascontiguousarray(add_mask, np.float32)
    lib().omx_oracle_dit_attention(
        _ptr(q), _ptr(k), _ptr(v), _ptr(out), ctypes.c_int(DT[dtype]),
        ctypes.c_int(B), ctypes.c_int(H), ctypes.c_int(Lq), ctypes.c_int(Lk),
        ctypes.c_int(D), ctypes.c_float(scale_or_div),
        ctypes.c_int(1 if use_mul else 0), None if m is None else _ptr(m))
    return out


def dit_rope(x, cos, sin, dtype):
    """Table-driven interleaved rope (klein_model.rs:124-162); x [B,S,H,D],
    cos/sin [B,S,D/2] in x's dtype."""
    x, cos, sin = _c(x, dtype), _c(cos, dtype), _c(sin, dtype)
    B, S, H, D = x.shape
    out = np.empty_like(x)
    lib().omx_oracle_dit_rope(_ptr(x), _ptr(cos), _ptr(sin), _ptr(out),
                              ctypes.c_int(DT[dtype]), ctypes.c_int(B),
                              ctypes.c_int(S), ctypes.c_int(H), ctypes.c_int(D))
    return out


def klein_rope_freqs(ids, axes_dim, theta):
    """compute_rope_freqs (flux-klein-mlx/src/klein_model.rs:53-107), f32.
    ids [B,S,len(axes)] f32 -> cos, sin [B,S,sum(axes)/2] (un-duplicated)."""
    ids = np.asarray(ids, np.float32)
    cs, sn = [], []
    for ax, dim in enumerate(axes_dim):
        half = dim // 2
        i = np.arange(half, dtype=np.float32)
        # 1.0 / theta.powf(2.0 * i / dim)  -- f32 throughout
        inv = (np.float32(1.0) / np.power(np.float32(theta), (np.float32(2.0) * i / np.float32(dim)).astype(np.float32)).astype(np.float32)).astype(np.float32)
        ang = (ids[:, :, ax:ax + 1] * inv[None, None, :]).astype(np.float32)
        cs.append(np.cos(ang.astype(np.float64)).astype(np.float32))
        sn.append(np.sin(ang.astype(np.float64)).astype(np.float32))
    return np.concatenate(cs, -1), np.concatenate(sn, -1)


# -------------------------------------------------------------------- KV cache

class ConcatKeyValueCache:
    """mlx-rs-core/src/cache.rs:45-85."""

    def __init__(self):
        self.keys = None
        self.values = None
        self._offset = 0

    def offset(self):
        return self._offset

    def max_size(self):
        return None

    def reset(self):  # trait default: does nothing (cache.rs:19)
        pass

    def update_and_fetch(self, keys, values):
        if self.keys is not None and self.values is not None:
            self.keys = np.concatenate([self.keys, keys], axis=-2)
            self.values = np.concatenate([self.values, values], axis=-2)
        else:
            self.keys, self.values = keys, values
        self._offset = self.keys.shape[-2]
        return self.keys, self.values


class KVCache:
    """mlx-rs-core/src/cache.rs:92-195, line for line."""

    def __init__(self, step=256):
        self.keys = None
        self.values = None
        self._offset = 0
        self.step = step

    def offset(self):
        return self._offset

    def max_size(self):
        return None

    def reset(self):  # cache.rs:130-132
        self._offset = 0

    def update_and_fetch(self, keys, values):
        prev = self._offset
        num_new = keys.shape[2]
        needs_grow = self.keys is None or (prev + num_new) > self.keys.shape[2]  # :141-144
        if needs_grow:
            b, n_kv = keys.shape[0], keys.shape[1]
            n_steps = (self.step + num_new - 1) // self.step  # :152
            new_size = n_steps * self.step
            new_k = np.zeros((b, n_kv, new_size, keys.shape[3]), keys.dtype)
            new_v = np.zeros((b, n_kv, new_size, values.shape[3]), values.dtype)
            if self.keys is not None and self.values is not None:
                old_k, old_v = self.keys, self.values
                if prev % self.step != 0:  # :165-172
                    old_k, old_v = old_k[:, :, :prev, :], old_v[:, :, :prev, :]
                self.keys = np.concatenate([old_k, new_k], axis=2)
                self.values = np.concatenate([old_v, new_v], axis=2)
            else:
                self.keys, self.values = new_k, new_v
        self._offset += num_new
        self.keys[:, :, prev:self._offset, :] = keys  # :187-188
        self.values[:, :, prev:self._offset, :] = values
        return self.keys[:, :, :self._offset, :], self.values[:, :, :self._offset, :]


# ----------------------------------------------------------------------- masks

def create_causal_mask(N, offset=0, window_size=None):
    """mlx-rs-core/src/utils.rs:134-153 -> bool [N, offset+N]."""
    offset = offset or 0
    rinds = np.arange(offset + N)[None, :]
    linds = np.arange(offset, offset + N)[:, None]
    mask = linds >= rinds
    if window_size is not None:
        mask = mask & (linds <= rinds + window_size)
    return mask


def create_attention_mask(T, cache_offset=None, cache_max_size=None, return_array=False):
    """mlx-rs-core/src/utils.rs:156-188; returns None | "causal" | bool array."""
    if T <= 1:
        return None
    offset, window = 0, None
    if cache_offset is not None:
        offset = cache_offset
        if cache_max_size is not None:
            window = cache_max_size
            offset = min(offset, window)
            return_array = return_array or (offset + T) > window
    if return_array:
        return create_causal_mask(T, offset, window)
    return "causal"
