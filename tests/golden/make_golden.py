"""Generator of the committed parity fixtures in tests/golden/ (run from the repo root:
`python tests/golden/make_golden.py`).

TEST INFRASTRUCTURE. Where the vectors come from:

* The reference itself cannot be executed in the build container (Rust toolchain absent, MLX un-vendored and
  macOS-only: SURVEY.md 8c), so no fixture here is an output of the reference binary.
* `ref_*.npz` carry the ONLY value-level vectors the reference's own tests hold for this path: the seeded
  inputs of `test_rope` (mlx-rs/src/fast.rs:231-251 == nn/positional_encoding.rs:432-463, seed 71) and
  `test_rms_norm` (fast.rs:253-274, seed 103), reproduced through the restated MLX threefry key schedule
  (oracle/mlx_random.py), together with the statistics the reference asserts (`ref_mean`, `ref_sum`, copied as
  numbers from those test bodies) and the oracle's full output tensor whose statistics meet them.
* `kvcache_appendix_a.json` is the hand-derived (offset, capacity) table of SURVEY.md Appendix A
  (mlx-rs-core/src/cache.rs:134-194); the reference has no cache tests.
* Every other file freezes the ORACLE's output on seeded inputs (inputs stored inside the fixture, so the
  tests never depend on an RNG implementation): they detect drift of the oracle and give the CUDA path fixed
  byte-level targets; they are not independent evidence about the reference ("parity unpinned" for sdpa / KV
  cache, as DESIGN.md section 2 states).

bf16 tensors are stored as uint16 bit patterns.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import mlx_random, oracle as orc  # noqa: E402


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name), **arrays)
    print(name, {k: (getattr(v, "shape", None), str(getattr(v, "dtype", type(v)))) for k, v in arrays.items()})


def cast(a, dtype):
    """f32 ndarray -> the oracle's array convention for `dtype`."""
    if dtype == "bf16":
        return orc.f32_to_bf16_bits(a)
    if dtype == "f16":
        return a.astype(np.float16)
    return a.astype(np.float32)


def main():
    # ---- reference-held vectors -------------------------------------------------------------------
    x = mlx_random.uniform_f32(mlx_random.RandomState(71), (2, 8, 16))
    save("ref_rope_seed71.npz", x=x, out=orc.rope(x, 8, False, 10000.0, 1.0, 0, dtype="f32"),
         in_mean=np.float64(0.5082664489746094), in_sum=np.float64(130.1162109375),
         ref_mean=np.float64(0.4562537670135498), ref_sum=np.float64(116.80096435546875))
    x = mlx_random.uniform_f32(mlx_random.RandomState(103), (2, 8, 16))
    save("ref_rms_norm_seed103.npz", x=x, out=orc.rms_norm(x, np.ones(16, np.float32), 1e-5, dtype="f32"),
         ref_mean=np.float64(0.87293875), ref_sum=np.float64(223.47232))

    # ---- Appendix A -------------------------------------------------------------------------------
    cases = {
        "A1_A2_A3": {"calls": [["append", 5]] + [["append", 1]] * 251 + [["append", 1]],
                     "expect_last": [257, 512], "expect_at": {"0": [5, 256], "251": [256, 256]}},
        "A4_A5": {"calls": [["append", 300], ["append", 300]], "expect_last": [600, 812],
                  "expect_at": {"0": [300, 512]}},
        "A6": {"calls": [["append", 300], ["reset"], ["append", 10]], "expect_last": [10, 512], "expect_at": {}},
        "A7": {"calls": [["append", 256], ["reset"], ["append", 300]], "expect_last": [300, 768], "expect_at": {}},
    }
    with open(os.path.join(HERE, "kvcache_appendix_a.json"), "w") as f:
        json.dump({"step": 256, "source": "SURVEY.md Appendix A / mlx-rs-core/src/cache.rs:134-194", "cases": cases},
                  f, indent=1)
    print("kvcache_appendix_a.json")

    # ---- rope variants (Appendix B rope matrix) -----------------------------------------------------
    rng = np.random.default_rng(20261017)
    out = {}
    variants = [  # name, shape, dims, traditional, base, scale, offset, freqs?
        ("qwen3_decode", (2, 4, 1, 128), 128, False, 1e6, 1.0, 2047, False),
        ("qwen3_prefill", (1, 3, 9, 128), 128, False, 1e6, 1.0, 0, False),
        ("glm4_partial_trad", (1, 2, 5, 128), 64, True, 10000.0, 1.0, 11, False),
        ("linear_scaled", (1, 2, 4, 64), 64, False, 10000.0, 0.25, 100, False),
        ("freqs", (1, 2, 3, 32), 32, False, None, 1.0, 7, True),
    ]
    meta = []
    for name, shape, dims, trad, base, scale, offset, use_f in variants:
        xf = rng.standard_normal(shape).astype(np.float32)
        fr = (rng.uniform(1.0, 500.0, dims // 2)).astype(np.float32) if use_f else None
        for dt in ("f32", "bf16", "f16"):
            xi = cast(xf, dt)
            out[f"{name}.{dt}.x"] = xi
            out[f"{name}.{dt}.out"] = orc.rope(xi, dims, trad, base, scale, offset, freqs=fr, dtype=dt)
        if fr is not None:
            out[f"{name}.freqs"] = fr
        meta.append([name, dims, trad, base, scale, offset, use_f])
    out["meta"] = np.array(json.dumps(meta))
    save("rope_cases.npz", **out)

    # ---- rms_norm ---------------------------------------------------------------------------------
    out = {}
    for D in (64, 128):
        xf = rng.standard_normal((3, 5, D)).astype(np.float32)
        wf = rng.standard_normal(D).astype(np.float32)
        for dt in ("f32", "bf16"):
            xi, wi = cast(xf, dt), cast(wf, dt)
            out[f"d{D}.{dt}.x"], out[f"d{D}.{dt}.w"] = xi, wi
            out[f"d{D}.{dt}.out"] = orc.rms_norm(xi, wi, 1e-6, dtype=dt)
    save("rms_norm_cases.npz", **out)

    # ---- sdpa: every mask mode, GQA, decode and prefill shapes ----------------------------------------
    out = {}
    meta = []
    shapes = [("decode_gqa", (1, 8, 2, 1, 130, 128)), ("prefill_gqa", (1, 4, 2, 24, 24, 128)),
              ("chunked_prefill", (1, 4, 2, 7, 50, 64)), ("mha_d64", (1, 2, 2, 33, 33, 64))]
    for name, (B, Hq, Hkv, Lq, Lk, D) in shapes:
        qf = rng.standard_normal((B, Hq, Lq, D)).astype(np.float32)
        kf = rng.standard_normal((B, Hkv, Lk, D)).astype(np.float32)
        vf = rng.standard_normal((B, Hkv, Lk, D)).astype(np.float32)
        mb = rng.random((Lq, Lk)) > 0.3
        mb[:, 0] = True
        ma = rng.standard_normal((B, 1, Lq, Lk)).astype(np.float32)
        win = orc.create_causal_mask(Lq, Lk - Lq, window_size=8)
        for dt in ("f32", "bf16"):
            q, k, v = cast(qf, dt), cast(kf, dt), cast(vf, dt)
            out[f"{name}.{dt}.q"], out[f"{name}.{dt}.k"], out[f"{name}.{dt}.v"] = q, k, v
            mad = cast(ma, dt)
            out[f"{name}.{dt}.mask_add"] = mad
            for mk, m in (("none", None), ("causal", "causal"), ("bool", mb), ("add", mad), ("window", win)):
                out[f"{name}.{dt}.out_{mk}"] = orc.sdpa(q, k, v, D ** -0.5, m, dtype=dt)
        out[f"{name}.mask_bool"], out[f"{name}.mask_window"] = mb, win
        meta.append([name, B, Hq, Hkv, Lq, Lk, D])
    out["meta"] = np.array(json.dumps(meta))
    save("sdpa_cases.npz", **out)

    # ---- composite decode step (Attention::forward, qwen3-mlx/src/model.rs:172-212, L = 1) --------------
    out = {}
    B, Hq, Hkv, S, D = 2, 4, 1, 256, 128   # offset 256 == cap 256: this very append grows the cache to 512 rows (A3)
    for dt in ("f32", "bf16"):
        k0 = cast(rng.standard_normal((B, Hkv, S, D)).astype(np.float32), dt)
        v0 = cast(rng.standard_normal((B, Hkv, S, D)).astype(np.float32), dt)
        q = cast(rng.standard_normal((B, Hq, 1, D)).astype(np.float32), dt)
        kn = cast(rng.standard_normal((B, Hkv, 1, D)).astype(np.float32), dt)
        vn = cast(rng.standard_normal((B, Hkv, 1, D)).astype(np.float32), dt)
        qw = cast(rng.standard_normal(D).astype(np.float32), dt)
        kw = cast(rng.standard_normal(D).astype(np.float32), dt)
        for normed in (False, True):
            c = orc.KVCache()
            c.update_and_fetch(k0, v0)
            qq, kk = (orc.rms_norm(q, qw, 1e-6, dtype=dt), orc.rms_norm(kn, kw, 1e-6, dtype=dt)) if normed else (q, kn)
            qr = orc.rope(qq, D, False, 1e6, 1.0, S, dtype=dt)
            kr = orc.rope(kk, D, False, 1e6, 1.0, S, dtype=dt)
            K, V = c.update_and_fetch(kr, vn)
            o = orc.sdpa(qr, np.ascontiguousarray(K), np.ascontiguousarray(V), D ** -0.5, None, dtype=dt)
            tag = f"{dt}.{'norm' if normed else 'plain'}"
            out[f"{tag}.out"], out[f"{tag}.k_row"], out[f"{tag}.v_row"] = o, c.keys[:, :, S], c.values[:, :, S]
            out[f"{tag}.cap"] = np.int64(c.keys.shape[2])
        for nm, a in (("k0", k0), ("v0", v0), ("q", q), ("k_new", kn), ("v_new", vn), ("q_w", qw), ("k_w", kw)):
            out[f"{dt}.{nm}"] = a
    save("decode_step.npz", **out)

    # ---- DiT joint attention ([txt;img], klein_model.rs:460-483) -----------------------------------------
    out = {}
    B, H, D, txt, img = 1, 2, 128, 8, 40
    S = txt + img
    axes, theta = [32, 32, 32, 32], 2000.0
    ids = rng.integers(0, 64, (B, S, 4)).astype(np.float32)
    c, s = orc.klein_rope_freqs(ids, axes, theta)
    for dt in ("f32", "bf16"):
        q, k, v = (cast(rng.standard_normal((B, S, H, D)).astype(np.float32), dt) for _ in range(3))
        cd, sd = cast(c, dt), cast(s, dt)
        qr, kr = orc.dit_rope(q, cd, sd, dt), orc.dit_rope(k, cd, sd, dt)
        tr = lambda a: np.ascontiguousarray(np.swapaxes(a, 1, 2))  # noqa: E731
        o = orc.dit_attention(tr(qr), tr(kr), tr(v), dt, np.float32(np.sqrt(D)))
        for nm, a in (("q", q), ("k", k), ("v", v), ("cos", cd), ("sin", sd), ("q_rope", qr), ("k_rope", kr), ("out", o)):
            out[f"{dt}.{nm}"] = a
    out["ids"] = ids
    save("dit_joint.npz", **out)


if __name__ == "__main__":
    main()
