#!/usr/bin/env python
"""bench.py -- the attention hot path on 1..8 B200 (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c5|c3|c4] [--impl reference]

Default workload = BASELINE.json configs[1] (C2): Qwen3-8B-shape GQA decode, 32 q / 8 kv heads,
d=128, bf16, batch 64 per GPU, ctx 8192.  One step = ONE fused decode launch (rope(q), rope(k),
KV append, split-K attention) over the whole batch; the cache offset is rewound by one row after
every step so each step does identical work.  The 2 GiB KV working set is far larger than the
126 MB L2, so every step streams from HBM (no explicit flush needed; stated in `config`).

Prints ONE JSON line (rank 0): value = whole-job tokens/s with inputs resident in HBM;
e2e = the same metric through the public API with pinned HOST buffers (H2D of q/k/v, D2H of the
output inside the timed region); roofline = algorithmic HBM bytes per launch / CUDA-event time
against MEASURED_PEAKS.json; cpu_baseline = the CPU oracle (a port of the reference's MLX-CPU
op chain -- the reference itself cannot be built here) on a bounded sample.
`--impl reference` times that CPU port as the reference arm.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, dict)
    "c1": ("decode", dict(B=1, Hq=16, Hkv=8, D=128, S=2048, dtype="f32", label="C1 Qwen3-0.6B decode fp32 B1 ctx2048")),
    "c2": ("decode", dict(B=64, Hq=32, Hkv=8, D=128, S=8192, dtype="bf16", label="C2 Qwen3-8B GQA decode bf16 B64 ctx8192")),
    "c5": ("decode", dict(B=1, Hq=32, Hkv=8, D=128, S=32768, dtype="bf16", label="C5 Mixtral-8x7B decode bf16 B1 ctx32768")),
    "c3": ("prefill", dict(B=8, Hq=32, Hkv=8, D=128, S=8192, dtype="bf16", causal=True, label="C3 Qwen3-8B causal prefill bf16 B8 seq8192")),
    "c4": ("prefill", dict(B=4, Hq=24, Hkv=24, D=128, S=4608, dtype="bf16", causal=False, label="C4 FLUX.2-klein DiT joint attention bf16 B4 512txt+4096img")),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """NVML sampling thread: SM clock + throttle reasons DURING the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ----------------------------------------------------------------------------- CPU arm

def cpu_decode_sample(cfg, rows, reps, warmup=1, threads=None):
    """The oracle's decode step (rope q/k -> KVCache append -> sdpa) on `rows` batch rows.
    Returns ([seconds per timed step], threads).  The KV fill is one random 2-row block tiled over
    the batch (values do not change the work; generating 2 GiB of normals would dominate)."""
    import numpy as np
    from oracle import oracle as orc
    if threads:
        orc.set_threads(threads)
    Hq, Hkv, D, S, dt = cfg["Hq"], cfg["Hkv"], cfg["D"], cfg["S"], cfg["dtype"]
    rng = np.random.default_rng(1234)

    def mk(b, *shape):
        base = rng.standard_normal((min(b, 2),) + shape, dtype=np.float32)
        x = np.concatenate([base] * ((b + 1) // 2), 0)[:b] if b > 2 else base
        return orc.f32_to_bf16_bits(x) if dt == "bf16" else np.ascontiguousarray(x)

    cache = orc.KVCache()
    cache.update_and_fetch(mk(rows, Hkv, S - 1, D), mk(rows, Hkv, S - 1, D))
    q, kn, vn = mk(rows, Hq, 1, D), mk(rows, Hkv, 1, D), mk(rows, Hkv, 1, D)
    times = []
    for i in range(reps + warmup):
        t0 = time.perf_counter()
        off = cache.offset()
        qr = orc.rope(q, D, False, 1e6, 1.0, off, dtype=dt)
        kr = orc.rope(kn, D, False, 1e6, 1.0, off, dtype=dt)
        K, V = cache.update_and_fetch(kr, vn)
        orc.sdpa(qr, K if K.flags.c_contiguous else np.ascontiguousarray(K),
                 V if V.flags.c_contiguous else np.ascontiguousarray(V), D ** -0.5, None, dtype=dt)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
        cache._offset -= 1  # rewind like the GPU arm
    return times, orc.num_threads()


def cpu_prefill_sample(cfg, heads, reps):
    """One batch item, `heads` query heads of the prefill / DiT attention on the oracle."""
    import numpy as np
    from oracle import oracle as orc
    Hq, Hkv, D, S, dt = cfg["Hq"], cfg["Hkv"], cfg["D"], cfg["S"], cfg["dtype"]
    G = Hq // Hkv
    hk = max(1, heads // G)
    rng = np.random.default_rng(1234)

    def mk(*shape):
        x = rng.standard_normal(shape, dtype=np.float32)
        return orc.f32_to_bf16_bits(x) if dt == "bf16" else x

    q, k, v = mk(1, hk * G, S, D), mk(1, hk, S, D), mk(1, hk, S, D)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.sdpa(q, k, v, D ** -0.5, "causal" if cfg["causal"] else None, dtype=dt)
        times.append(time.perf_counter() - t0)
    times.sort()
    return times[len(times) // 2], orc.num_threads(), hk * G


def run_reference(args, kind, cfg, rank):
    """--impl reference: the CPU port of the reference's MLX-CPU op chain, all host threads."""
    if rank != 0:
        return
    t_start = time.time()
    if kind == "decode":
        t1, cores = cpu_decode_sample(cfg, 1, 1)
        budget = 100.0 / max(1, args.steps + args.warmup)
        rows = int(max(1, min(cfg["B"], budget / max(t1[0], 1e-6))))
        ts, cores = cpu_decode_sample(cfg, rows, args.steps, args.warmup)
        ms = 1e3 * sum(ts) / len(ts)
        value = rows / (ms / 1e3)
        metric, unit = "attn_decode_tokens_per_s", "tokens/s"
        sample = f"{rows} of {cfg['B']} batch rows per step (same ctx/heads), tokens/s = rows / step time"
        extra = dict(global_batch=cfg["B"], ctx=cfg["S"])
    else:
        t, cores, heads = cpu_prefill_sample(cfg, cfg["Hq"] // cfg["Hkv"], 1)
        ts = [cpu_prefill_sample(cfg, heads, 1)[0] for _ in range(min(args.steps, 3))]
        ms = 1e3 * sum(ts) / len(ts)
        flops = 4.0 * heads * cfg["S"] * cfg["S"] * cfg["D"] * (0.5 if cfg["causal"] else 1.0)
        value = flops / (ms / 1e3) / 1e12
        metric, unit = "attn_prefill_tflops", "TFLOP/s"
        sample = f"1 of {cfg['B']} batch items x {heads} of {cfg['Hq']} heads per step (FLOPs of the slice / time)"
        extra = dict(global_batch=cfg["B"], seq_len=cfg["S"])
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
        "config": dict(workload=cfg["label"], **extra),
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference (Rust + MLX v0.30.1, macOS-only build) cannot be built in this image; this is the "
                "oracle port of its MLX-CPU op chain on the host cores",
        "wall_s": round(time.time() - t_start, 1),
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="omx", choices=["omx", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="override per-GPU batch (parity/debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--graph", action="store_true", help="replay the steps from a CUDA graph")
    ap.add_argument("--graph-steps", type=int, default=16, help="steps captured per graph replay")
    ap.add_argument("--gather", default="peer", choices=["peer", "collective"],
                    help="c5 with --gpus > 1 (kv-head-sharded single sequence): exchange of the output heads "
                         "by peer stores fused into the decode kernel, or by an NCCL all-gather")
    ap.add_argument("--layout", default="head", choices=["head", "seq"],
                    help="c5 with --gpus > 1: shard the kv HEADS (BASELINE C5) or the SEQUENCE (rows of position p on "
                         "rank p %% N, all heads everywhere, float32 partials + log-sum-exp merge)")
    ap.add_argument("--mask", default="string", choices=["string", "array"],
                    help="c3: pass the causal mask as the \"causal\" mode string or as the bool [T,T] array the LLM "
                         "crates build with create_causal_mask (classified per tile, same tiles skipped)")
    ap.add_argument("--composite", action="store_true",
                    help="c3 / c4: time the whole composite entry point (q_norm/k_norm + rope + KV append / [txt;img] "
                         "concat in ONE prologue launch, then attention) instead of the attention call alone; FLOPs "
                         "counted are still the attention's")
    ap.add_argument("--rotate", type=int, default=1,
                    help="decode: cycle through R distinct KV caches so a working set smaller than L2 "
                         "is still read from HBM (R x KV bytes should exceed 126 MB)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    kind, cfg = WORKLOADS[args.workload]
    cfg = dict(cfg)
    if args.batch:
        cfg["B"] = args.batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, kind, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback for the product path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    omx = importlib.import_module("ominix-mlx_b200")
    pk = peaks()

    B, Hq, Hkv, D, S = cfg["B"], cfg["Hq"], cfg["Hkv"], cfg["D"], cfg["S"]
    tdt = torch.bfloat16 if cfg["dtype"] == "bf16" else torch.float32
    es = 2 if cfg["dtype"] == "bf16" else 4
    g = torch.Generator(device=dev).manual_seed(1234 + rank)

    def rn(*shape):
        return torch.randn(shape, generator=g, device=dev, dtype=torch.float32).to(tdt)

    scale = D ** -0.5
    e2e_drain = None
    sharded = kind == "decode" and args.workload == "c5" and world > 1
    if sharded:
        # C5: ONE sequence; rank r holds kv heads [r*Hkv/N, ...) and computes their q heads; the full
        # [B,Hq,1,D] output lands on every rank (strong scaling: total work is fixed).
        g0 = torch.Generator(device=dev).manual_seed(1234)  # replicated step inputs
        def rn0(*shape):
            return torch.randn(shape, generator=g0, device=dev, dtype=torch.float32).to(tdt)
        rope = omx.nn.Rope(D, False, 1e6, 1.0)
        if args.layout == "seq":
            eng = omx.parallel.SeqShardedDecode(Hq, Hkv, D, tdt, rope, scale, batch=B, gather=args.gather)

            def rewind(n=1):
                eng.position -= 1
                if omx.parallel.seq_shard_owner(eng.position, world) == rank:
                    eng.cache.trim(1)
            eng.rewind = rewind
        else:
            eng = omx.parallel.HeadShardedDecode(Hq, Hkv, D, tdt, rope, scale, batch=B, gather=args.gather)
        # the exchange waits on "my own signal count" instead of a host-side step number, so that a captured
        # step can be replayed (--graph): what a compiled host's decode loop would launch
        eng.auto_wait = bool(args.graph)
        for s0 in range(0, S - 1, 4096):
            n = min(4096, S - 1 - s0)
            kk, vv = rn0(B, Hkv, n, D), rn0(B, Hkv, n, D)
            eng.prefill(kk, vv)
        assert args.layout == "seq" or eng.cache.offset() == S - 1
        q, kn, vn = rn0(B, Hq, 1, D), rn0(B, Hkv, 1, D), rn0(B, Hkv, 1, D)

        def step():
            eng.step(q, kn, vn)
            eng.rewind(1)

        units = B
        alg_bytes = (2 * B * Hkv * S * D * es + 2 * (2 * B * Hkv * D * es)) // world + 2 * B * Hq * D * es
        alg_flops = 4.0 * B * Hq * S * D / world
        metric, unit = "attn_decode_tokens_per_s", "tokens/s"
        bound, peak, peak_unit = "hbm", pk["hbm"], "GB/s"
        hq_pin = [torch.empty_like(t, device="cpu").pin_memory() for t in (q, kn, vn)]
        for hp, t in zip(hq_pin, (q, kn, vn)):
            hp.copy_(t)
        out_pin = torch.empty((B, Hq, 1, D), dtype=tdt).pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in hq_pin)
        d2h = out_pin.numel() * out_pin.element_size()

        def step_e2e():
            q.copy_(hq_pin[0], non_blocking=True)
            kn.copy_(hq_pin[1], non_blocking=True)
            vn.copy_(hq_pin[2], non_blocking=True)
            o = eng.step(q, kn, vn)
            eng.rewind(1)
            out_pin.copy_(o, non_blocking=True)
    elif kind == "decode":
        caches = []
        for _ in range(max(1, args.rotate)):
            cache = omx.KVCache()
            CH = 8  # fill in chunks to bound temporary memory
            for s0 in range(0, S - 1, (S - 1 + CH - 1) // CH):
                n = min((S - 1 + CH - 1) // CH, S - 1 - s0)
                cache.update_and_fetch(rn(B, Hkv, n, D), rn(B, Hkv, n, D))
            assert cache.offset() == S - 1
            caches.append(cache)
        turn = [0]
        q, kn, vn = rn(B, Hq, 1, D), rn(B, Hkv, 1, D), rn(B, Hkv, 1, D)
        out = torch.empty((B, Hq, 1, D), dtype=tdt, device=dev)
        rope = omx.nn.Rope(D, False, 1e6, 1.0)

        def step():
            cache = caches[turn[0] % len(caches)]
            turn[0] += 1
            omx.attn_decode_fused(q, kn, vn, cache, rope, scale, out=out)
            cache.trim(1)

        units = B  # tokens per step per GPU
        alg_bytes = 2 * B * Hkv * S * D * es + 2 * B * Hq * D * es + 2 * (2 * B * Hkv * D * es)
        alg_flops = 4.0 * B * Hq * S * D
        metric, unit = "attn_decode_tokens_per_s", "tokens/s"
        bound, peak, peak_unit = "hbm", pk["hbm"], "GB/s"
        hq_pin = [torch.empty_like(t, device="cpu").pin_memory() for t in (q, kn, vn)]
        for hp, t in zip(hq_pin, (q, kn, vn)):
            hp.copy_(t)
        h2d = sum(t.numel() * t.element_size() for t in hq_pin)
        d2h = out.numel() * out.element_size()
        # e2e: the serving loop a host runs -- step t+1's q/k/v upload (copy stream) and step t-1's output
        # download (second copy stream) overlap step t's kernel through double-buffered device tensors; every
        # step's H2D and D2H are inside the timed region and ordered by events, nothing is skipped.
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        dbuf = [dict(q=torch.empty_like(q), k=torch.empty_like(kn), v=torch.empty_like(vn), o=torch.empty_like(out),
                     pin=torch.empty_like(out, device="cpu").pin_memory(),
                     ev_in=torch.cuda.Event(), ev_k=torch.cuda.Event(), ev_out=torch.cuda.Event())
                for _ in range(2)]
        for d in dbuf:
            d["ev_k"].record()
            d["ev_out"].record()
        e2e_turn = [0]

        def step_e2e():
            d = dbuf[e2e_turn[0] & 1]
            e2e_turn[0] += 1
            cur = torch.cuda.current_stream()
            s_in.wait_event(d["ev_k"])  # the kernel that last read these inputs has finished
            with torch.cuda.stream(s_in):
                d["q"].copy_(hq_pin[0], non_blocking=True)
                d["k"].copy_(hq_pin[1], non_blocking=True)
                d["v"].copy_(hq_pin[2], non_blocking=True)
                d["ev_in"].record()
            cur.wait_event(d["ev_in"])
            cur.wait_event(d["ev_out"])  # the previous download of this output buffer has finished
            cache = caches[turn[0] % len(caches)]
            turn[0] += 1
            omx.attn_decode_fused(d["q"], d["k"], d["v"], cache, rope, scale, out=d["o"])
            cache.trim(1)
            d["ev_k"].record()
            s_out.wait_event(d["ev_k"])
            with torch.cuda.stream(s_out):
                d["pin"].copy_(d["o"], non_blocking=True)
                d["ev_out"].record()

        def e2e_drain():
            cur = torch.cuda.current_stream()
            for d in dbuf:
                cur.wait_event(d["ev_out"])
    else:
        q, k, v = rn(B, Hq, S, D), rn(B, Hkv, S, D), rn(B, Hkv, S, D)
        out = torch.empty_like(q)
        mask = omx.fast.ScaledDotProductAttentionMask.Causal if cfg["causal"] else None
        if cfg["causal"] and args.mask == "array":
            mask = omx.create_causal_mask(S, 0, device=dev)

        def step():
            omx.fast.scaled_dot_product_attention(q, k, v, scale, mask, out=out)

        if args.composite and cfg["causal"]:
            # Attention::forward for L = S new tokens (qwen3-mlx/src/model.rs:172-212), caller layouts:
            # projections [B,L,H,D] viewed [B,H,L,D], merged-head output
            qc, kc, vc = (rn(B, S, h, D).transpose(1, 2) for h in (Hq, Hkv, Hkv))
            qn = omx.nn.RmsNorm((1 + 0.1 * rn(D).float()).to(tdt), 1e-6)
            kn_ = omx.nn.RmsNorm((1 + 0.1 * rn(D).float()).to(tdt), 1e-6)
            rope = omx.nn.Rope(D, False, 1e6, 1.0)
            pcache = omx.KVCache()
            merged = torch.empty((B, S, Hq, D), dtype=tdt, device=dev).transpose(1, 2)

            def step():
                pcache.reset()
                omx.attn_prefill_fused(qc, kc, vc, pcache, rope, scale, mask, out=merged, q_norm=qn, k_norm=kn_)
        elif args.composite:
            # one FLUX.2-klein double-stream block's attention (klein_model.rs:443-489): 512 txt + 4096 img tokens
            lens = (512, S - 512)
            qs, ks, vs = ([rn(B, n, h, D) for n in lens] for h in (Hq, Hkv, Hkv))
            ang = torch.rand((B, S, D // 2), generator=g, device=dev) * 6.28
            ct, st = ang.cos().to(tdt), ang.sin().to(tdt)
            nw = [omx.nn.RmsNorm((1 + 0.1 * rn(D).float()).to(tdt), 1e-6) for _ in range(4)]

            def step():
                omx.dit.attn_fused(qs, ks, vs, scale, cos=ct, sin=st, q_norm=nw[:2], k_norm=nw[2:])

        units = B * S
        alg_flops = 4.0 * B * Hq * S * S * D * (0.5 if cfg["causal"] else 1.0)
        alg_bytes = (2 * B * Hq * S * D + 2 * B * Hkv * S * D) * es
        metric, unit = "attn_prefill_tflops", "TFLOP/s"
        bound, peak, peak_unit = "tensor", pk["tf_burst"], "TFLOP/s"
        pins = [torch.empty_like(t, device="cpu").pin_memory() for t in (q, k, v)]
        out_pin = torch.empty_like(out, device="cpu").pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in pins)
        d2h = out.numel() * out.element_size()

        def step_e2e():
            q.copy_(pins[0], non_blocking=True)
            k.copy_(pins[1], non_blocking=True)
            v.copy_(pins[2], non_blocking=True)
            omx.fast.scaled_dot_product_attention(q, k, v, scale, mask, out=out)
            out_pin.copy_(out, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, graph=False, drain=None):
        """Returns the device time (ms) of exactly `steps` steps (max over ranks).  `drain` makes the timing
        stream wait for work the steps put on side streams before the closing event is recorded."""
        for _ in range(warmup):
            fn()
        run, reps = fn, steps
        if graph:
            per = max(1, min(args.graph_steps, steps))
            while steps % per:
                per -= 1
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm the capture stream's scratch before capturing
                for _ in range(3):
                    fn()
            side.synchronize()
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg, stream=side):
                for _ in range(per):
                    fn()
            run, reps = cg.replay, steps // per
            for _ in range(3):
                run()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        if drain is not None:
            drain()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    sampler = ClockSampler(physical_gpu_index(local))
    omx.launch_count(reset=True)
    sampler.start()
    total_ms = timed(step, args.steps, args.warmup, graph=args.graph)
    clocks = sampler.stop()
    launches = omx.launch_count(reset=True)
    launches_timed = launches * args.steps // (args.steps + args.warmup) if not args.graph else args.steps
    kernel = omx.last_kernel()
    ms_step = total_ms / args.steps
    e2e_ms = timed(step_e2e, max(3, min(args.steps, 200)), 3, drain=e2e_drain) / max(3, min(args.steps, 200))

    if sharded:
        value = units / (ms_step / 1e3)
        e2e_value = units / (e2e_ms / 1e3)
        achieved = alg_bytes / (ms_step / 1e3) / 1e9
    elif kind == "decode":
        value = units * world / (ms_step / 1e3)
        e2e_value = units * world / (e2e_ms / 1e3)
        achieved = alg_bytes / (ms_step / 1e3) / 1e9
    else:
        value = alg_flops * world / (ms_step / 1e3) / 1e12
        e2e_value = alg_flops * world / (e2e_ms / 1e3) / 1e12
        achieved = alg_flops / (ms_step / 1e3) / 1e12

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            if kind == "decode":
                rows = min(B, 16)
                ts, cores = cpu_decode_sample(cfg, rows, 3)
                t = sorted(ts)[1]
                cpu = {"value": rows / t, "unit": unit, "cores": cores, "kind": "port",
                       "sample": f"{rows} of {B} batch rows (same ctx/heads), median of 3 steps; tokens/s = rows/time"}
            else:
                t, cores, heads = cpu_prefill_sample(cfg, Hq // Hkv, 1)
                fl = 4.0 * heads * S * S * D * (0.5 if cfg["causal"] else 1.0)
                cpu = {"value": fl / t / 1e12, "unit": unit, "cores": cores, "kind": "port",
                       "sample": f"1 of {B} batch items x {heads} of {Hq} heads, one pass; FLOPs of the slice / time"}
        except Exception as e:  # the checker must never take the product number down with it
            cpu = {"value": None, "unit": unit, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
            "config": {"workload": cfg["label"], "per_gpu_batch": B, "global_batch": B if sharded else B * world,
                       "ctx" if kind == "decode" else "seq_len": S, "q_heads": Hq, "kv_heads": Hkv, "head_dim": D,
                       "parallelism": ((f"kv-head-sharded x{world}, output heads exchanged by " if args.layout == "head"
                                        else f"sequence-sharded x{world} (rows of position p on rank p % N), float32 "
                                             "partials exchanged by ")
                                       + ("peer stores fused into the decode kernel" if args.gather == "peer"
                                          else "NCCL all-gather")) if sharded
                       else f"batch-sharded x{world}, no data-path collective",
                       "l2_policy": "working set >> 126 MB L2 (streams from HBM every step)"
                       if alg_bytes * max(1, args.rotate) > 512e6 else "working set fits L2: warm-L2 number",
                       "kernel": kernel, "cuda_graph": bool(args.graph), "rotate_caches": args.rotate,
                       "mask": args.mask if kind == "prefill" and cfg.get("causal") else None,
                       "composite": bool(args.composite) if kind == "prefill" else None},
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms},
            "gpu_launches": launches_timed,
            "clocks": clocks,
            "roofline": {"bound": bound, "achieved": achieved, "peak": peak, "unit": peak_unit,
                         "frac": achieved / peak, "traffic": None, "peak_source": pk["src"],
                         "algorithmic_bytes_per_launch": alg_bytes if bound == "hbm" else None,
                         "algorithmic_flops_per_launch": alg_flops},
            "cpu_baseline": cpu,
        }
        tr = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
        if os.path.exists(tr):
            try:
                line["roofline"]["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
            except Exception:
                pass
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
