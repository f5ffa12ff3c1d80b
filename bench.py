#!/usr/bin/env python
"""bench.py -- the attention hot path on 1..8 B200 (one process per GPU), every BASELINE.json config in ONE line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|c1|c2|c3|c4|c5] [--impl reference]

Headline (top level of the JSON line) = BASELINE.json configs[1] (C2): Qwen3-8B-shape GQA decode, 32 q / 8 kv
heads, d = 128, bf16, GLOBAL batch 64 at ctx 8192, **batch-sharded 64/N per GPU** (`scaling: "strong"`).  One
step = ONE fused decode launch per GPU (rope(q), rope(k), KV append, split-K attention over the GPU's rows),
replayed from a CUDA graph the way a compiled host drives a decode loop.  The per-GPU KV working set
(2 GiB / N) is far larger than the 126 MB L2, so every step streams from HBM.

`workloads` carries one sub-record per other config, each with ms_per_step / value / roofline / clocks / e2e:
  c1  Qwen3-0.6B-shape decode fp32 B1 ctx 2048 (the reference's own CPU-runnable case; replicas at N > 1),
      rotated through enough distinct caches that the reads are L2-cold
  c3  Qwen3-8B causal prefill bf16 B8 seq 8192 (tcgen05 flash attention; batch-sharded 8/N)
  c4  FLUX.2-klein / Z-Image DiT joint attention bf16 B4, 512 txt + 4096 img, 24 heads (batch-sharded; at N = 8
      additionally 12 / 12 heads within a batch item)
  c5  Mixtral-shape single-sequence decode ctx 32768: one GPU at N = 1; kv-head-sharded over N GPUs at N > 1 with
      the output heads exchanged by peer stores fused into the decode kernel (`c5`) and by ncclAllGather
      (`c5_collective`)
  c2_paged the headline shape through the PAGED cache (page pool + block table + per-sequence lengths)
  c2_weak (N > 1 only) the r01 spelling: batch 64 PER GPU.
At N > 1 every sharded workload also runs one un-timed step whose result is compared on rank 0 with the
unsharded computation of the same inputs (`parity_check`).

Timing: W warm-up steps, then blocks of exactly K steps bracketed by barrier + synchronize, CUDA events on the
launching stream, max over ranks; blocks repeat until >= 0.5 s have been timed and the MEDIAN block is
reported.  value = units / time with inputs resident in HBM; e2e = the same call with pinned HOST q/k/v
uploaded and the output downloaded inside the timed region.  roofline = algorithmic bytes (decode) or FLOPs
(prefill) per launch / CUDA-event time against MEASURED_PEAKS.json.  cpu_baseline / `--impl reference` = the
CPU oracle (a port of the reference's MLX-CPU op chain -- the reference itself cannot be built here) on all
host cores on a bounded sample.
"""
import argparse
import importlib
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c1": ("decode", dict(B=1, Hq=16, Hkv=8, D=128, S=2048, dtype="f32",
                          label="C1 Qwen3-0.6B decode fp32 B1 ctx2048")),
    "c2": ("decode", dict(B=64, Hq=32, Hkv=8, D=128, S=8192, dtype="bf16",
                          label="C2 Qwen3-8B GQA decode bf16 B64 ctx8192")),
    "c5": ("decode", dict(B=1, Hq=32, Hkv=8, D=128, S=32768, dtype="bf16",
                          label="C5 Mixtral-8x7B decode bf16 B1 ctx32768")),
    "c3": ("prefill", dict(B=8, Hq=32, Hkv=8, D=128, S=8192, dtype="bf16", causal=True,
                           label="C3 Qwen3-8B causal prefill bf16 B8 seq8192")),
    "c4": ("prefill", dict(B=4, Hq=24, Hkv=24, D=128, S=4608, dtype="bf16", causal=False,
                           label="C4 FLUX.2-klein DiT joint attention bf16 B4 512txt+4096img")),
    # not a BASELINE config: the PREFILL of config 1's model in its own dtype (float32 -> the tiled FFMA kernel,
    # DESIGN 4.3), so that the float32 multi-row path has a driver-side number too (N = 1 only)
    "c1_prefill": ("prefill", dict(B=1, Hq=16, Hkv=8, D=128, S=2048, dtype="f32", causal=True,
                                   label="C1 model prefill: Qwen3-0.6B causal prefill fp32 B1 seq2048")),
    # not BASELINE configs either: GLM-4.7-Flash's absorbed MLA (glm-4.7-flash-mlx/src/model.rs:263-299; 20 query heads
    # over ONE latent kv head, keys 512 + 64, values 512) through fast::sdpa on the mma.sync kernel, so
    # that the Dk != Dv path has driver-side numbers (N = 1 only)
    # Qwen3.5-35B-A3B full-attention layers (qwen3.5-35B-mlx/src/attention.rs: 16 q / 2 kv heads, head_dim 256, rope on
    # the first 64 features): the fused decode step outside head dim 128 -- prologue + mma.sync key groups
    "gqa256_decode": ("decode", dict(B=64, Hq=16, Hkv=2, D=256, S=8192, dtype="bf16", rope_dims=64,
                                     label="Qwen3.5-35B-A3B full-attention decode bf16 B64 ctx8192 (head_dim 256, 16/2 heads)")),
    "mla_decode": ("mla", dict(B=64, Hq=20, Hkv=1, D=576, Dv=512, S=4096, L=1, dtype="bf16", causal=False,
                               label="GLM-4.7-Flash absorbed-MLA decode bf16 B64 ctx4096 (keys 576 / values 512)")),
    "mla_prefill": ("mla", dict(B=1, Hq=20, Hkv=1, D=576, Dv=512, S=4096, L=4096, dtype="bf16", causal=True,
                                label="GLM-4.7-Flash absorbed-MLA causal prefill bf16 B1 seq4096 (keys 576 / values 512)")),
}
L2_BYTES = 126e6
MIN_TIMED_S = 0.5
BAD_REASONS = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


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
            self._stop.wait(0.01)

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


def workload_config(name, world):
    """The `config` object of a workload: a pure function of (name, N), printed identically by both arms."""
    kind, cfg = WORKLOADS[name.replace("_weak", "").replace("_collective", "").replace("_paged", "")]
    B, S = cfg["B"], cfg["S"]
    c = {"workload": cfg["label"], "q_heads": cfg["Hq"], "kv_heads": cfg["Hkv"], "head_dim": cfg["D"],
         "ctx" if kind == "decode" else "seq_len": S}
    if name in ("c2", "c2_paged"):
        c.update(global_batch=B, per_gpu_batch=B // world,
                 parallelism=f"batch-sharded {B}/{world} rows per GPU, no data-path collective")
        if name == "c2_paged":
            c["kv_cache"] = "paged: 64-row pages, block table, per-sequence lengths"
    elif name == "c2_weak":
        c.update(global_batch=B * world, per_gpu_batch=B,
                 parallelism=f"batch {B} per GPU x{world} (weak), no data-path collective")
    elif name == "c1":
        c.update(global_batch=B, per_gpu_batch=B,
                 parallelism="single GPU" if world == 1 else f"{world} independent replicas (the path does not shard)")
    elif name == "gqa256_decode":
        c.update(global_batch=B, per_gpu_batch=B, rope_dims=cfg["rope_dims"], parallelism="single GPU")
    elif name in ("c5", "c5_collective"):
        how = "peer stores fused into the decode kernel" if name == "c5" else "ncclAllGather"
        c.update(global_batch=B, per_gpu_batch=B,
                 parallelism="single GPU" if world == 1 else
                 f"kv-head-sharded {cfg['Hkv']}/{world} kv heads per GPU, output heads exchanged by {how}")
    elif name == "c1_prefill":
        c.update(global_batch=B, per_gpu_batch=B, mask="causal", parallelism="single GPU")
    elif kind == "mla":
        c.update(global_batch=B, per_gpu_batch=B, value_dim=cfg["Dv"], mask="causal" if cfg["causal"] else None,
                 parallelism="single GPU")
    elif name == "c3":
        c.update(global_batch=B, per_gpu_batch=B // world if world <= B else 1, mask="causal",
                 parallelism=f"batch-sharded {B}/{world} items per GPU, no data-path collective")
    elif name == "c4":
        split = "" if world <= B else f" x {world // B} head groups of {cfg['Hq'] * B // world}"
        c.update(global_batch=B, per_gpu_batch=max(1, B // world), mask=None,
                 parallelism=f"batch-sharded {B}/{min(world, B)} items per GPU{split}, no data-path collective")
    return c


# ----------------------------------------------------------------------------- CPU arm (the checker, timed)

def cpu_decode_sample(cfg, rows, reps, warmup=1, threads=None):
    """The oracle's decode step (rope q/k -> KVCache append -> sdpa) on `rows` batch rows.
    Returns ([seconds per timed step], threads).  The KV fill is one random 2-row block tiled over
    the batch (values do not change the work; generating 2 GiB of normals would dominate)."""
    import numpy as np
    from oracle import oracle as orc
    orc.set_threads(threads or host_cores())  # torchrun exports OMP_NUM_THREADS=1: pin the count ourselves
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


def cpu_prefill_sample(cfg, heads, reps, rows=None, threads=None):
    """One batch item, `heads` query heads of the prefill / DiT attention on the oracle.  `rows` bounds the
    sample to the LAST `rows` query rows against all S keys (bottom-right aligned causal mask, the heaviest
    rows): the MLX-CPU chain materialises the [Lq, Lk] score tensor, so full 8192 x 8192 slices cost seconds."""
    import numpy as np
    from oracle import oracle as orc
    orc.set_threads(threads or host_cores())
    Hq, Hkv, D, S, dt = cfg["Hq"], cfg["Hkv"], cfg["D"], cfg["S"], cfg["dtype"]
    G = Hq // Hkv
    hk = max(1, heads // G)
    Lq = min(S, rows or S)
    rng = np.random.default_rng(1234)

    def mk(*shape):
        x = rng.standard_normal(shape, dtype=np.float32)
        return orc.f32_to_bf16_bits(x) if dt == "bf16" else x

    q, k, v = mk(1, hk * G, Lq, D), mk(1, hk, S, D), mk(1, hk, S, D)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.sdpa(q, k, v, D ** -0.5, "causal" if cfg["causal"] else None, dtype=dt)
        times.append(time.perf_counter() - t0)
    times.sort()
    # FLOPs the slice really performs on the CPU chain: the dense [Lq, S] products (the mask is applied after)
    flops = 4.0 * hk * G * Lq * S * D
    return times[len(times) // 2], orc.num_threads(), hk * G, Lq, flops


def cpu_record(name, steps, warmup, budget_s):
    """One CPU-arm measurement of workload `name`, bounded to about `budget_s` seconds."""
    kind, cfg = WORKLOADS[name]
    if kind == "decode":
        t1, cores = cpu_decode_sample(cfg, 1, 1)
        per_step = budget_s / max(1, steps + warmup)
        rows = int(max(1, min(cfg["B"], per_step / max(t1[0], 1e-6))))
        ts, cores = cpu_decode_sample(cfg, rows, steps, warmup)
        ms = 1e3 * sorted(ts)[len(ts) // 2]
        return {"value": rows / (ms / 1e3), "unit": "tokens/s", "cores": cores, "kind": "port", "ms_per_step": ms,
                "sample": f"{rows} of {cfg['B']} batch rows per step (same ctx / heads), median of {len(ts)} steps; "
                          "tokens/s = rows / step time"}
    G = cfg["Hq"] // cfg["Hkv"]
    rows = 1024
    t, cores, heads, lq, fl = cpu_prefill_sample(cfg, G, max(1, min(steps, 3)), rows=rows)
    return {"value": fl / t / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "port", "ms_per_step": 1e3 * t,
            "sample": f"1 of {cfg['B']} batch items x {heads} of {cfg['Hq']} heads x the last {lq} of {cfg['S']} query "
                      f"rows against all {cfg['S']} keys; dense FLOPs of the slice (the CPU chain masks after the "
                      "product) / median time"}


def run_reference(args, rank, world):
    """--impl reference: the CPU port of the reference's MLX-CPU op chain, all host threads, same config."""
    if rank != 0:
        return
    t_start = time.time()
    names = ["c2", "c1", "c5", "c3", "c4"] if args.workload == "all" else [args.workload]
    head = names[0]
    kind, cfg = WORKLOADS[head]
    recs = {}
    for n in names:
        recs[n] = cpu_record(n, args.steps if n == head else min(args.steps, 5), args.warmup if n == head else 1,
                             60.0 if n == head else 8.0)
    r = recs[head]
    line = {
        "impl": "reference", "metric": "attn_decode_tokens_per_s" if kind == "decode" else "attn_prefill_tflops",
        "value": r["value"], "unit": r["unit"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if head in ("c2", "c3", "c4", "c5") else "weak", "vs_baseline": None,
        "dtype": cfg["dtype"], "data": "synthetic", "config": workload_config(head, world),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "workloads": {n: dict(config=workload_config(n, world), **recs[n]) for n in names if n != head},
        "note": "reference (Rust + MLX v0.30.1, macOS-only build) cannot be built in this image; this is the "
                "oracle port of its MLX-CPU op chain on the host cores (a CPU process: N does not change it)",
        "wall_s": round(time.time() - t_start, 1),
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm

class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback for the product path"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.omx = importlib.import_module("ominix-mlx_b200")
        self.pk = peaks()
        self.gpu_index = physical_gpu_index(self.local)

    # ---- data: every global batch row has its own seed, so any rank can regenerate any row (parity_check)
    def row_randn(self, seed, shape, tdt):
        g = self.torch.Generator(device=self.dev).manual_seed(int(seed))
        return self.torch.randn(shape, generator=g, device=self.dev, dtype=self.torch.float32).to(tdt)

    def rows_randn(self, tag, rows, shape, tdt):
        """[len(rows), *shape]: row r of the global batch comes from seed (tag, r)."""
        return self.torch.stack([self.row_randn(1234 + 1009 * tag + 7919 * r, shape, tdt) for r in rows])

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    # ---- timing
    def capture(self, fn, steps):
        """`steps` calls of fn captured into ONE CUDA graph (what a compiled host's decode loop replays)."""
        torch = self.torch
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm the capture stream's scratch before capturing (one whole block)
            for i in range(steps):
                fn(i)
        side.synchronize()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg, stream=side):
            for i in range(steps):
                fn(i)
        return cg

    def timed_blocks(self, run_block, steps, min_s=None, max_blocks=2000):
        """Blocks of exactly `steps` steps, each bracketed by barrier + synchronize and timed with CUDA events on
        the launching stream (max over ranks); repeated until min_s of device time.  Returns the block times."""
        torch = self.torch
        out, total = [], 0.0
        min_s = MIN_TIMED_S if min_s is None else min(min_s, MIN_TIMED_S)
        while (total < min_s * 1e3 or len(out) < 3) and len(out) < max_blocks:
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run_block()
            e1.record()
            torch.cuda.synchronize()
            ms = self.max_over_ranks(e0.elapsed_time(e1))
            out.append(ms)
            total += ms
        self.barrier()
        return out

    def measure(self, fn, steps, warmup, graph, min_s=None):
        """fn(i): step i of a block.  Returns dict(ms_per_step = median block / steps, blocks, clocks, launches)."""
        omx = self.omx
        for i in range(warmup):
            fn(i)
        self.torch.cuda.synchronize()
        omx.launch_count(reset=True)
        for i in range(steps):
            fn(i)
        per_block = omx.launch_count(reset=True)
        if graph:
            cg = self.capture(fn, steps)
            run_block = cg.replay
            for _ in range(2):
                run_block()
        else:
            def run_block():
                for i in range(steps):
                    fn(i)
        for attempt in range(2):
            sampler = ClockSampler(self.gpu_index)
            sampler.start()
            blocks = self.timed_blocks(run_block, steps, min_s)
            clocks = sampler.stop()
            if not (BAD_REASONS & set(clocks["reasons"])):
                break
            clocks["remeasured"] = True
        s = sorted(blocks)
        med = s[len(s) // 2]
        return {"ms_per_step": med / steps, "ms_per_step_min": s[0] / steps, "blocks": len(blocks),
                "timed_ms": sum(blocks), "clocks": clocks, "launches_per_block": per_block,
                "gpu_launches": per_block * len(blocks), "cuda_graph": bool(graph), "kernel": omx.last_kernel()}

    # ---- decode workloads (c1, c2, c2_weak, c5 at N = 1): rows [row0, row0 + Bl) of the global batch on this rank
    def run_decode(self, name, cfg, rows, value_scale, steps, warmup, want_cpu, paged=False):
        torch, omx = self.torch, self.omx
        Hq, Hkv, D, S = cfg["Hq"], cfg["Hkv"], cfg["D"], cfg["S"]
        tdt = torch.bfloat16 if cfg["dtype"] == "bf16" else torch.float32
        es = 2 if cfg["dtype"] == "bf16" else 4
        Bl = len(rows)
        tag = {"c1": 1, "c2": 2, "c2_weak": 2, "c2_paged": 2, "c5": 5, "gqa256_decode": 8}[name]
        kv_bytes = 2 * Bl * Hkv * S * D * es
        # L2 policy: a working set below ~2x L2 is rotated through R distinct caches so every step reads HBM
        R = 1
        if kv_bytes < 2 * L2_BYTES:
            need = int(math.ceil(2 * L2_BYTES / kv_bytes))
            divs = [d for d in range(1, steps + 1) if steps % d == 0 and d >= need]
            R = divs[0] if divs else steps
        caches = []
        for c in range(R):
            if paged:  # page pool sized for the batch's context + one spare page per sequence
                n_seq_pages = (S + steps + 63) // 64 + 1
                cache = omx.PagedKVCache(Bl, Hkv, D, tdt, n_pages=Bl * n_seq_pages, max_pages_per_seq=n_seq_pages)
            else:
                cache = omx.KVCache()
                cache.reserve(S + 256)
            if c == 0:
                for s0 in range(0, S - 1, 1024):  # fill in chunks to bound temporary memory
                    n = min(1024, S - 1 - s0)
                    kk = torch.stack([self.row_randn(1234 + 1009 * tag + 7919 * r + 31 * (s0 // 1024 + 1),
                                                     (Hkv, n, D), tdt) for r in rows])
                    vv = torch.stack([self.row_randn(4321 + 1009 * tag + 7919 * r + 31 * (s0 // 1024 + 1),
                                                     (Hkv, n, D), tdt) for r in rows])
                    if paged:
                        cache.update_and_fetch(kk, vv, fetch=False)
                    else:
                        cache.update_and_fetch(kk, vv)
            else:  # rotation copies: same values, distinct HBM addresses
                k0, v0 = caches[0].state()
                cache.update_and_fetch(k0[:, :, :S - 1], v0[:, :, :S - 1])
            assert cache.offset() == S - 1
            if paged:
                cache.reserve(steps + 2)  # no host-side page allocation inside the timed / captured region
            caches.append(cache)
        q = self.rows_randn(tag + 10, rows, (Hq, 1, D), tdt)
        kn = self.rows_randn(tag + 20, rows, (Hkv, 1, D), tdt)
        vn = self.rows_randn(tag + 30, rows, (Hkv, 1, D), tdt)
        out = torch.empty((Bl, Hq, 1, D), dtype=tdt, device=self.dev)
        rope = omx.nn.Rope(cfg.get("rope_dims", D), False, 1e6, 1.0)
        scale = D ** -0.5

        if paged:
            steps += steps & 1  # the device lengths are double-buffered by step parity: even blocks

        def fused(qq, kk, vv, oo, i, n):
            cache = caches[i % R]
            if paged:
                # the rewind is a (tiny) kernel here, so it runs once per block: rows S-1 .. S-2+n are appended,
                # then all n dropped (context 8192 .. 8191+n within a block: +0.1 % bytes, not counted)
                omx.attn_decode_fused_paged(qq, kk, vv, cache, rope, scale, out=oo)
                if i == n - 1:
                    cache.trim(n)
            else:
                omx.attn_decode_fused(qq, kk, vv, cache, rope, scale, out=oo)
                cache.trim(1)  # host-side rewind: every step appends row S-1 and attends S keys

        def step(i):
            fused(q, kn, vn, out, i % steps, steps)

        # whole blocks only (the paged rewind closes a block)
        res = self.measure(step, steps, steps * max(1, (warmup + steps - 1) // steps), graph=True)
        alg_bytes = 2 * Bl * Hkv * S * D * es + 2 * Bl * Hq * D * es + 2 * (2 * Bl * Hkv * D * es)
        alg_flops = 4.0 * Bl * Hq * S * D

        # ---- e2e: the serving loop a host runs, captured as a graph with three streams: step t+1's q/k/v upload
        # and step t-1's output download ride two copy streams under step t's kernel (double-buffered device
        # tensors, every H2D and D2H inside the timed region, ordered by events)
        pins = [torch.empty_like(t, device="cpu").pin_memory() for t in (q, kn, vn)]
        for hp, t in zip(pins, (q, kn, vn)):
            hp.copy_(t)
        h2d = sum(t.numel() * t.element_size() for t in pins)
        d2h = out.numel() * out.element_size()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        dbuf = [dict(q=torch.empty_like(q), k=torch.empty_like(kn), v=torch.empty_like(vn), o=torch.empty_like(out),
                     pin=torch.empty_like(out, device="cpu").pin_memory(),
                     ev_in=torch.cuda.Event(), ev_k=torch.cuda.Event(), ev_out=torch.cuda.Event())
                for _ in range(2)]

        def e2e_block(n):
            """n pipelined steps issued from the current stream; all side-stream work is joined at the end."""
            cur = torch.cuda.current_stream()
            fork = torch.cuda.Event()
            fork.record(cur)
            for i in range(n):
                d = dbuf[i & 1]
                s_in.wait_event(d["ev_k"] if i >= 2 else fork)  # the kernel that last read these inputs is done
                with torch.cuda.stream(s_in):
                    d["q"].copy_(pins[0], non_blocking=True)
                    d["k"].copy_(pins[1], non_blocking=True)
                    d["v"].copy_(pins[2], non_blocking=True)
                    d["ev_in"].record()
                cur.wait_event(d["ev_in"])
                if i >= 2:
                    cur.wait_event(d["ev_out"])  # the previous download of this output buffer is done
                fused(d["q"], d["k"], d["v"], d["o"], i, n)
                d["ev_k"].record(cur)
                s_out.wait_event(d["ev_k"])
                with torch.cuda.stream(s_out):
                    d["pin"].copy_(d["o"], non_blocking=True)
                    d["ev_out"].record()
            for d in dbuf[:min(n, 2)]:
                cur.wait_event(d["ev_out"])

        e2e = self.measure_e2e(e2e_block, steps)
        # the downloaded result is the kernel's output
        torch.cuda.synchronize()
        e2e_ok = None  # paged blocks append at advancing positions: the last resident / e2e steps are not the same step
        if not paged:
            e2e_ok = bool(torch.equal(dbuf[0]["pin"].view(torch.int16 if es == 2 else torch.int32),
                                      out.cpu().view(torch.int16 if es == 2 else torch.int32)))
        units = Bl * value_scale  # tokens per step, whole job
        ms, e2e_ms = res["ms_per_step"], e2e["ms_per_step"]
        rec = {
            "metric": "attn_decode_tokens_per_s", "unit": "tokens/s", "value": units / (ms / 1e3),
            "ms_per_step": ms, "ms_per_step_min": res["ms_per_step_min"], "blocks": res["blocks"], "steps": steps,
            "dtype": cfg["dtype"],
            "e2e": {"value": units / (e2e_ms / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "blocks": e2e["blocks"],
                    "cuda_graph": e2e["cuda_graph"], "output_matches_resident_run": e2e_ok,
                    "how": "double-buffered: upload of step t+1 and download of step t-1 overlap step t's kernel"},
            "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms / 1e3) / 1e9, "peak": self.pk["hbm"],
                         "unit": "GB/s", "frac": alg_bytes / (ms / 1e3) / 1e9 / self.pk["hbm"],
                         "peak_source": self.pk["src"] + " copy bandwidth (MEASURED_PEAKS.json hbm_gbs)",
                         "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_flops_per_launch": alg_flops,
                         "frac_of_8TBs_nominal": alg_bytes / (ms / 1e3) / 8e12},
            "run": {"kernel": res["kernel"], "cuda_graph": True, "rotate_caches": R,
                    "l2_policy": ("working set >> 126 MB L2: every step streams from HBM" if R == 1 else
                                  f"KV working set {kv_bytes / 1e6:.0f} MB < 2 x L2: rotated through {R} distinct "
                                  "caches so every step reads HBM (L2-cold)"),
                    "launches_per_step": res["launches_per_block"] / steps},
        }
        self.attach_traffic(rec, name)
        if paged:
            rec["run"]["cache"] = (f"PagedKVCache: pool [{caches[0].n_pages}][{Hkv}][64][{D}] per tensor, block table "
                                   f"[{Bl}][{caches[0].max_pages_per_seq}]; rewind = one 1-block kernel per block")
        state = dict(caches=caches, q=q, kn=kn, vn=vn, out=out, rope=rope, scale=scale)
        return rec, state

    def measure_e2e(self, block_fn, steps):
        """block_fn(n) issues n pipelined e2e steps.  Captured into a graph when the capture succeeds."""
        torch = self.torch
        block_fn(steps)
        torch.cuda.synchronize()
        run, graphed = (lambda: block_fn(steps)), False
        if not self.args.eager_e2e:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg, stream=side):
                    block_fn(steps)
                cg.replay()
                torch.cuda.synchronize()
                run, graphed = cg.replay, True
            except Exception as e:  # keep the eager pipeline
                sys.stderr.write(f"[bench] e2e graph capture failed, running the pipeline eagerly: {e}\n")
                torch.cuda.synchronize()
        blocks = self.timed_blocks(run, steps, min_s=MIN_TIMED_S / 2)
        s = sorted(blocks)
        return {"ms_per_step": s[len(s) // 2] / steps, "blocks": len(blocks), "cuda_graph": graphed}

    def attach_traffic(self, rec, name):
        """DRAM bytes per launch from the committed ncu capture of this workload (static, per round)."""
        rec["roofline"]["traffic"] = None
        tr = os.path.join(ROOT, "profiles", f"traffic_{name}.json")
        if self.world == 1 and os.path.exists(tr):
            try:
                d = json.load(open(tr))
                rec["roofline"]["traffic"] = d.get("dram_bytes_per_launch")
                rec["roofline"]["traffic_source"] = "static: " + d.get("source", f"profiles/traffic_{name}.json")
            except Exception:
                pass

    # ---- C5 at N > 1: kv-head-sharded single sequence
    def run_c5_sharded(self, cfg, gather, steps, warmup):
        torch, omx = self.torch, self.omx
        Hq, Hkv, D, S, B = cfg["Hq"], cfg["Hkv"], cfg["D"], cfg["S"], cfg["B"]
        tdt, es = torch.bfloat16, 2
        rope = omx.nn.Rope(D, False, 1e6, 1.0)
        scale = D ** -0.5
        eng = omx.parallel.HeadShardedDecode(Hq, Hkv, D, tdt, rope, scale, batch=B, gather=gather)
        eng.auto_wait = True  # the exchange waits on "my own signal count": capturable
        kv_bytes = 2 * B * (Hkv // self.world) * S * D * es
        need = int(math.ceil(2 * L2_BYTES / kv_bytes))
        divs = [d for d in range(1, steps + 1) if steps % d == 0 and d >= need]
        R = divs[0] if divs else steps
        R += R & 1  # the peer path double-buffers the output by step parity: keep blocks even
        caches = []
        for c in range(R):
            cache = omx.KVCache()
            cache.reserve(S + 256)
            eng.cache = cache
            for s0 in range(0, S - 1, 4096):  # replicated prompt K/V (same seeds on every rank)
                n = min(4096, S - 1 - s0)
                kk = self.row_randn(77 + s0, (B, Hkv, n, D), tdt)
                vv = self.row_randn(99 + s0, (B, Hkv, n, D), tdt)
                eng.prefill(kk, vv)
            assert cache.offset() == S - 1
            caches.append(cache)
        q, kn, vn = (self.row_randn(5 + i, (B, h, 1, D), tdt) for i, h in enumerate((Hq, Hkv, Hkv)))

        def step(i):
            eng.cache = caches[i % R]
            eng.step(q, kn, vn)
            eng.rewind(1)

        res = self.measure(step, steps + (steps & 1), warmup, graph=True)
        pins = [torch.empty_like(t, device="cpu").pin_memory() for t in (q, kn, vn)]
        for hp, t in zip(pins, (q, kn, vn)):
            hp.copy_(t)
        out_pin = torch.empty((B, Hq, 1, D), dtype=tdt).pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in pins)
        d2h = out_pin.numel() * out_pin.element_size()

        def e2e_block(n):
            for i in range(n + (n & 1)):
                q.copy_(pins[0], non_blocking=True)
                kn.copy_(pins[1], non_blocking=True)
                vn.copy_(pins[2], non_blocking=True)
                eng.cache = caches[i % R]
                o = eng.step(q, kn, vn)
                eng.rewind(1)
                out_pin.copy_(o, non_blocking=True)

        e2e = self.measure_e2e(e2e_block, steps + (steps & 1))
        ms, e2e_ms = res["ms_per_step"], e2e["ms_per_step"]
        alg_bytes = (2 * B * Hkv * S * D * es + 2 * (2 * B * Hkv * D * es)) // self.world + 2 * B * Hq * D * es
        rec = {
            "metric": "attn_decode_tokens_per_s", "unit": "tokens/s", "value": B / (ms / 1e3), "ms_per_step": ms,
            "ms_per_step_min": res["ms_per_step_min"], "blocks": res["blocks"], "steps": steps, "dtype": "bf16",
            "e2e": {"value": B / (e2e_ms / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "blocks": e2e["blocks"],
                    "cuda_graph": e2e["cuda_graph"], "how": "serial upload -> step -> download on one stream"},
            "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms / 1e3) / 1e9, "peak": self.pk["hbm"],
                         "unit": "GB/s", "frac": alg_bytes / (ms / 1e3) / 1e9 / self.pk["hbm"],
                         "peak_source": self.pk["src"], "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "per-GPU bytes of the rank's kv heads; the step is latency-bound (launch + exchange)",
                         "traffic": None},
            "run": {"kernel": res["kernel"], "cuda_graph": True, "rotate_caches": R, "gather": gather,
                    "launches_per_step": res["launches_per_block"] / (steps + (steps & 1)),
                    "l2_policy": f"per-GPU KV {kv_bytes / 1e6:.0f} MB: rotated through {R} caches (L2-cold)"},
        }
        # ---- parity: one sharded step vs the unsharded computation of the same inputs on rank 0
        eng.cache = caches[0]
        o_sh = eng.step(q, kn, vn).clone()
        krow, vrow = (t[:, :, S - 1:S].contiguous() for t in caches[0].state())
        eng.rewind(1)
        rows = [torch.empty_like(krow) for _ in range(self.world)], [torch.empty_like(vrow) for _ in range(self.world)]
        self.dist.all_gather(rows[0], krow)
        self.dist.all_gather(rows[1], vrow)
        if self.rank == 0:
            full = omx.KVCache()
            for s0 in range(0, S - 1, 4096):
                n = min(4096, S - 1 - s0)
                full.update_and_fetch(self.row_randn(77 + s0, (B, Hkv, n, D), tdt),
                                      self.row_randn(99 + s0, (B, Hkv, n, D), tdt))
            o_ref = omx.attn_decode_fused(q, kn, vn, full, rope, scale)
            kf, vf = (t[:, :, S - 1:S] for t in full.state())
            rec["parity_check"] = {
                "against": "unsharded fused step on rank 0 (same inputs)",
                "max_abs_err": float((o_sh.float() - o_ref.float()).abs().max()),
                "kv_rows_bit_exact": bool(torch.equal(torch.cat(rows[0], 1).view(torch.int16), kf.view(torch.int16)) and
                                          torch.equal(torch.cat(rows[1], 1).view(torch.int16), vf.view(torch.int16)))}
            del full
        del eng, caches
        return rec

    # ---- prefill / DiT workloads: batch items [b0, b0 + Bl), heads [h0, h0 + Hl)
    def run_prefill(self, name, cfg, items, heads, value_scale, steps, warmup):
        torch, omx = self.torch, self.omx
        Hq, Hkv, D, S = cfg["Hq"], cfg["Hkv"], cfg["D"], cfg["S"]
        f32 = cfg["dtype"] == "f32"
        tdt, es = (torch.float32, 4) if f32 else (torch.bfloat16, 2)
        h0, Hl = heads
        G = Hq // Hkv
        Hkl = max(1, Hl // G)
        Bl = len(items)
        tag = {"c3": 3, "c4": 4, "c1_prefill": 6}[name]
        q = self.rows_randn(tag + 10, items, (Hq, S, D), tdt)[:, h0:h0 + Hl].contiguous()
        k = self.rows_randn(tag + 20, items, (Hkv, S, D), tdt)[:, h0 // G:h0 // G + Hkl].contiguous()
        v = self.rows_randn(tag + 30, items, (Hkv, S, D), tdt)[:, h0 // G:h0 // G + Hkl].contiguous()
        out = torch.empty_like(q)
        scale = D ** -0.5
        mask = omx.fast.ScaledDotProductAttentionMask.Causal if cfg["causal"] else None

        def step(i):
            omx.fast.scaled_dot_product_attention(q, k, v, scale, mask, out=out)

        res = self.measure(step, steps, warmup, graph=False)
        flops = 4.0 * Bl * Hl * S * S * D * (0.5 if cfg["causal"] else 1.0)
        pins = [torch.empty_like(t, device="cpu").pin_memory() for t in (q, k, v)]
        for hp, t in zip(pins, (q, k, v)):
            hp.copy_(t)
        out_pin = torch.empty_like(out, device="cpu").pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in pins)
        d2h = out.numel() * out.element_size()

        def e2e_block(n):
            for _ in range(n):
                q.copy_(pins[0], non_blocking=True)
                k.copy_(pins[1], non_blocking=True)
                v.copy_(pins[2], non_blocking=True)
                omx.fast.scaled_dot_product_attention(q, k, v, scale, mask, out=out)
                out_pin.copy_(out, non_blocking=True)

        n_e2e = max(2, min(steps, 5))
        e2e_block(1)
        blocks = self.timed_blocks(lambda: e2e_block(n_e2e), n_e2e, min_s=0.1, max_blocks=3)
        e2e_ms = sorted(blocks)[len(blocks) // 2] / n_e2e
        ms = res["ms_per_step"]
        tf = flops / (ms / 1e3) / 1e12
        rec = {
            "metric": "attn_prefill_tflops", "unit": "TFLOP/s", "value": tf * value_scale, "ms_per_step": ms,
            "ms_per_step_min": res["ms_per_step_min"], "blocks": res["blocks"], "steps": steps, "dtype": cfg["dtype"],
            "tokens_per_s": Bl * S * value_scale / (ms / 1e3),
            "e2e": {"value": flops * value_scale / (e2e_ms / 1e3) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "blocks": len(blocks), "cuda_graph": False,
                    "how": "serial upload -> kernel -> download on one stream: PCIe-bound ("
                           f"{(h2d + d2h) / 1e6:.0f} MB per step over the host link)"},
            "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
            # the kernel is timed back to back for >= 0.5 s (power-capped clocks, like the 4 s cuBLAS run behind
            # bf16_tflops_sustained), so the sustained figure is the denominator; the burst one is beside it
            "roofline": {"bound": "tensor", "achieved": tf, "peak": self.pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": tf / self.pk["tf_sustained"], "frac_of_burst_peak": tf / self.pk["tf_burst"],
                         "achieved_best_block": flops / (res["ms_per_step_min"] / 1e3) / 1e12,
                         "frac_of_2250_nominal": tf / 2250.0,
                         "peak_source": self.pk["src"] + " cuBLAS bf16 sustained (MEASURED_PEAKS.json "
                                        "bf16_tflops_sustained; burst = bf16_tflops)",
                         "algorithmic_flops_per_launch": flops,
                         "flops_convention": "4*B*Hq*Lq*Lk*D, causal counted as half"},
            "run": {"kernel": res["kernel"], "cuda_graph": False,
                    "l2_policy": f"q+k+v+out {(2 * q.numel() + 2 * k.numel()) * es / 1e6:.0f} MB per GPU vs 126 MB L2",
                    "launches_per_step": res["launches_per_block"] / steps},
        }
        if f32:  # the FFMA pipe bounds a float32 product that has to meet 1e-4: 2 FLOP x 128 lanes x SMs x clock
            props = torch.cuda.get_device_properties(self.dev)
            peak = 2.0 * 128 * props.multi_processor_count * self.pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
            rec["roofline"] = {"bound": "fp32_ffma", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                               "achieved_best_block": flops / (res["ms_per_step_min"] / 1e3) / 1e12,
                               "peak_source": "2 x 128 FFMA lanes x SM count x max SM clock (no measured FFMA figure "
                                              "in MEASURED_PEAKS.json)",
                               "algorithmic_flops_per_launch": flops,
                               "flops_convention": "4*B*Hq*Lq*Lk*D, causal counted as half"}
        self.attach_traffic(rec, name)
        return rec, dict(q=q, k=k, v=v, out=out, scale=scale, mask=mask)

    # ---- absorbed MLA through fast::scaled_dot_product_attention (keys wider than values, one latent kv head)
    def run_mla(self, name, cfg, steps, warmup):
        torch, omx = self.torch, self.omx
        B, Hq, Hkv, D, Dv, S, L = (cfg[k] for k in ("B", "Hq", "Hkv", "D", "Dv", "S", "L"))
        tdt, es = torch.bfloat16, 2
        decode = L == 1
        kv_bytes = B * Hkv * S * (D + Dv) * es
        R = 1
        if decode and kv_bytes < 2 * L2_BYTES:
            need = int(math.ceil(2 * L2_BYTES / kv_bytes))
            divs = [d for d in range(1, steps + 1) if steps % d == 0 and d >= need]
            R = divs[0] if divs else steps
        q = self.rows_randn(71, list(range(B)), (Hq, L, D), tdt)
        ks = [self.rows_randn(72 + 10 * c, list(range(B)), (Hkv, S, D), tdt) for c in range(R)]
        vs = [self.rows_randn(73 + 10 * c, list(range(B)), (Hkv, S, Dv), tdt) for c in range(R)]
        out = torch.empty((B, Hq, L, Dv), dtype=tdt, device=self.dev)
        scale = D ** -0.5
        mask = omx.fast.ScaledDotProductAttentionMask.Causal if cfg["causal"] else None

        def step(i):
            omx.fast.scaled_dot_product_attention(q, ks[i % R], vs[i % R], scale, mask, out=out)

        res = self.measure(step, steps, warmup, graph=decode)
        ms = res["ms_per_step"]
        flops = 2.0 * B * Hq * L * S * (D + Dv) * (0.5 if cfg["causal"] else 1.0)
        alg_bytes = kv_bytes + B * Hq * L * (D + Dv) * es
        pins = [torch.empty_like(t, device="cpu").pin_memory() for t in (q, ks[0], vs[0])]
        for hp, t in zip(pins, (q, ks[0], vs[0])):
            hp.copy_(t)
        out_pin = torch.empty_like(out, device="cpu").pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in pins)
        d2h = out.numel() * out.element_size()

        def e2e_block(n):
            for _ in range(n):
                q.copy_(pins[0], non_blocking=True)
                ks[0].copy_(pins[1], non_blocking=True)
                vs[0].copy_(pins[2], non_blocking=True)
                omx.fast.scaled_dot_product_attention(q, ks[0], vs[0], scale, mask, out=out)
                out_pin.copy_(out, non_blocking=True)

        n_e2e = max(2, min(steps, 5))
        e2e_block(1)
        blocks = self.timed_blocks(lambda: e2e_block(n_e2e), n_e2e, min_s=0.1, max_blocks=3)
        e2e_ms = sorted(blocks)[len(blocks) // 2] / n_e2e
        tf = flops / (ms / 1e3) / 1e12
        gbs = alg_bytes / (ms / 1e3) / 1e9
        per_s = (B if decode else B * S) / (ms / 1e3)
        rec = {
            "metric": "mla_decode_tokens_per_s" if decode else "mla_prefill_tflops",
            "unit": "tokens/s" if decode else "TFLOP/s", "value": per_s if decode else tf,
            "ms_per_step": ms, "ms_per_step_min": res["ms_per_step_min"], "blocks": res["blocks"], "steps": steps,
            "dtype": cfg["dtype"], "tokens_per_s": per_s, "tflops": tf, "hbm_gbs": gbs,
            "e2e": {"value": ((B if decode else flops / 1e12) / (e2e_ms / 1e3)), "unit": "tokens/s" if decode else "TFLOP/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "blocks": len(blocks),
                    "cuda_graph": False,
                    "how": "serial upload of q + the whole K / V -> kernel -> download on one stream: PCIe-bound ("
                           f"{(h2d + d2h) / 1e6:.0f} MB per step over the host link; a serving loop keeps K / V resident)"},
            "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
            "run": {"kernel": res["kernel"], "cuda_graph": bool(decode),
                    "l2_policy": (f"rotating {R} K/V sets of {kv_bytes / 1e6:.0f} MB" if R > 1 else
                                  f"K/V {kv_bytes / 1e6:.0f} MB per step vs 126 MB L2"),
                    "launches_per_step": res["launches_per_block"] / steps},
        }
        if decode:
            rec["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": self.pk["hbm"], "unit": "GB/s",
                               "frac": gbs / self.pk["hbm"], "peak_source": self.pk["src"] + " copy bandwidth",
                               "algorithmic_bytes_per_launch": alg_bytes,
                               "bytes_convention": "K + V rows once (B*Hkv*S*(Dk+Dv)*2) + q + out"}
        else:
            # mma.sync path: its own measured ceiling on this chip is 552 TFLOP/s (scripts/microbench/hmma.cu,
            # profiles/r02_mma.md); the fraction printed is against the same cuBLAS figure as C3 / C4
            rec["roofline"] = {"bound": "tensor", "achieved": tf, "peak": self.pk["tf_sustained"], "unit": "TFLOP/s",
                               "frac": tf / self.pk["tf_sustained"], "frac_of_legacy_mma_sync_peak_552": tf / 552.0,
                               "achieved_best_block": flops / (res["ms_per_step_min"] / 1e3) / 1e12,
                               "peak_source": self.pk["src"] + " cuBLAS bf16 sustained",
                               "algorithmic_flops_per_launch": flops,
                               "flops_convention": "2*B*Hq*Lq*Lk*(Dk+Dv), causal counted as half"}
        return rec

    # ---- parity of the batch split (C2, C3, C4): gathered shard outputs vs the unsharded launch on rank 0
    def parity_decode(self, name, cfg, rows, st):
        torch, omx, dist = self.torch, self.omx, self.dist
        S = cfg["S"]
        cache = st["caches"][0]
        o_sh = omx.attn_decode_fused(st["q"], st["kn"], st["vn"], cache, st["rope"], st["scale"]).clone()
        krow, vrow = (t[:, :, S - 1:S].contiguous() for t in cache.state())
        cache.trim(1)
        outs = [torch.empty_like(o_sh) for _ in range(self.world)]
        krs = [torch.empty_like(krow) for _ in range(self.world)]
        vrs = [torch.empty_like(vrow) for _ in range(self.world)]
        dist.all_gather(outs, o_sh)
        dist.all_gather(krs, krow)
        dist.all_gather(vrs, vrow)
        if self.rank != 0:
            return None
        tdt = o_sh.dtype
        tag = 2
        allrows = list(range(cfg["B"]))
        full = omx.KVCache()
        full.reserve(S + 256)
        Hkv, Hq, D = cfg["Hkv"], cfg["Hq"], cfg["D"]
        for s0 in range(0, S - 1, 1024):
            n = min(1024, S - 1 - s0)
            kk = torch.stack([self.row_randn(1234 + 1009 * tag + 7919 * r + 31 * (s0 // 1024 + 1), (Hkv, n, D), tdt)
                              for r in allrows])
            vv = torch.stack([self.row_randn(4321 + 1009 * tag + 7919 * r + 31 * (s0 // 1024 + 1), (Hkv, n, D), tdt)
                              for r in allrows])
            full.update_and_fetch(kk, vv)
        q = self.rows_randn(tag + 10, allrows, (Hq, 1, D), tdt)
        kn = self.rows_randn(tag + 20, allrows, (Hkv, 1, D), tdt)
        vn = self.rows_randn(tag + 30, allrows, (Hkv, 1, D), tdt)
        o_ref = omx.attn_decode_fused(q, kn, vn, full, st["rope"], st["scale"])
        kf, vf = (t[:, :, S - 1:S] for t in full.state())
        i16 = torch.int16
        res = {"against": f"one unsharded B = {cfg['B']} fused step on rank 0 (same per-row seeds)",
               "max_abs_err": float((torch.cat(outs, 0).float() - o_ref.float()).abs().max()),
               "kv_rows_bit_exact": bool(torch.equal(torch.cat(krs, 0).view(i16), kf.contiguous().view(i16)) and
                                         torch.equal(torch.cat(vrs, 0).view(i16), vf.contiguous().view(i16)))}
        del full
        return res

    def parity_prefill(self, name, cfg, st, items, heads):
        """Sampled check: rank 0 recomputes (batch item, head) slices that OTHER ranks own with its own launch."""
        torch, omx, dist = self.torch, self.omx, self.dist
        Hq, Hkv, D, S = cfg["Hq"], cfg["Hkv"], cfg["D"], cfg["S"]
        tag = {"c3": 3, "c4": 4}[name]
        G = Hq // Hkv
        h0, Hl = heads
        # every rank contributes the first G heads of its first item
        mine = st["out"][0:1, 0:G].contiguous()
        got = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(got, mine)
        meta = torch.tensor([items[0], h0], device=self.dev, dtype=torch.int64)
        metas = [torch.empty_like(meta) for _ in range(self.world)]
        dist.all_gather(metas, meta)
        if self.rank != 0:
            return None
        worst = 0.0
        for r in range(self.world):
            b, hh = int(metas[r][0]), int(metas[r][1])
            q = self.rows_randn(tag + 10, [b], (Hq, S, D), torch.bfloat16)[:, hh:hh + G].contiguous()
            k = self.rows_randn(tag + 20, [b], (Hkv, S, D), torch.bfloat16)[:, hh // G:hh // G + 1].contiguous()
            v = self.rows_randn(tag + 30, [b], (Hkv, S, D), torch.bfloat16)[:, hh // G:hh // G + 1].contiguous()
            o = omx.fast.scaled_dot_product_attention(q, k, v, st["scale"], st["mask"])
            worst = max(worst, float((o.float() - got[r].float()).abs().max()))
        return {"against": f"rank 0 recomputes one (item, {G}-head group) slice of every rank's shard from the "
                           "per-row seeds", "max_abs_err": worst, "kv_rows_bit_exact": None}

    # ---- one workload by name
    def run(self, name, steps, warmup):
        torch = self.torch
        W, r = self.world, self.rank
        kind, cfg = WORKLOADS[name.replace("_weak", "").replace("_collective", "").replace("_paged", "")]
        cfg = dict(cfg)
        rec = None
        if name == "c2":
            start, cnt = self.omx.parallel.batch_shard(cfg["B"], W, r)
            rows = list(range(start, start + cnt))
            rec, st = self.run_decode(name, cfg, rows, 1, steps, warmup, False)
            rec["value"] = cfg["B"] / (rec["ms_per_step"] / 1e3)  # whole job: all 64 rows per (max-over-ranks) step
            rec["e2e"]["value"] = cfg["B"] / (rec["e2e"]["ms_per_step"] / 1e3)
            rec["scaling"] = "strong"
            if W > 1:
                pc = self.parity_decode(name, cfg, rows, st)
                if pc:
                    rec["parity_check"] = pc
        elif name == "c2_paged":
            start, cnt = self.omx.parallel.batch_shard(cfg["B"], W, r)
            rows = list(range(start, start + cnt))
            rec, st = self.run_decode(name, cfg, rows, 1, steps, warmup, False, paged=True)
            rec["value"] = cfg["B"] / (rec["ms_per_step"] / 1e3)
            rec["e2e"]["value"] = cfg["B"] / (rec["e2e"]["ms_per_step"] / 1e3)
            rec["scaling"] = "strong"
        elif name == "c2_weak":
            rows = list(range(r * cfg["B"], (r + 1) * cfg["B"]))
            rec, st = self.run_decode(name, cfg, rows, W, steps, warmup, False)
            rec["scaling"] = "weak"
        elif name == "c1":
            rec, st = self.run_decode(name, cfg, [0], 1, steps, warmup, False)
            rec["scaling"] = "replicas only" if W > 1 else "single GPU"
        elif name == "gqa256_decode":
            rec, st = self.run_decode(name, cfg, list(range(cfg["B"])), 1, steps, warmup, False)
            rec["scaling"] = "single GPU"
        elif name == "c5":
            if W == 1:
                rec, st = self.run_decode(name, cfg, [0], 1, steps, warmup, False)
                rec["scaling"] = "single GPU"
            else:
                # "peer" = data + flag words (omx_attn_decode_fused_sharded_ll); OMX_BENCH_C5_GATHER=peer_flags
                # selects the r01 peer stores + arrival counters for A/B runs
                rec = self.run_c5_sharded(cfg, os.environ.get("OMX_BENCH_C5_GATHER", "peer"), steps, warmup)
                rec["scaling"] = "strong"
        elif name == "c5_collective":
            rec = self.run_c5_sharded(cfg, "collective", steps, warmup)
            rec["scaling"] = "strong"
        elif name == "c1_prefill":
            rec, st = self.run_prefill(name, cfg, [0], (0, cfg["Hq"]), 1, steps, warmup)
            rec["scaling"] = "single GPU"
        elif kind == "mla":
            rec = self.run_mla(name, cfg, steps, warmup)
            rec["scaling"] = "single GPU"
        elif name in ("c3", "c4"):
            B, Hq = cfg["B"], cfg["Hq"]
            if W <= B:
                start, cnt = self.omx.parallel.batch_shard(B, W, r)
                items, heads = list(range(start, start + cnt)), (0, Hq)
            else:  # more GPUs than batch items: split the heads of an item as well (MHA / whole GQA groups)
                per = W // B
                Hl = Hq // per
                items, heads = [r // per], ((r % per) * Hl, Hl)
            rec, st = self.run_prefill(name, cfg, items, heads, 1, steps, warmup)
            # whole job = all items per (max-over-ranks) step
            full_flops = 4.0 * B * Hq * cfg["S"] * cfg["S"] * cfg["D"] * (0.5 if cfg["causal"] else 1.0)
            rec["value"] = full_flops / (rec["ms_per_step"] / 1e3) / 1e12
            rec["tokens_per_s"] = B * cfg["S"] / (rec["ms_per_step"] / 1e3)
            rec["e2e"]["value"] = full_flops / (rec["e2e"]["ms_per_step"] / 1e3) / 1e12
            rec["scaling"] = "strong"
            if W > 1:
                pc = self.parity_prefill(name, cfg, st, items, heads)
                if pc:
                    rec["parity_check"] = pc
        rec["config"] = workload_config(name, W)
        rec["n_gpus"] = W
        st = None
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        return rec


def main():
    global MIN_TIMED_S
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="all", choices=["all", "c2_paged", "c2_weak", "c5_collective"] + sorted(WORKLOADS))
    ap.add_argument("--impl", default="omx", choices=["omx", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--eager-e2e", action="store_true", help="run the e2e pipeline eagerly instead of from a graph")
    ap.add_argument("--batch", type=int, default=None, help="override C2's GLOBAL batch (sweeps / debugging)")
    ap.add_argument("--min-seconds", type=float, default=MIN_TIMED_S,
                    help="device time to accumulate per workload (default 0.5 s); profiler runs pass 0 (3 blocks)")
    args = ap.parse_args()
    MIN_TIMED_S = max(0.0, args.min_seconds)
    args.warmup = max(args.warmup, 3)
    args.steps = max(args.steps, 2)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.batch:
        WORKLOADS["c2"][1]["B"] = args.batch

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    t_start = time.time()
    b = Bench(args)
    names = ["c2"] if args.workload in ("all", "c2") else [args.workload]
    if args.workload == "all":
        names += ["c2_paged", "c1", "c5"] + (["c5_collective"] if world > 1 and 8 % world == 0 else []) + ["c3", "c4"]
        if world == 1:
            names += ["c1_prefill", "gqa256_decode", "mla_decode", "mla_prefill"]
        if world > 1:
            names.append("c2_weak")
    recs = {}
    for n in names:
        try:
            recs[n] = b.run(n, args.steps, args.warmup)
        except Exception as e:  # a failing side workload must not take the headline down (and is visible)
            if n == names[0]:
                raise
            recs[n] = {"error": f"{type(e).__name__}: {e}"}
            b.torch.cuda.synchronize()
            b.torch.cuda.empty_cache()

    head = names[0]
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        kind = WORKLOADS[head][0]
        try:
            r = cpu_record(head, 3, 1, 20.0 if kind == "decode" else 10.0)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            if args.workload == "all":
                r1 = cpu_record("c1", 5, 1, 3.0)
                recs["c1"]["cpu_baseline"] = {k: r1[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the checker must never take the product number down with it
            cpu = {"value": None, "unit": recs[head]["unit"], "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        h = recs[head]
        line = {
            "metric": h["metric"], "value": h["value"], "unit": h["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": h["ms_per_step"], "higher_is_better": True,
            "scaling": h["scaling"] if h["scaling"] in ("strong", "weak") else "weak",
            "vs_baseline": None, "dtype": h["dtype"], "data": "synthetic", "config": h["config"],
            "e2e": h["e2e"], "gpu_launches": h["gpu_launches"], "clocks": h["clocks"], "roofline": h["roofline"],
            "cpu_baseline": cpu, "run": h["run"], "blocks": h["blocks"], "ms_per_step_min": h["ms_per_step_min"],
            "timing": f"median of {h['blocks']} blocks of {args.steps} steps (>= {MIN_TIMED_S} s timed), CUDA events, "
                      "max over ranks",
        }
        if "parity_check" in h:
            line["parity_check"] = h["parity_check"]
        if len(names) > 1:
            line["workloads"] = {n: recs[n] for n in names[1:]}
        line["wall_s"] = round(time.time() - t_start, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        b.dist.destroy_process_group()


if __name__ == "__main__":
    main()
