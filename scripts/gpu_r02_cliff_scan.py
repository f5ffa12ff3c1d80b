"""Scan of the path's standalone ops over head dims / dtypes for performance cliffs: bytes moved per call against
the copy peak (rope, rms_norm, KV append) -- any line far below ~0.3 of the peak deserves a look."""
import importlib, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
dev = "cuda"
def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us
out = []
for dt in (torch.bfloat16, torch.float32):
    es = 2 if dt == torch.bfloat16 else 4
    for D in (64, 80, 128, 256, 576):
        # prefill-sized rows in the callers' layout: [B, L, H, D] viewed [B, H, L, D]
        B, L, H = 4, 2048, 8
        x = torch.randn((B, L, H, D), device=dev).to(dt).transpose(1, 2)
        w = torch.ones(D, device=dev, dtype=dt)
        nbytes = 2 * x.numel() * es
        us = timeit(lambda: omx.fast.rms_norm(x, w, 1e-6))
        out.append(("rms_norm", str(dt)[6:], D, f"{B}x{H}x{L}", round(us, 1), round(nbytes / us / 1e3, 1)))
        dims = D if D != 576 else 64
        us = timeit(lambda: omx.fast.rope(x, dims, False, 1e6, 1.0, 0))
        out.append(("rope", str(dt)[6:], D, f"{B}x{H}x{L} dims{dims}", round(us, 1), round(nbytes / us / 1e3, 1)))
        # decode-sized rows
        xd = torch.randn((64, 1, 32, D), device=dev).to(dt).transpose(1, 2)
        us = timeit(lambda: omx.fast.rms_norm(xd, w, 1e-6))
        out.append(("rms_norm", str(dt)[6:], D, "64x32x1", round(us, 1), None))
        us = timeit(lambda: omx.fast.rope(xd, dims, False, 1e6, 1.0, 777))
        out.append(("rope", str(dt)[6:], D, "64x32x1", round(us, 1), None))
        # KV append of a prefill chunk
        k = torch.randn((B, H, L, D), device=dev).to(dt)
        def app():
            c = omx.KVCache(); c.reserve(L); c.update_and_fetch(k, k)
        c0 = omx.KVCache(); c0.reserve(L)
        def app2():
            c0.reset(); c0.update_and_fetch(k, k)
        us = timeit(app2)
        out.append(("kv_append", str(dt)[6:], D, f"{B}x{H}x{L}", round(us, 1), round(4 * k.numel() * es / us / 1e3, 1)))
for r in out:
    print(json.dumps({"op": r[0], "dtype": r[1], "D": r[2], "shape": r[3], "us": r[4], "GB/s": r[5]}), flush=True)
