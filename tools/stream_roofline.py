#!/usr/bin/env python
"""HBM roofline of the three streaming kernels around the ray loop (run on the GPU box).

    python tools/stream_roofline.py [WxH] [launches]

pe_k_quantize_rgba8   20 B/pixel (16 read + 4 written)              -- main.rs:2939-2943's RGBA8 target
pe_k_deinterleave     32 B/pixel (16 + 16)                          -- SURVEY.md 8e
pe_k_average_rgba8    4 n + 4 B/pixel (n RGBA8 frames -> one)        -- main.rs:640-722 average_images
Timed with CUDA events on the launching stream over `launches` back-to-back launches whose buffers rotate through a ring
larger than the 126 MB L2; `frac` is against MEASURED_PEAKS.json's hbm_gbs (else the B200_PROFILING.md fallback).
Prints one JSON line per kernel.  Results are also checked against numpy restatements."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from portal_b200 import capi  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


def timed(fn, n, stream):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(n):
        fn(i)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    w, h = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "3840x2160").split("x"))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    lib = capi.lib()
    ctx = lib.pe_create(0)
    assert ctx, lib.pe_last_error(None)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    pk, src = peak()
    npx = w * h
    gen = torch.Generator(device="cuda").manual_seed(1)

    def check(rc):
        assert rc == 0, lib.pe_last_error(ctx).decode()

    def report(name, bpp, ms, ok, note=""):
        gbs = bpp * npx / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "frame": f"{w}x{h}", "bytes_per_pixel": bpp, "ms": round(ms, 5), "achieved_gbs": round(gbs, 1),
                          "peak_gbs": pk, "peak_source": src, "frac": round(gbs / pk, 4), "result_matches_numpy": bool(ok), "note": note}), flush=True)

    # ---- quantize: ring of 3 float frames (3 x 132.7 MB at 4K) + 3 RGBA8 frames
    fin = [torch.rand((h, w, 4), generator=gen, device="cuda") * 1.2 - 0.1 for _ in range(3)]
    fin[0][0, 0, 0] = float("nan")
    q8 = [torch.empty((h, w, 4), dtype=torch.uint8, device="cuda") for _ in range(3)]
    ms = timed(lambda i: check(lib.pe_quantize_rgba8(ctx, fin[i % 3].data_ptr(), q8[i % 3].data_ptr(), npx, sp)), n, stream)
    x = fin[0].cpu().numpy()
    want = np.rint(np.fmin(np.fmax(x, np.float32(0)), np.float32(1)) * np.float32(255)).astype(np.uint8)
    check(lib.pe_quantize_rgba8(ctx, fin[0].data_ptr(), q8[0].data_ptr(), npx, sp))
    torch.cuda.synchronize()
    report("pe_k_quantize_rgba8", 20, ms, np.array_equal(q8[0].cpu().numpy(), want))

    # ---- deinterleave: 8 ranks x cyclic 16-row strips
    from portal_b200.distributed import deinterleave_numpy, strips_per_rank
    world, srows = 8, 16
    spr = strips_per_rank(h, world, srows)
    gath = [torch.rand((world, spr, srows, w, 4), generator=gen, device="cuda") for _ in range(2)]
    frames = [torch.empty((h, w, 4), device="cuda") for _ in range(2)]
    ms = timed(lambda i: check(lib.pe_deinterleave_strips(ctx, gath[i % 2].data_ptr(), frames[i % 2].data_ptr(), w, h, srows, world, spr, sp)),
               n, stream)
    check(lib.pe_deinterleave_strips(ctx, gath[0].data_ptr(), frames[0].data_ptr(), w, h, srows, world, spr, sp))
    torch.cuda.synchronize()
    ok = np.array_equal(frames[0].cpu().numpy(), deinterleave_numpy(gath[0].cpu().numpy(), h, world, srows))
    report("pe_k_deinterleave", 32, ms, ok, f"{world} ranks x {srows}-row strips")
    del gath, frames, fin

    # ---- average: n_frames RGBA8 sub-frames -> one (ring of 2 sets)
    for nf in (4, 16):
        sets = [[torch.randint(0, 256, (h, w, 4), generator=gen, dtype=torch.uint8, device="cuda") for _ in range(nf)] for _ in range(2)]
        outs = [torch.empty((h, w, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
        arrs = [(C.c_void_p * nf)(*[t.data_ptr() for t in s]) for s in sets]
        ms = timed(lambda i: check(lib.pe_average_frames_rgba8(ctx, arrs[i % 2], nf, outs[i % 2].data_ptr(), npx, sp)), n, stream)
        check(lib.pe_average_frames_rgba8(ctx, arrs[0], nf, outs[0].data_ptr(), npx, sp))
        torch.cuda.synchronize()
        st = np.stack([t.cpu().numpy().astype(np.uint32) for t in sets[0]])
        mean = (st * st).sum(axis=0) // nf
        want = (np.sqrt(mean.astype(np.float32)) + np.float32(0.5)).astype(np.uint8)
        want[..., 3] = 255
        report("pe_k_average_rgba8", 4 * nf + 4, ms, np.array_equal(outs[0].cpu().numpy(), want), f"{nf} sub-frames")
        del sets, outs
    lib.pe_destroy(ctx)


if __name__ == "__main__":
    main()
