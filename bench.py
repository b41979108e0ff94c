#!/usr/bin/env python
"""bench.py -- Mpixels/s of the portal ray loop on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one frame of the headline workload (portal_in_portal, 3840x2160, depth 40, saved
camera, aa 1): uniform upload + ONE launch of the scene's sm_100a ray-loop kernel (+ for N > 1 the
NCCL gather of the row strips and the de-interleave kernel on rank 0).  `value` is timed with CUDA
events on the launching stream with the scene resident on the GPU; `e2e` goes through the
reference-facing call (render_frame's product: RGBA8 pixels in HOST memory) and includes the
host->device uniform upload and the device->host readback every step.
`--impl reference` times the CPU oracle (oracle/, the restatement of the reference's GLSL path --
the reference has no CPU implementation and cannot be built here) on the host
cores, on a bounded sample of the same frame.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENE_DIR = os.path.join(ROOT, "tests", "golden", "scenes")
WORKLOADS = {  # BASELINE.json configs
    "portal_in_portal": (3840, 2160, 40),
    "triple_portal": (3840, 2160, 40),
    "monoportal": (1920, 1080, 20),
    "mobius_monoportal": (7680, 4320, 64),
    "basics": (256, 256, 4),
}
STRIP_ROWS = 16
METRIC = "Mpixels/s @ 3840x2160 depth-40 portal_in_portal"  # BASELINE.json metric (headline workload)


def metric_name(scene, w, h, depth):
    return f"Mpixels/s @ {w}x{h} depth-{depth} {scene}"


def load_ir(scene):
    with open(os.path.join(SCENE_DIR, f"{scene}.scene.json")) as f:
        return json.load(f)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.path = f"/tmp/_pe_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None, t2=None):
        """t0, t1: wall-clock (time.time()) bounds of the timed region.  If it was too short to contain two samples
        (nvidia-smi cannot sample faster than ~20 ms) the samples up to t2 -- the end of an untimed repeat of the very
        same steps that run_ours() appends in that case -- are used, and failing that everything since start()."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        import datetime
        rows = []
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(p[2]), float(p[3]), p[6:10]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        window = "timed region"
        if len(inside) < 2 and t2 is not None:
            inside = [r for r in rows if t0 <= r[0] <= t2]
            window = "timed region + untimed repeat of the same steps (timed region shorter than the sampling period)"
        if len(inside) < 2:
            inside, window = rows, "warm-up + timed region (timed region shorter than the sampling period)"
        sm = sorted(r[1] for r in inside)
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": inside[0][2] if inside else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def usable_cores() -> int:
    """Host threads this process may really use: CPU affinity capped by the cgroup CPU quota (a container
    can report 128 CPUs and be throttled to a few)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(math.ceil(int(txt[0]) / int(txt[1])))))
            else:
                q = int(txt[0])
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, int(math.ceil(q / per))))
            break
        except Exception:
            continue
    return n


# ------------------------------------------------------------------------------ CPU oracle legs
def cpu_oracle_rate(scene, w, h, depth, budget_s=20.0, threads=0):
    """Mpx/s of the oracle's fast build on a bounded sample: bands of 16 rows spread over the frame."""
    from oracle import runner
    tex = runner.load_texture_npz(os.path.join(SCENE_DIR, f"{scene}.textures.npz"))
    orc = runner.Oracle(load_ir(scene), "fast", textures=tex)
    cores = threads or usable_cores()
    n_bands = 4
    t0 = time.perf_counter()
    B = min(64, h)
    orc.render(w, h, depth, rows=(h // 2, min(h, h // 2 + B)), threads=cores)  # warm up the thread pool
    t0 = time.perf_counter()
    orc.render(w, h, depth, rows=(h // 2, min(h, h // 2 + B)), threads=cores)  # calibrate
    per_band = max(time.perf_counter() - t0, 1e-4)
    n_bands = int(max(2, min(h // B, budget_s / per_band)))
    starts = [int(i * (h - B) / max(n_bands - 1, 1)) for i in range(n_bands)]
    t0 = time.perf_counter()
    px = 0
    for s in starts:
        orc.render(w, h, depth, rows=(s, s + B), threads=cores)
        px += B * w
    dt = time.perf_counter() - t0
    return px / dt / 1e6, cores, f"{n_bands} bands x {B} rows of the {w}x{h} frame ({px} px, {dt:.1f} s, OpenMP over 64-pixel runs)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, h, depth = args.width, args.height, args.depth
    rates = []
    for _ in range(min(args.warmup, 3)):
        cpu_oracle_rate(args.scene, w, h, depth, budget_s=0.5)
    t0 = time.perf_counter()
    sample = ""
    for _ in range(args.steps):
        # bounded sample per step: the whole arm stays around 1.5 minutes whatever --steps is
        r, cores, sample = cpu_oracle_rate(args.scene, w, h, depth, budget_s=min(20.0, max(0.25, 90.0 / max(args.steps, 1))))
        rates.append(r)
    dt = time.perf_counter() - t0
    v = sum(rates) / len(rates)
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args.scene, w, h, depth), "value": round(v, 4), "unit": "Mpixels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(w * h / (v * 1e6) * 1e3, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.scene}.ron {w}x{h} depth {depth}, " + (f"{args.orbit}-frame camera orbit" if args.orbit else "saved camera") + ", aa 1"},
        "cpu_baseline": {"value": round(v, 4), "unit": "Mpixels/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 4), "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": f"CPU oracle (restatement of the reference's GLSL path; the reference itself has no CPU path); wall {dt:.1f} s",
    }))


# ------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from portal_b200.renderer import SceneRenderer, load_textures

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h, depth = args.width, args.height, args.depth
    ir = load_ir(args.scene)
    r = SceneRenderer(ir, textures=load_textures(os.path.join(SCENE_DIR, f"{args.scene}.textures.npz")), device=local,
                      persistent=bool(args.persistent))
    r.render_depth = depth
    stream = torch.cuda.Stream()          # a real (non-NULL) stream: NULL means "the context's own stream" to the C ABI
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0

    from portal_b200.distributed import FrameSharder, STRIP_ROWS as SR
    mode = args.mode if world > 1 else "gather"
    if world > 1 and mode == "gather":
        args.format = "f32"                      # the NCCL gather path assembles float frames
    if world == 1:
        target = r.full_target(w, h)
        # ring of output frames larger than the 126 MB L2: 2 float frames (265 MB at 4K) or 6 RGBA8 frames (199 MB)
        n_outs = 2 if args.format == "f32" else 6
        outs = [torch.empty((h, w, 4), dtype=torch.float32 if args.format == "f32" else torch.uint8, device="cuda") for _ in range(n_outs)]
        sharder = None
    else:
        # default for N > 1: the render kernels store straight into rank 0's frame over NVLink (CUDA IPC).
        # If the box cannot map peer memory between processes, every rank agrees to use the NCCL gather
        # instead (both are GPU paths; the mode actually used is reported in `config.parallelism`).
        sharder, ok = None, 1
        try:
            sharder = FrameSharder(r, w, h, rank, world, mode=mode, strip_rows=SR, fmt=args.format)
        except Exception as e:  # noqa: BLE001
            ok = 0
            print(f"[bench] rank {rank}: {mode} set-up failed ({e}); falling back to gather", file=sys.stderr)
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if sharder is not None:
                sharder.close()
            mode = "gather"
            args.format = "f32"                  # the gather path assembles float frames
            sharder = FrameSharder(r, w, h, rank, world, mode=mode, strip_rows=SR)
        target = sharder.target

    cam0 = dict(r.cam)
    frame_counter = [0]

    def step(i, consumer=None):
        """One frame.  `consumer` (rank 0) enqueues whatever reads the assembled frame; after it the frame's
        buffer is released to the other ranks (p2p mode)."""
        _step_render(i)
        if consumer is not None:
            consumer()
        if sharder is not None:
            sharder.release(sptr)

    def _step_render(i):
        if args.orbit:
            # BASELINE config 5: alpha_k = alpha_0 + 2*pi*k/orbit; only `_camera` changes
            k = frame_counter[0] % args.orbit
            frame_counter[0] += 1
            r.set_cam(cam0["look_at"], cam0["alpha"] + 2.0 * math.pi * k / args.orbit, cam0["beta"], cam0["r"])
        if world == 1:
            if args.format == "f32":
                r.draw_texture(target, outs[i & 1].data_ptr(), 0, sptr)
            else:
                r.draw_texture_rgba8(target, outs[i % len(outs)].data_ptr(), sptr)
        else:
            sharder.render(i, sptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # nvidia-smi needs ~0.2 s to produce its first sample: start it early
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    # ---- timed region: K steps, device time, max over ranks
    l0 = r.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_epoch0 = time.time()
    ev[0].record(stream)
    for i in range(args.steps):
        kev[i][0].record(stream)
        step(i)
        kev[i][1].record(stream)
    ev[1].record(stream)
    barrier()
    t_epoch1 = time.time()
    launches = r.launch_count() - l0
    dev_ms = ev[0].elapsed_time(ev[1])
    step_ms = dev_ms      # both modes are stream-ordered end to end: device time between the two events
    ms = torch.tensor([step_ms], dtype=torch.float64, device="cuda")
    kms = torch.tensor([sum(a.elapsed_time(b) for a, b in kev) / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(ms.item()), float(kms.item())
    # A timed region of a few tens of ms (many GPUs, short frames) ends before nvidia-smi delivers two samples: keep the
    # GPUs under the identical load, untimed, for ~1.5 s so that the clocks / throttle reasons are observed under it.
    # The count is derived from the all-reduced time, so every rank runs the same number of steps.
    if total_ms < 100.0:
        extra = max(1, min(50000, int(1500.0 / max(total_ms / args.steps, 1e-3))))
        for i in range(extra):
            step(args.steps + i)
        barrier()
    clocks = sampler.stop(t_epoch0, t_epoch1, time.time()) if rank == 0 else None

    # ---- e2e: the reference-facing call with HOST buffers (RGBA8 readback = get_texture_data)
    e2e_steps = max(3, min(args.steps, 20))
    from portal_b200.distributed import gpu_numa_affinity
    with gpu_numa_affinity(local):      # pinned pages on the GPU's NUMA node: the D2H copy stays off the socket interconnect
        host8 = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() if rank == 0 else None
        host8_b = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() if (rank == 0 and world == 1) else None
    q8 = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda") if (rank == 0 and world > 1) else None
    cam = r.cam
    n_px_local = int(r._lib.pe_target_pixels(C.byref(target))) if world == 1 or mode == "gather" else \
        sum(1 for y in __import__("portal_b200.distributed", fromlist=["x"]).local_rows(h, rank, world, SR) if y >= 0) * w

    def e2e_step(i):
        r.set_cam(cam["look_at"], cam["alpha"], cam["beta"], cam["r"])   # per-frame host work: camera -> _camera
        if world == 1:
            r.render_host_ptr(w, h, host8.data_ptr(), rgba8=True)
        else:
            def consume():
                if rank == 0 and args.format == "f32":
                    r._check(r._lib.pe_quantize_rgba8(r._ctx, sharder.frame_ptr, q8.data_ptr(), w * h, sptr))
                    host8.copy_(q8, non_blocking=True)
                elif rank == 0:                  # the assembled frame is RGBA8 already
                    stream.synchronize()
                    r._check(r._lib.pe_memcpy_d2h(r._ctx, host8.data_ptr(), sharder.frame_ptr, w * h * 4, sptr))
            step(i, consume)
            torch.cuda.synchronize()
    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_rate = w * h * e2e_steps / float(e2e_s.item()) / 1e6
    e2e_sync_rate = None
    if world > 1:
        # host delivery without a device-side gather: every rank renders its strips as RGBA8 and copies them over its
        # own PCIe link into one shared, page-locked host frame (HostFrameSharder); rank 0 consumes whole frames
        from portal_b200.distributed import HostFrameSharder
        try:
            hfs = HostFrameSharder(r, w, h, rank, world)     # raises on EVERY rank alike if /dev/shm is too small
        except RuntimeError as e:
            hfs = None
            if rank == 0:
                print(f"[bench] {e}; e2e stays the blocking rank-0 path", file=sys.stderr)
    if world > 1 and hfs is not None:
        e2e_sync_rate = e2e_rate
        def pipelined_n(n):
            prev = None
            for i in range(n):
                r.set_cam(cam["look_at"], cam["alpha"], cam["beta"], cam["r"])
                f = hfs.submit()
                if prev is not None:
                    hfs.complete(prev)
                    if rank == 0:
                        hfs.wait_frame(prev)          # the whole frame i-1 is in host memory: the consumer may read it
                        hfs.release(prev)
                prev = f
            hfs.complete(prev)
            if rank == 0:
                hfs.wait_frame(prev)
                hfs.release(prev)
        pipelined_n(3)
        barrier()
        t0 = time.perf_counter()
        pipelined_n(e2e_steps)
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_rate = w * h * e2e_steps / float(e2e_s.item()) / 1e6
        hfs.close()
    if world == 1:
        # the frame-sequence form of the same call (the offline `render` loop reads one frame after another):
        # pe_submit_host_rgba8 / pe_wait_host, frame i's D2H overlapping frame i+1's kernel; every frame still
        # uploads its uniforms and lands, complete, in pinned host memory before the clock stops
        e2e_sync_rate = e2e_rate
        ring = [host8, host8_b]
        def pipelined(n):
            prev = None
            for i in range(n):
                r.set_cam(cam["look_at"], cam["alpha"], cam["beta"], cam["r"])
                tk = r.submit_host_rgba8(w, h, ring[i % 2].data_ptr())
                if prev is not None:
                    r.wait_host(prev)            # frame i-1 is in host memory: the consumer may read it now
                prev = tk
            r.wait_host(prev)
        pipelined(3)
        r.sync()
        t0 = time.perf_counter()
        pipelined(e2e_steps)
        e2e_rate = w * h * e2e_steps / (time.perf_counter() - t0) / 1e6

    if rank == 0:
        peaks, peak_src = measured_peaks()
        value = w * h * args.steps / (total_ms * 1e-3) / 1e6
        # roofline of the dominant kernel (pe_render_kernel): algorithmic bytes = 16 B/pixel written
        #  x pixels one launch shades, / its mean launch duration (CUDA events)
        bpp = 16.0 if args.format == "f32" else 4.0
        alg_bytes = bpp * n_px_local
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(f"{args.scene}_{w}x{h}_d{depth}") if args.format == "f32" else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_oracle_rate(args.scene, w, h, depth, budget_s=20.0)
            v1, _, _ = cpu_oracle_rate(args.scene, w, h, depth, budget_s=4.0, threads=1)
            cpu = {"value": round(v, 4), "unit": "Mpixels/s", "cores": cores, "kind": "port", "sample": sample,
                   "single_thread_value": round(v1, 4)}
        line = {
            "metric": metric_name(args.scene, w, h, depth), "value": round(value, 2), "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.scene}.ron {w}x{h} depth {depth}, " + (f"{args.orbit}-frame camera orbit" if args.orbit else "saved camera") + ", aa 1" + (" (BASELINE.json headline config)" if (args.scene, w, h, depth) == ("portal_in_portal", 3840, 2160, 40) else ""),
                       "parallelism": "1 GPU" if world == 1 else (
                           f"{world} GPUs x cyclic {STRIP_ROWS}-row strips + 1 NCCL gather + de-interleave" if mode == "gather" else
                           f"{world} GPUs x cyclic {STRIP_ROWS}-row strips, kernels store into rank 0's frame over NVLink (CUDA IPC), stream-ordered flag words, no collective"),
                       "scheduler": "persistent warps + per-bounce refill" if args.persistent else "one thread per pixel, 8x4 warp tiles, 512-thread blocks, <= 64 regs",
                       "frame_format": "float RGBA (16 B/pixel)" if args.format == "f32" else "RGBA8 quantised by the kernel (4 B/pixel)",
                       "l2": ("each step writes a 132.7 MB frame (> 126 MB L2) into alternating buffers" if args.format == "f32" else
                              ("each step writes a 33.2 MB frame into a ring of 6 (199 MB > 126 MB L2)" if world == 1 else
                               "each step writes a 33.2 MB frame into one of rank 0's two buffers (L2-resident: this format is a bandwidth experiment)")) +
                             "; inputs are a <8 KB constant block"},
            "kernel_ms": round(kernel_ms, 4),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": round(achieved / peaks["hbm_gbs"], 5), "traffic": traffic, "peak_source": peak_src,
                         "note": f"{int(bpp)} B/pixel algorithmic; the loop is fp32-ALU/latency bound, see DESIGN.md §6"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_rate, 2), "unit": "Mpixels/s", "h2d_bytes_per_step": e2e_h2d_bytes(r),
                    "d2h_bytes_per_step": w * h * 4, "steps": e2e_steps,
                    "call": "pe_submit_host_rgba8 / pe_wait_host per frame (RGBA8 into pinned host memory, 2 frames in flight)" if world == 1 else
                            (f"pe_submit_host_strips_rgba8 / pe_wait_host on every rank: RGBA8 strips over {world} PCIe links into one shared pinned host frame, no gather"
                             if e2e_sync_rate is not None else
                             f"pe_render ({mode}) into rank 0's float frame + pe_quantize_rgba8 + D2H from rank 0, blocking per frame")},
            "gpu_launches": int(launches),
        }
        if e2e_sync_rate is not None:
            line["e2e"]["sync_call_value"] = round(e2e_sync_rate, 2)
            line["e2e"]["sync_call"] = "pe_render_host_rgba8 (one blocking call per frame)" if world == 1 else \
                f"pe_render ({mode}) into rank 0's float frame + pe_quantize_rgba8 + D2H from rank 0, blocking per frame"
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if sharder is not None:
        barrier()
        sharder.close()
    if world > 1:
        dist.destroy_process_group()


def e2e_h2d_bytes(r):
    """Bytes of the constant uniform block uploaded every step (scene matrices + camera + scalars)."""
    src = r.source()
    import re
    m = re.search(r"sizeof\(PeConstBlock\) == (\d+)", src)
    return int(m.group(1)) if m else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="portal_in_portal")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--persistent", type=int, default=0)
    ap.add_argument("--orbit", type=int, default=0, help="camera orbit of this many frames per turn (config 5: 360)")
    ap.add_argument("--mode", default="p2p", choices=["gather", "p2p"], help="N > 1: how strips reach rank 0")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--format", default="f32", choices=["f32", "rgba8"],
                    help="frame format of the timed steps: f32 = float RGBA, 16 B/pixel (the metric's definition, SURVEY.md 8d); "
                         "rgba8 = what the reference's RGBA8 render target holds, 4 B/pixel, quantised by the kernel (p2p mode for N > 1)")
    args = ap.parse_args()
    w, h, d = WORKLOADS[args.scene]
    args.width, args.height, args.depth = args.width or w, args.height or h, args.depth or d
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
