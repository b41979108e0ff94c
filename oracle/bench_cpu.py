"""CPU baseline of bench.py -- ORACLE / TEST INFRASTRUCTURE (see oracle/README.md): times the oracle's fast build, the
restatement of the reference's GLSL ray loop (the reference itself has no CPU path), on a bounded sample of one frame.

Run as a CLEAN SUBPROCESS by bench.py (`python -m oracle.bench_cpu ...`), never inside the GPU process: a process launched
by torchrun inherits OMP_NUM_THREADS=1 and, on some boxes, a narrowed CPU affinity; measured from there the same oracle gave
3x-5x different rates on the same box.  Here, before OpenMP starts:
  * the CPU affinity is reset to every CPU the cgroup allows,
  * OMP_NUM_THREADS = the thread count asked for (default: usable cores), OMP_PROC_BIND=spread, OMP_PLACES=cores,
and the line printed at the end reports the thread count an OpenMP parallel region REALLY had (`threads_used`).

    python -m oracle.bench_cpu --scene portal_in_portal --width 3840 --height 2160 --depth 40 [--threads N] [--budget 20]
prints one JSON line: {"value": Mpixels/s, "threads_used": n, "sample": "...", "pixels": p, "seconds": s}
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time


def usable_cores() -> int:
    """Host threads this process may really use: CPU affinity capped by the cgroup CPU quota."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", required=True)
    ap.add_argument("--width", type=int, required=True)
    ap.add_argument("--height", type=int, required=True)
    ap.add_argument("--depth", type=int, required=True)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--budget", type=float, default=20.0, help="seconds of CPU rendering to spend on the sample (all steps together)")
    ap.add_argument("--steps", type=int, default=1, help="split the sample into this many timed steps (the reference arm's --steps); 0 = as many passes over the frame as fit --budget")
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--scene-dir", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scenes"))
    args = ap.parse_args()
    # ---- a clean slate for OpenMP, BEFORE the oracle library (and libgomp) is loaded
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except OSError:
        pass
    threads = args.threads or usable_cores()
    if "PORTAL_B200_BENCH_CPU_REEXEC" not in os.environ:
        env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="spread" if threads > 1 else "false", OMP_PLACES="cores",
                   PORTAL_B200_BENCH_CPU_REEXEC="1")
        env.pop("KMP_AFFINITY", None)
        os.execve(sys.executable, [sys.executable, "-m", "oracle.bench_cpu", *sys.argv[1:]], env)   # libgomp reads these at load
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import runner
    w, h, depth = args.width, args.height, args.depth
    with open(os.path.join(args.scene_dir, f"{args.scene}.scene.json")) as f:
        ir = json.load(f)
    orc = runner.Oracle(ir, "fast", textures=runner.load_texture_npz(os.path.join(args.scene_dir, f"{args.scene}.textures.npz")))
    used = orc.omp_threads(threads)
    B = min(64, h)
    mid = (h // 2, min(h, h // 2 + B))
    for _ in range(max(1, args.warmup)):
        orc.render(w, h, depth, rows=mid, threads=threads)                   # warm up the thread pool
    t0 = time.perf_counter()
    orc.render(w, h, depth, rows=mid, threads=threads)                       # calibrate
    per_band = max(time.perf_counter() - t0, 1e-4)
    steps = args.steps if args.steps > 0 else max(1, int(args.budget / (per_band * max(1, h // B))))    # 0: as many whole frames as fit the budget
    per_step = int(max(1, min(h // B, args.budget / per_band / steps)))      # bands of one step: a bounded sample of the frame
    rates, px, dt = [], 0, 0.0
    for k in range(steps):
        # all steps together sample steps * per_step bands spread uniformly over the frame; step k takes every steps-th of them,
        # so one step and the whole run both see the frame's cheap and expensive rows in proportion
        total = steps * per_step
        starts = [int((k + i * steps + 0.5) / total * (h - B)) for i in range(per_step)]
        t0 = time.perf_counter()
        for s in starts:
            orc.render(w, h, depth, rows=(s, s + B), threads=threads)
        t = time.perf_counter() - t0
        rates.append(per_step * B * w / t / 1e6)
        px += per_step * B * w
        dt += t
    print(json.dumps({"value": px / dt / 1e6, "threads_used": used, "threads_requested": threads, "pixels": px, "seconds": dt, "steps": steps,
                      "step_rates_min_max": [min(rates), max(rates)],
                      "sample": f"{steps} step(s) x {per_step} bands x {B} rows spread over the {w}x{h} frame ({px} px, {dt:.1f} s; OpenMP dynamic "
                                f"schedule over 64-pixel runs, {used} threads, OMP_PROC_BIND={os.environ.get('OMP_PROC_BIND')})"}))


if __name__ == "__main__":
    main()
