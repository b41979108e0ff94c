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
    ap.add_argument("--budget", type=float, default=20.0, help="seconds of CPU rendering to spend on the sample")
    ap.add_argument("--scene-dir", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scenes"))
    args = ap.parse_args()
    # ---- a clean slate for OpenMP, BEFORE the oracle library (and libgomp) is loaded
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except OSError:
        pass
    threads = args.threads or usable_cores()
    if "OMP_NUM_THREADS_SET_BY_BENCH" not in os.environ:
        env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="spread" if threads > 1 else "false", OMP_PLACES="cores",
                   OMP_NUM_THREADS_SET_BY_BENCH="1")
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
    orc.render(w, h, depth, rows=mid, threads=threads)                       # warm up the thread pool
    t0 = time.perf_counter()
    orc.render(w, h, depth, rows=mid, threads=threads)                       # calibrate
    per_band = max(time.perf_counter() - t0, 1e-4)
    n_bands = int(max(2, min(h // B, args.budget / per_band)))
    starts = [int(i * (h - B) / max(n_bands - 1, 1)) for i in range(n_bands)]
    t0 = time.perf_counter()
    px = 0
    for s in starts:
        orc.render(w, h, depth, rows=(s, s + B), threads=threads)
        px += B * w
    dt = time.perf_counter() - t0
    print(json.dumps({"value": px / dt / 1e6, "threads_used": used, "threads_requested": threads, "pixels": px, "seconds": dt,
                      "sample": f"{n_bands} bands x {B} rows spread over the {w}x{h} frame ({px} px, {dt:.1f} s; OpenMP dynamic schedule over "
                                f"64-pixel runs, {used} threads, OMP_PROC_BIND={os.environ.get('OMP_PROC_BIND')})"}))


if __name__ == "__main__":
    main()
