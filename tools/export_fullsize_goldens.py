#!/usr/bin/env python
"""sha256 pins of the oracle's frames at the FULL BASELINE.json sizes -> tests/golden/fullsize_sha256.json.

    python tools/export_fullsize_goldens.py [scene ...]

For every config scene: the saved-camera frame at the config's size and depth, as raw float32 RGBA (`sha256_f32_rgba`) and
as the RGBA8 frame a render target holds (`sha256_rgba8`: clamp, x255, round to nearest even -- numpy restatement of
pe_k_quantize_rgba8 / GL's unorm conversion).  For config 5 (mobius_monoportal 7680x4320 depth 64, 360-frame orbit) also
orbit frames 0 and 180 (`orbit_360`).  The oracle's OpenMP build is used (same bits as the strict build: contraction is off
in both; tests/test_oracle.py holds them equal), so an 8K frame takes seconds per core-minute instead of an hour.
bench.py compares the frames it timed with these pins (`parity` in its JSON line); tests/test_zz_fullsize_gpu.py holds the
GPU to them."""
import hashlib
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
from conftest import load_ir, load_tex  # noqa: E402
from oracle.runner import Oracle  # noqa: E402
from portal_b200.renderer import camera_scale, orbit_camera_matrix  # noqa: E402

CONFIGS = {"basics": (256, 256, 4), "monoportal": (1920, 1080, 20), "triple_portal": (3840, 2160, 40),
           "portal_in_portal": (3840, 2160, 40), "mobius_monoportal": (7680, 4320, 64)}
EXTRA_SIZES = {"mobius_monoportal": [(3840, 2160, 64)]}      # the size the 1-GPU sweeps use
ORBIT = {"mobius_monoportal": (360, [0, 180])}


def quantize(img):
    return np.rint(np.fmin(np.fmax(img, np.float32(0)), np.float32(1)) * np.float32(255)).astype(np.uint8)


def pins(img):
    return {"sha256_f32_rgba": hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest(),
            "sha256_rgba8": hashlib.sha256(np.ascontiguousarray(quantize(img)).tobytes()).hexdigest()}


def main():
    path = os.path.join(ROOT, "tests", "golden", "fullsize_sha256.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for scene in (sys.argv[1:] or list(CONFIGS)):
        w, h, d = CONFIGS[scene]
        ir = load_ir(scene)
        orc = Oracle(ir, "fast", textures=load_tex(scene))
        t0 = time.time()
        rec = {"width": w, "height": h, "depth": d, **pins(orc.render(w, h, d))}
        print(f"{scene} {w}x{h} d{d}: {time.time() - t0:.1f} s  f32 {rec['sha256_f32_rgba'][:16]}  rgba8 {rec['sha256_rgba8'][:16]}", flush=True)
        for (w2, h2, d2) in EXTRA_SIZES.get(scene, []):
            rec[f"{w2}x{h2}"] = {"depth": d2, **pins(orc.render(w2, h2, d2))}
        if scene in ORBIT:
            n, ks = ORBIT[scene]
            cam = ir["cam"]
            rec[f"orbit_{n}"] = {}
            for k in ks:
                # BASELINE config 5: alpha_k = alpha_0 + 2 pi k / n; only `_camera` changes
                m = orbit_camera_matrix(cam["look_at"], cam["alpha"] + 2.0 * math.pi * k / n, cam["beta"], cam["r"])
                t0 = time.time()
                rec[f"orbit_{n}"][str(k)] = pins(orc.render(w, h, d, camera=m, camera_scale=camera_scale(m)))
                print(f"{scene} orbit frame {k}/{n}: {time.time() - t0:.1f} s  f32 {rec[f'orbit_{n}'][str(k)]['sha256_f32_rgba'][:16]}", flush=True)
        out[scene] = rec
    out["_note"] = ("sha256 of the oracle's frame at the BASELINE.json size: raw float32 RGBA (row-major) and the RGBA8 render-target "
                    "quantisation of it; orbit_<n>: frame k of the n-frame camera orbit of config 5; written by "
                    "tools/export_fullsize_goldens.py (tools/fullsize_host_check.py ran the generated program on the host at these sizes)")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
