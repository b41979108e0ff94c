#!/usr/bin/env python
"""Full-size cross-check WITHOUT a GPU: the generated scene program run on the host (tests/host_harness) against the
strict oracle at the BASELINE.json frame size, bit for bit, plus the sha256 of the frame.  Slow (minutes); a tool,
not a test.  `python tools/fullsize_host_check.py portal_in_portal [width height]`"""
import hashlib
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
from conftest import DEPTH, load_ir, load_tex  # noqa: E402
from oracle.runner import Oracle  # noqa: E402
from portal_b200.renderer import SceneRenderer  # noqa: E402


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "portal_in_portal"
    w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160)
    ir = load_ir(scene)
    r = SceneRenderer(ir, device=-1)
    r.render_depth = DEPTH[scene]
    d = tempfile.mkdtemp(prefix="pe_fullsize_")
    open(os.path.join(d, "prog.cu"), "w").write(r.source())
    open(os.path.join(d, "block.bin"), "wb").write(r.uniform_block(w, h))
    hh = os.path.join(ROOT, "tests", "host_harness")
    subprocess.run(["g++", "-std=c++20", "-O2", "-fopenmp", "-ffp-contract=off", f'-DPROGRAM_FILE="{d}/prog.cu"', "-I", hh,
                    os.path.join(hh, "run_program.cpp"), "-o", f"{d}/run"], check=True)
    args = [f"{d}/run", f"{d}/block.bin", str(w), str(h), f"{d}/out.f32"]
    tex = load_tex(scene) or {}
    for t in ir["textures"]:
        arr = np.ascontiguousarray(tex[t["name"]], dtype=np.uint8)
        arr.tofile(f"{d}/{t['name']}.rgba")
        args += [f"{d}/{t['name']}.rgba", str(arr.shape[1]), str(arr.shape[0])]
    t0 = time.time()
    subprocess.run(args, check=True)
    t1 = time.time()
    got = np.fromfile(f"{d}/out.f32", dtype=np.float32).reshape(h, w, 4)
    want = Oracle(ir, "strict", textures=tex).render(w, h, DEPTH[scene])
    t2 = time.time()
    same = np.array_equal(got.view(np.uint32), want.view(np.uint32))
    print(f"{scene} {w}x{h} depth {DEPTH[scene]}: program-on-host {t1 - t0:.1f} s, oracle {t2 - t1:.1f} s, bit-identical: {same}, "
          f"sha256 {hashlib.sha256(want.tobytes()).hexdigest()}")
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
