#!/usr/bin/env python
"""Every reference scene that compiles: generated program run on the host (tests/host_harness) vs the strict oracle at a
small frame, bit for bit.  A tool (minutes), not a test; needs the reference checkout for the .ron files and textures.

    python tools/all_scenes_host_check.py /root/reference [--size 96x54] [--depth 30] [--only a,b]
"""
import argparse
import glob
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
from oracle import frontend  # noqa: E402
from oracle.runner import Oracle  # noqa: E402
from portal_b200.capi import PortalB200Error  # noqa: E402
from portal_b200.renderer import SceneRenderer  # noqa: E402

HH = os.path.join(ROOT, "tests", "host_harness")


def one(job):
    path, reference, w, h, depth = job
    from PIL import Image
    name = os.path.basename(path)[:-4]
    args = argparse.Namespace(reference=reference, depth=depth)
    ok = None
    if True:
        t0 = time.time()
        try:
            ir = frontend.scene_ir(frontend.load_scene(path), name)
            tex = {}
            for k, t in enumerate(ir["textures"]):
                if t.get("video") or not os.path.exists(os.path.join(args.reference, t["path"])):
                    # a video sampler (the clips are not in the reference checkout) or an image the checkout lacks
                    # (scenes/img/text.png of boot.dev): a synthetic image, the same on both sides
                    yy, xx = np.mgrid[0:36, 0:64]
                    tex[t["name"]] = np.ascontiguousarray(np.stack([(xx * 4 + k * 40) % 256, (yy * 7 + k * 90) % 256,
                                                                    ((xx ^ yy) * 5 + k * 17) % 256, np.full_like(xx, 255)], axis=-1).astype(np.uint8))
                    continue
                img = Image.open(os.path.join(args.reference, t["path"])).convert("RGBA")
                tex[t["name"]] = np.ascontiguousarray(np.asarray(img, dtype=np.uint8))
            r = SceneRenderer(ir, device=-1)
            r.render_depth = args.depth
            d = tempfile.mkdtemp(prefix=f"pe_all_{name}_")
            open(f"{d}/prog.cu", "w").write(r.source())
            open(f"{d}/block.bin", "wb").write(r.uniform_block(w, h))
            subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", f'-DPROGRAM_FILE="{d}/prog.cu"', "-I", HH,
                            os.path.join(HH, "run_program.cpp"), "-o", f"{d}/run"], check=True, capture_output=True)
            cmd = [f"{d}/run", f"{d}/block.bin", str(w), str(h), f"{d}/out.f32"]
            for t in ir["textures"]:
                tex[t["name"]].tofile(f"{d}/{t['name']}.rgba")
                cmd += [f"{d}/{t['name']}.rgba", str(tex[t["name"]].shape[1]), str(tex[t["name"]].shape[0])]
            subprocess.run(cmd, check=True, timeout=1200)
            got = np.fromfile(f"{d}/out.f32", dtype=np.float32).reshape(h, w, 4)
            want = Oracle(ir, "strict", textures=tex).render(w, h, args.depth)
            ok = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
            status = "bit-identical" if ok else f"DIFFERS in {(got.view(np.uint32) != want.view(np.uint32)).any(axis=-1).sum()} pixels"
        except (PortalB200Error, NotImplementedError, subprocess.CalledProcessError, KeyError, FileNotFoundError) as e:
            status = "skipped: " + str(e).strip().splitlines()[0][:120]
        return f"{name:40s} {time.time() - t0:6.1f} s  {status}", ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference")
    ap.add_argument("--size", default="96x54")
    ap.add_argument("--depth", type=int, default=30)
    ap.add_argument("--only", default="")
    ap.add_argument("--jobs", type=int, default=8)
    args = ap.parse_args()
    w, h = (int(x) for x in args.size.split("x"))
    only = {x for x in args.only.split(",") if x}
    jobs = [(p, args.reference, w, h, args.depth) for p in sorted(glob.glob(os.path.join(args.reference, "scenes", "*.ron")))
            if os.path.basename(p) != "empty.ron" and (not only or os.path.basename(p)[:-4] in only)]
    import concurrent.futures as cf
    same = diff = skipped = 0
    with cf.ProcessPoolExecutor(args.jobs) as ex:
        for line, ok in ex.map(one, jobs):
            print(line, flush=True)
            same += ok is True
            diff += ok is False
            skipped += ok is None
    print(f"{same} scenes bit-identical, {diff} differ, {skipped} skipped")
    return 0 if diff == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
