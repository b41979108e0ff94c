#!/usr/bin/env python
"""Every reference scene as a parity input for the GPU.

Run in the build container only (needs /root/reference):

    python tools/export_all_scenes.py [--jobs 8]

For each scene of /root/reference/scenes that needs no video clip (81 of 82; `boot.dev` samples four, which the checkout does not hold) writes
  tests/golden/scenes_all/<scene>.scene.json.gz   scene IR (the ORACLE's front-end; tests/test_host_frontend.py holds the
                                                  product's C++ front-end to the same tables)
  tests/golden/scenes_all/textures.npz            decoded RGBA8 texels, one entry per distinct image file, down-sampled
                                                  copies are NOT made: only the images these scenes reference
  tests/golden/scenes_all/sha256.json             sha256 of the strict oracle's float frame at SIZE, depth DEPTH, saved camera
tests/test_all_scenes_gpu.py renders the same frames on the B200 and compares hashes (bit-exact parity)."""
import argparse
import concurrent.futures as cf
import glob
import gzip
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "scenes_all")
W, H, DEPTH = 192, 108, 30


def tex_key(path: str) -> str:
    return path.replace("/", "__").replace(".", "_")


def one(path):
    import numpy as np
    from PIL import Image
    from oracle import frontend, runner
    from portal_b200.capi import PortalB200Error
    from portal_b200.renderer import SceneRenderer
    name = os.path.basename(path)[:-4]
    try:
        ir = frontend.scene_ir(frontend.load_scene(path), name)
        tex, used = {}, {}
        if ir.get("videos"):          # the clips are not in the reference checkout: checked on the host with synthetic frames instead
            return name, None, None, "skipped: samples video textures (tools/all_scenes_host_check.py, tests/test_video_textures.py)"
        for t in ir["textures"]:
            img = np.ascontiguousarray(np.asarray(Image.open(os.path.join(REF, t["path"])).convert("RGBA"), dtype=np.uint8))
            tex[t["name"]] = img
            used[tex_key(t["path"])] = img
        SceneRenderer(ir, device=-1).close()          # the sm_100a generator + NVRTC accept it too
        frame = runner.Oracle(ir, "strict", textures=tex).render(W, H, DEPTH)
    except (PortalB200Error, NotImplementedError, KeyError, FileNotFoundError) as e:
        return name, None, None, "skipped: " + str(e).strip().splitlines()[0][:140]
    with gzip.GzipFile(os.path.join(OUT, f"{name}.scene.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(ir, separators=(",", ":")).encode())
    return name, hashlib.sha256(frame.tobytes()).hexdigest(), used, "ok"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=8)
    args = ap.parse_args()
    import numpy as np
    os.makedirs(OUT, exist_ok=True)
    paths = [p for p in sorted(glob.glob(os.path.join(REF, "scenes", "*.ron"))) if os.path.basename(p) != "empty.ron"]
    shas, textures = {}, {}
    with cf.ProcessPoolExecutor(args.jobs) as ex:
        for name, sha, used, status in ex.map(one, paths):
            print(f"{name:40s} {status} {(sha or '')[:16]}", flush=True)
            if sha:
                shas[name] = sha
                textures.update(used)
    with open(os.path.join(OUT, "sha256.json"), "w") as f:
        json.dump({"width": W, "height": H, "depth": DEPTH, "what": "sha256 of the strict oracle's float32 RGBA frame, saved camera",
                   "frames": shas}, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, "textures.npz"), **textures)
    print(f"{len(shas)} scenes exported, {len(paths) - len(shas)} skipped, {len(textures)} textures")


if __name__ == "__main__":
    main()
