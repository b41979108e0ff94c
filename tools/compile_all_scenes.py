#!/usr/bin/env python
"""Generate + NVRTC-compile (sm_100a) the program of every scene under a directory of reference `.ron` files and,
optionally, build the CPU oracle for the same scenes.  No GPU needed.  Prints one line per scene and a summary;
this is how the "81 of the 82 non-empty reference scenes compile" figure in DESIGN.md is obtained.

    python tools/compile_all_scenes.py /root/reference/scenes [--oracle] [--only name,name]
"""
import argparse
import glob
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from portal_b200.capi import PortalB200Error  # noqa: E402
from portal_b200.host import HostRenderer, HostScene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene_dir")
    ap.add_argument("--oracle", action="store_true", help="also generate and build the CPU oracle (g++)")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = {x for x in args.only.split(",") if x}
    ok = bad = 0
    for path in sorted(glob.glob(os.path.join(args.scene_dir, "*.ron"))):
        name = os.path.basename(path)[:-4]
        if name == "empty" or (only and name not in only):
            continue
        t0 = time.time()
        try:
            hs = HostScene.from_file(path)
            hr = HostRenderer(hs, device=-1)       # compile-only context: program generation + NVRTC
            status = f"cuda ok ({len(hr.source()) // 1024} KiB source)"
            hr.close()
            if args.oracle:
                from oracle import frontend, gen_oracle
                gen_oracle.build(frontend.scene_ir(frontend.load_scene(path), name), "fast")
                status += ", oracle ok"
            ok += 1
        except (PortalB200Error, NotImplementedError, RuntimeError) as e:
            bad += 1
            status = "FAILED: " + str(e).strip().splitlines()[0][:160]
        print(f"{name:40s} {time.time() - t0:5.1f} s  {status}", flush=True)
    print(f"{ok} scenes compile, {bad} do not")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
