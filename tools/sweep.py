#!/usr/bin/env python
"""Kernel-time sweep over generator options (run on the GPU box).  Prints one line per combination:
scene, options, registers, kernel ms (CUDA events, mean of N launches after warm-up), Mpx/s.

    python tools/sweep.py <scene[,scene]> '<json: option -> list of values>' [launches] [WxHxDEPTH]"""
import itertools
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from portal_b200.renderer import SceneRenderer, load_scene_ir, load_textures  # noqa: E402

SCENES = {"portal_in_portal": (3840, 2160, 40), "triple_portal": (3840, 2160, 40), "monoportal": (1920, 1080, 20),
          "mobius_monoportal": (3840, 2160, 64), "basics": (1024, 1024, 4)}


def regs_of(cubin):
    p = f"/tmp/_sweep_{os.getpid()}.cubin"
    open(p, "wb").write(cubin)
    out = subprocess.run(["cuobjdump", "-res-usage", p], capture_output=True, text=True).stdout
    m = re.search(r"REG:(\d+) STACK:(\d+)", out)
    return (int(m.group(1)), int(m.group(2))) if m else (None, None)


def main():
    scenes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["portal_in_portal"]
    grid = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {"persistent": [0], "min_blocks": [1, 4, 5], "unroll_loops": [1, 0]}
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    size = tuple(int(x) for x in sys.argv[4].split("x")) if len(sys.argv) > 4 else None
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    for scene in scenes:
        w, h, d = size or SCENES[scene]
        ir = load_scene_ir(os.path.join(ROOT, "tests/golden/scenes", f"{scene}.scene.json"))
        tex = load_textures(os.path.join(ROOT, "tests/golden/scenes", f"{scene}.textures.npz"))
        out = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
        keys = list(grid)
        ref = None
        for combo in itertools.product(*[grid[k] for k in keys]):
            opts = dict(zip(keys, combo))
            pers = bool(opts.pop("persistent", 0))
            try:
                r = SceneRenderer(ir, textures=tex, device=0, persistent=pers, options=opts)
            except Exception as e:
                print(scene, dict(zip(keys, combo)), "FAILED", str(e)[:200])
                continue
            r.render_depth = d
            t = r.full_target(w, h)
            for _ in range(3):
                r.draw_texture(t, out.data_ptr(), 0, stream.cuda_stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n):
                r.draw_texture(t, out.data_ptr(), 0, stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            img = out.cpu()
            if ref is None:
                ref = img
            same = bool(torch.equal(img.view(torch.int32), ref.view(torch.int32)))
            reg, stack = regs_of(r.cubin())
            print(f"{scene:18s} {json.dumps(dict(zip(keys, combo))):60s} regs {reg} stack {stack} {ms:8.3f} ms {w*h/ms/1e3:9.1f} Mpx/s same_bits={same}",
                  flush=True)
            r.close()


if __name__ == "__main__":
    main()
