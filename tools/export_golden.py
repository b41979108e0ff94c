#!/usr/bin/env python
"""Regenerate the committed golden fixtures from the reference's scene files.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tools/export_golden.py

Writes, for each BASELINE.json config scene:
  tests/golden/scenes/<scene>.scene.json    scene IR: objects / materials / snippets (the
                                            scene's own GLSL, i.e. benchmark *input data*), the
                                            evaluated uniform table (f64), saved camera
  tests/golden/scenes/<scene>.textures.npz  decoded RGBA8 texels of the scene's textures
  tests/golden/frames/<scene>_<W>x<H>_d<D>.npz   oracle (strict f32) framebuffer + sha256
The front-end used is the ORACLE's (oracle/frontend.py); the product's C++ front-end is
checked against these files by tests/test_host_frontend.py.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import frontend, runner  # noqa: E402

REF = "/root/reference"
SCENES = ["basics", "monoportal", "triple_portal", "portal_in_portal", "mobius_monoportal"]
# (W, H, depth): reduced-size pins; depth as in BASELINE.json configs
FRAMES = {
    "basics": (256, 256, 4),
    "monoportal": (160, 90, 20),
    "triple_portal": (160, 90, 40),
    "portal_in_portal": (160, 90, 40),
    "mobius_monoportal": (160, 90, 64),
}


# Further reference scenes kept as parity inputs for the GPU tests (beyond the five
# configs).  Scene IR only (compact JSON); their single texture, where they have one, is the config
# scenes' scenes/img/monoportal.png, already stored in monoportal.textures.npz.
EXTRA = ["borromean_rings", "hopf_link", "recursive_space", "mobius", "sphere_intersection", "cone", "recursive_room",
         "spherical_geometry", "cylinder", "non_linear", "time_portal_spacetime", "matryoshka"]


# Scenes checked on the host harness only so far (tests/test_program_on_host.py): same IR format, separate folder so that
# the GPU parametrisation (scenes_extra) grows only with scenes that have been seen on a GPU.
HOST_ONLY = ["trefoil"]


def export_extra():
    for folder, names in (("scenes_extra", EXTRA), ("scenes_host", HOST_ONLY)):
        _export_ir(os.path.join(ROOT, "tests/golden", folder), names)


def _export_ir(d, names):
    os.makedirs(d, exist_ok=True)
    for name in names:
        ir = frontend.scene_ir(frontend.load_scene(f"{REF}/scenes/{name}.ron"), name)
        assert all(t["path"] == "scenes/img/monoportal.png" for t in ir["textures"]), name
        with open(os.path.join(d, f"{name}.scene.json"), "w") as f:
            json.dump(ir, f, separators=(",", ":"))
        print(f"extra {name}: {len(ir['objects'])} objects, {len(ir['uniforms'])} uniforms")


def main():
    from PIL import Image
    export_extra()

    os.makedirs(os.path.join(ROOT, "tests/golden/scenes"), exist_ok=True)
    os.makedirs(os.path.join(ROOT, "tests/golden/frames"), exist_ok=True)
    for name in SCENES:
        scene = frontend.load_scene(f"{REF}/scenes/{name}.ron")
        ir = frontend.scene_ir(scene, name)
        with open(os.path.join(ROOT, f"tests/golden/scenes/{name}.scene.json"), "w") as f:
            json.dump(ir, f, indent=1)
        tex = {}
        for t in ir["textures"]:
            im = Image.open(os.path.join(REF, t["path"])).convert("RGBA")
            tex[t["name"]] = np.asarray(im, dtype=np.uint8)
        tpath = os.path.join(ROOT, f"tests/golden/scenes/{name}.textures.npz")
        if tex:
            np.savez_compressed(tpath, **tex)
        elif os.path.exists(tpath):
            os.remove(tpath)
        w, h, d = FRAMES[name]
        orc = runner.Oracle(ir, "strict", textures=tex)
        img = orc.render(w, h, d)
        digest = hashlib.sha256(img.tobytes()).hexdigest()
        np.savez_compressed(os.path.join(ROOT, f"tests/golden/frames/{name}_{w}x{h}_d{d}.npz"), frame=img,
                            sha256=np.array(digest))
        print(f"{name}: {len(ir['objects'])} objects, {len(ir['uniforms'])} uniforms, frame {w}x{h} d{d} sha256 {digest[:16]}")


if __name__ == "__main__":
    main()
