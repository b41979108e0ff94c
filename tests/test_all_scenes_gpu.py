"""SURVEY.md section 8 (f1) on hardware: EVERY reference scene both generators accept (81 of 82) rendered on the B200 through the
C ABI and held, bit for bit, to the strict oracle's frame -- by sha256 of the float frame, so that the 81 checkers need not
travel (tools/export_all_scenes.py wrote tests/golden/scenes_all/ from /root/reference in the build container).  Scenes that
call exp / log / pow are in here like the rest: those functions are part of the pinned numeric profile.
On a mismatch the oracle is built on the spot to say how many pixels differ."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

ALL = os.path.join(GOLDEN, "scenes_all")
with open(os.path.join(ALL, "sha256.json")) as _f:
    PINS = json.load(_f)
W, H, DEPTH = PINS["width"], PINS["height"], PINS["depth"]


def load_scene(name):
    with gzip.open(os.path.join(ALL, f"{name}.scene.json.gz"), "rb") as f:
        ir = json.loads(f.read().decode())
    tex = {}
    if ir["textures"]:
        with np.load(os.path.join(ALL, "textures.npz")) as z:
            for t in ir["textures"]:
                tex[t["name"]] = np.ascontiguousarray(z[t["path"].replace("/", "__").replace(".", "_")])
    return ir, tex


@pytest.mark.parametrize("name", sorted(PINS["frames"]))
def test_reference_scene_bit_exact_on_the_gpu(name, torch_cuda):
    from portal_b200.renderer import SceneRenderer
    ir, tex = load_scene(name)
    r = SceneRenderer(ir, textures=tex, device=0)
    r.render_depth = DEPTH
    img = r.render_host(W, H)
    assert r.launch_count() >= 1
    r.close()
    got = hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()
    if got != PINS["frames"][name]:
        from oracle import runner
        ref = runner.Oracle(ir, "strict", textures=tex).render(W, H, DEPTH)
        bad = (np.ascontiguousarray(img).view(np.uint32) != np.ascontiguousarray(ref).view(np.uint32)).any(axis=-1)
        with np.errstate(invalid="ignore"):
            err = np.nanmax(np.abs(img - ref))
        raise AssertionError(f"{name}: {int(bad.sum())} of {bad.size} pixels differ from the oracle (max |diff| {err}); "
                             f"oracle frame sha matches the pin: {hashlib.sha256(ref.tobytes()).hexdigest() == PINS['frames'][name]}")
