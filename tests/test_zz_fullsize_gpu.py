"""Last file of the suite on purpose.  Full BASELINE.json frame sizes on the GPU against the committed sha256 of the
strict oracle's frame (tests/golden/fullsize_sha256.json, written by tools/fullsize_host_check.py, which also ran the
generated program on the host at that size): the whole frame, every bit."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_ir, load_tex

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN, "fullsize_sha256.json")) as f:
    FULL = {k: v for k, v in json.load(f).items() if not k.startswith("_")}


@pytest.mark.parametrize("scene", sorted(FULL))
def test_full_frame_hash_equals_the_oracles(scene, torch_cuda):
    from portal_b200.renderer import SceneRenderer
    cfg = FULL[scene]
    r = SceneRenderer(load_ir(scene), textures=load_tex(scene), device=0)
    r.render_depth = cfg["depth"]
    img = r.render_host(cfg["width"], cfg["height"])
    assert img.dtype == np.float32 and img.shape == (cfg["height"], cfg["width"], 4)
    assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == cfg["sha256_f32_rgba"]
