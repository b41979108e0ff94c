"""Last file of the suite on purpose.  Full BASELINE.json frame sizes on the GPU against the committed sha256 of the
strict oracle's frame (tests/golden/fullsize_sha256.json, written by tools/fullsize_host_check.py, which also ran the
generated program on the host at that size): the whole frame, every bit.  Also here: the strips of ph_render_target, and the
analytic scenes of tests/test_analytic.py on the GPU (oracle bit for bit, closed forms to rounding)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_ir, load_tex

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN, "fullsize_sha256.json")) as f:
    FULL = {k: v for k, v in json.load(f).items() if not k.startswith("_")}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("scene", sorted(FULL))
def test_full_frame_hash_equals_the_oracles(scene, torch_cuda):
    """Every BASELINE config at its full size (config 5: 7680x4320 depth 64), float frame and RGBA8 frame, through the JSON
    scene IR route."""
    from portal_b200.renderer import SceneRenderer
    cfg = FULL[scene]
    r = SceneRenderer(load_ir(scene), textures=load_tex(scene), device=0)
    r.render_depth = cfg["depth"]
    img = r.render_host(cfg["width"], cfg["height"])
    assert img.dtype == np.float32 and img.shape == (cfg["height"], cfg["width"], 4)
    assert _sha(img) == cfg["sha256_f32_rgba"]
    del img
    assert _sha(r.render_host_rgba8(cfg["width"], cfg["height"])) == cfg["sha256_rgba8"]
    r.close()


@pytest.mark.parametrize("scene", sorted(FULL))
def test_full_frame_through_the_products_own_front_end(scene, torch_cuda):
    """The same frames through `.ron` -> C++ host front-end (ph_scene_*, include/portal_b200_host.h) -> pe_*: the route
    bench.py's e2e leg measures.  Equal to the oracle's pin, hence bit-identical to the JSON-IR route above."""
    from conftest import ROOT
    from portal_b200.host import HostRenderer, HostScene
    cfg = FULL[scene]
    w, h, d = cfg["width"], cfg["height"], cfg["depth"]
    hs = HostScene.from_file(os.path.join(ROOT, "tests", "golden", "ron", f"{scene}.ron"))
    hr = HostRenderer(hs, device=0, textures=load_tex(scene))
    assert _sha(hr.render_frame(w, h, d, rgba8=True)) == cfg["sha256_rgba8"]
    if w * h <= 3840 * 2160:
        assert _sha(hr.render_frame(w, h, d)) == cfg["sha256_f32_rgba"]
    for n, frames in ((int(k.split("_")[1]), v) for k, v in cfg.items() if k.startswith("orbit_")):
        cam = hs.camera()
        for k, pin in frames.items():
            if int(k) == 0:
                continue                                   # frame 0 is the saved camera (checked above)
            import math
            c = dict(cam, alpha=cam["alpha"] + 2.0 * math.pi * int(k) / n)
            assert _sha(hr.render_frame(w, h, d, camera=c, rgba8=True)) == pin["sha256_rgba8"], (scene, k)
    hr.close()


def test_host_render_target_strips_reassemble_the_frame(torch_cuda):
    """ph_render_target (C++ host: uniform setup + asynchronous strip render into device memory): the strips of three
    "ranks" put back in row order equal ph_render_frame's frame."""
    torch = torch_cuda
    from conftest import ROOT
    from portal_b200.distributed import local_rows
    from portal_b200.host import HostRenderer, HostScene
    from portal_b200.renderer import SceneRenderer
    fixture = os.path.join(ROOT, "tests", "fixtures", "two_spheres.ron")
    hr = HostRenderer(HostScene.from_file(fixture), device=0)
    w, h, depth, world = 160, 100, 12, 3
    whole = hr.render_frame(w, h, depth)
    got = np.zeros_like(whole)
    for rank in range(world):
        t = SceneRenderer.strip_target(w, h, 16, rank, world)
        rows = local_rows(h, rank, world, 16)
        buf = torch.zeros((len(rows), w, 4), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()                                  # the fill runs on torch's stream, the render on the context's
        hr.render_target(t, buf.data_ptr(), depth)
        hr.sync()
        part = buf.cpu().numpy()
        for k, y in enumerate(rows):
            if y >= 0:
                got[y] = part[k]
    assert np.array_equal(got.view(np.uint32), whole.view(np.uint32))


def test_kernel_on_the_analytic_scenes_equals_oracle_and_closed_form(torch_cuda):
    """tests/fixtures/analytic{,2,...,7}.ron on the GPU with the identity camera: bit-identical to the oracle's frame,
    and within rounding of the closed forms written down from the reference's shader text (tests/test_analytic.py)."""
    from conftest import ROOT
    from oracle import frontend, runner
    from portal_b200.renderer import SceneRenderer
    from test_analytic import DEPTH, H, IDENTITY, W, closed_form, closed_form2
    for name, form in (("analytic", closed_form), ("analytic2", closed_form2), ("analytic3", None), ("analytic4", None), ("analytic5", None), ("analytic6", None), ("analytic7", None)):
        ir = frontend.scene_ir(frontend.load_scene(os.path.join(ROOT, "tests", "fixtures", name + ".ron")), name)
        ref = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0)
        r = SceneRenderer(ir, device=0)
        r.render_depth = DEPTH
        r.camera_matrix = list(IDENTITY)
        img = r.render_host(W, H)
        assert np.array_equal(np.ascontiguousarray(img).view(np.uint32), np.ascontiguousarray(ref).view(np.uint32)), name
        if form is not None:
            want, safe = form(W, H)[:2]
            assert np.abs(img[..., :3].astype(np.float64) - want)[safe].max() < 2e-5, name
