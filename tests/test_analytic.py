"""Analytic pixel checks of the oracle.

The reference has no pixel pins, so the oracle is tied to the reference's shader TEXT here by a second, independent route:
for the hand-written scenes tests/fixtures/analytic{,2,...,7}.ron every pixel of a frame has a closed form that can be
written down from frag.glsl / library.glsl / the generator (scene.rs:720-1063) by hand -- no ray loop, no generated code,
float64 numpy -- and the oracle's frame must agree with it to rounding (2e-5; the oracle computes in float32) everywhere
except within a hair of a decision boundary.  The generated sm_100a program, run on the host, must equal the oracle bit
for bit on the same inputs, and tests/test_zz_fullsize_gpu.py holds the GPU to both.  Covered:

  * pixel -> ray: R2 sample offsets, row order, AA mean of linear colours + sqrt, 60 degree / Panini / 360 / VR180 maps,
    side-by-side stereo, camera scale (scene.rs:1688-1693, frag.glsl:408-526, 550);
  * objects: Flat/Simple, Flat/Portal (jump, offset step, tmul under scaling, rotation, both `back` polarities and the way
    back), Complex/Simple (sphere under a scaling matrix; library triangle() / cylinder()), Complex/Portal (spherical gate),
    DebugMatrix (axis capsules), an intersection material competing with objects, ignored teleport codes, subspaces;
  * shading: material_simple2's angle term and the three grid patterns, Reflect, Refract, miss colour and skybox lookup,
    darkening (start, ramp, clamp, camera scale), depth-map ramp;
  * uniforms: time -> formula -> matrix -> block -> pixels; the camera-teleportation probe's known answers."""
import os

import numpy as np
import pytest

from conftest import ROOT

SCENE = os.path.join(ROOT, "tests", "fixtures", "analytic.ron")
W, H, DEPTH = 96, 54, 8
IDENTITY = [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0]


def closed_form(w, h, s=1.0, sample=0, linear=False, projection=None, depth_map=None, stereo=None, grid="grid", sky=None, near_shift=0.0, far_z=30.0, see_far_directly=True, cs=1.0, near_is_gridded=False, no_near=False, eye=0.0):
    """(frame [h, w, 3] float64, mask of pixels further than a hair from every decision boundary, region masks).
    s: scale of the gate's far side (gate_b) -- the jump then magnifies by s about the gate's centre, the offset step is taken
    along the UN-normalised direction (length s) and normalize_ray leaves tmul = 1 / s (library.glsl:108-113, 366-371)."""
    px, py = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    m = float(min(w, h))
    # scene.rs:1688-1693 + frag.glsl:515-526 with aa_count = 1, aa_start = 0: uv + R2(0) * pixel_size * 2, R2(0) = (.5, .5)
    # sample i sits at R2(i) = fract(0.5 + (0.7548776662, 0.5698402910) * i) (frag.glsl:506-513)
    r2x, r2y = np.mod(0.5 + 0.7548776662 * sample, 1.0), np.mod(0.5 + 0.5698402910 * sample, 1.0)
    a = ((px + 0.5) - w / 2) / m * 2 + r2x * (1 / m) * 2
    b = ((py + 0.5) - h / 2) / m * 2 + r2y * (1 / m) * 2
    ex = eye                                                 # the camera sits at (eye, 0, 0) (an eye of anaglyph stereo)
    split = np.zeros((h, w), dtype=bool)
    if stereo is not None:
        # frag.glsl:476-497: side by side -- each half of the frame is its own image for one eye, here at (-+stereo, 0, 0)
        posx, posy = a / 2 * m + w / 2, b / 2 * m + h / 2
        left = posx < w / 2
        split = np.abs(posx - w / 2) < 1e-6
        m2 = float(min(w / 2, h))
        a = np.where(left, posx - w / 4, posx - w / 2 - w / 4) / m2 * 2
        b = (posy - h / 2) / m2 * 2
        ex = np.where(left, -stereo, stereo)
    # frag.glsl:449-455 with the identity camera and view angle 90 degrees: d = normalize(a * tan(45), b * tan(45), 1);
    # another projection hands in its own direction field: everything below only needs a = dx/dz, b = dy/dz, n = 1/dz of
    # rays going forward, the pixels whose ray goes backwards (they see nothing) and the pixels outside the image (black)
    backwards = black = edge = np.zeros((h, w), dtype=bool)
    ray = None
    if projection is not None:
        dx, dy, dzz, black, edge = projection(a, b)
        ray = (dx, dy, dzz)
        backwards = ~(dzz > 1e-6)
        dzs = np.where(backwards, 1.0, dzz)
        a, b = dx / dzs, dy / dzs
    n = np.sqrt(a * a + b * b + 1)
    dz = 1 / n
    out = np.empty((h, w, 3))
    safe = ~edge & ~split                                      # |position| equal to the image border / the eye split up to rounding

    def near_boundary(x, edge, eps=2e-3):
        return np.abs(x - edge) < eps

    # gate: plane z = 3, inside the unit circle -> jump by (+100, 0, 0), step 2e-5 along d, carry on to the wall at z = 30
    gx, gy = ex + 3 * a, 3 * b
    in_gate = gx * gx + gy * gy < 1
    safe &= ~near_boundary(gx * gx + gy * gy, 1)
    t1 = 3 * n
    step = s * 2e-5                                          # r.o += r.d * offset with |r.d| = s, before the normalisation
    t2 = (far_z - (3 + step * dz)) * n                       # from the stepped origin to the far wall, along the unit direction
    u, v = s * gx + a * dz * (step + t2), s * gy + b * dz * (step + t2)   # far wall's local (x, y): the jump moved x by exactly +100
    all_t = t1 + t2 / s                                      # all_t += t * r.tmul (frag.glsl:119, 125)
    def gridded(u, v, all_t, where):
        """material_simple2 of the far wall (angle term, color_grid) + the darkening, for hits at local (u, v)."""
        nonlocal safe
        green = np.array([0.2, 0.9, 0.5])
        c = green * (1 - 0.25) + green * dz[..., None] * 0.25    # color_add_weighted(c, c * |cos|, normal_coef); cos = d.z here
        if grid == "grid":
            fu, fv = np.mod(u * 1.0 * 0.25, 1.0), np.mod(v * 1.0 * 0.25, 1.0)     # color_grid: fract(uv * grid_scale * 0.25)
            sx, sy = (fu <= 0.5).astype(np.float64), (fv <= 0.5).astype(np.float64)   # step(edge = uv, x = 0.5) = 0.5 < uv ? 0 : 1
            for f in (fu, fv):
                safe &= ~(where & (near_boundary(f, 0.5) | near_boundary(f, 0.0) | near_boundary(f, 1.0)))
            low, high = 0.7 + (1.1 - 0.7) * sx, 1.1 + (0.7 - 1.1) * sx
            factor = low + (high - low) * sy
        elif grid == "grid2":                                     # color_grid2 / circle_sdf, library.glsl:199-211
            sxy = np.array([2.0, np.sqrt(3.0) * 2.0])
            pos = np.stack([u, v], axis=-1) / sxy
            d1 = (np.mod(pos, 1.0) - 0.5) * sxy
            d2 = (np.mod(pos + 0.5, 1.0) - 0.5) * sxy
            dd = np.sqrt(np.minimum((d1 * d1).sum(axis=-1), (d2 * d2).sum(axis=-1))) - 1.0
            factor = np.where(dd < -0.2, 1.1, 0.7)
            safe &= ~(where & near_boundary(dd, -0.2, 5e-3))
        else:                                                     # color_grid3, library.glsl:268-283
            qx, qy = np.mod(u * 0.5, 1.0) - 0.5, np.mod(v * 0.5, 1.0) - 0.5
            dist = np.maximum(np.abs(qx), np.abs(qy)) * 2
            factor = np.where(dist > 0.985, 0.4, np.where(dist < 0.94, 1.0, np.where(qx > qy, 0.7, 1.2)))
            safe &= ~(where & (near_boundary(dist, 0.985, 5e-3) | near_boundary(dist, 0.94, 5e-3) |
                               ((dist >= 0.94) & (dist <= 0.985) & near_boundary(qx - qy, 0.0, 5e-3))))
        c = c * (1 - 0.3) + (c * factor[..., None]) * 0.3
        # frag.glsl:131-141 with camera_scale cs: darker only beyond _t_start * cs, clamped at _t_end * cs
        gray = np.where(all_t > 10.0 * cs, (np.minimum(all_t, 210.0 * cs) - 10.0 * cs) / 200.0 / cs, 0.0)
        safe &= ~(where & (np.abs(all_t - 10.0 * cs) < 1e-3))
        return c * ((1 - gray) ** 4)[..., None]
    far = gridded(u, v, all_t, in_gate)
    # wide-angle projections also see the far wall directly (x in 60..140 at z = 30), and could graze the gate's far disc
    direct = ~in_gate & (np.abs(ex + 30 * a - 100) < 40) & (np.abs(30 * b) < 40) & see_far_directly
    safe &= ~(~in_gate & (near_boundary(np.abs(ex + 30 * a - 100), 40, 0.05) | near_boundary(np.abs(30 * b), 40, 0.05)))
    safe &= ~((np.abs(ex + 3 * a - 100) < 1.5) & (np.abs(3 * b) < 1.5))
    far_direct = gridded(ex + 30 * a - 100, 30 * b, 30 * n, direct)

    # near wall: plane z = 6, |x| < 4, -2 < y < 3.5 (asymmetric in y: pins the row order), reached only outside the gate
    nx, ny = ex + 6 * a - near_shift, 6 * b                  # the near wall's own x: it may be moved along x
    on_near = (np.abs(nx) < 4) & (ny > -2) & (ny < 3.5) & ~in_gate & (not no_near)
    safe &= ~(~in_gate & (near_boundary(np.abs(nx), 4) | near_boundary(ny, -2) | near_boundary(ny, 3.5)))
    red = np.array([0.8, 0.4, 0.2])
    near = red * (1 - 0.5) + red * dz[..., None] * 0.5
    near_t = 6 * n                                           # darkened only if the camera scale pulls _t_start below it
    near = near * ((1 - np.where(near_t > 10.0 * cs, (np.minimum(near_t, 210.0 * cs) - 10.0 * cs) / 200.0 / cs, 0.0)) ** 4)[..., None]
    safe &= ~(on_near & (np.abs(near_t - 10.0 * cs) < 1e-3))
    if near_is_gridded:
        near = gridded(nx, ny, near_t, on_near)
    miss = np.full(3, 0.6 * 0.6)                             # current_color (1) * color(0.6, 0.6, 0.6), scene.rs:1060
    if sky is not None:
        # scene.rs:1052-1058: with a skybox the miss colour is the texture looked up by the direction of the CAMERA ray
        # (`_camera_mul_inv` = identity here): u = atan(z, x), v = atan(|xz|, y), at ((u / pi + 1) / 2, v / pi), squared
        rx, ry, rz = ray if ray is not None else (a / n, b / n, 1 / n)
        su = (np.arctan2(rz, rx) / np.pi + 1) / 2
        sv = np.arctan2(np.sqrt(rx * rx + rz * rz), ry) / np.pi
        th, tw = sky.shape[:2]
        tx, ty = su * tw - 0.5, sv * th - 0.5                  # pinned sampling rule: texel centres at (i + .5) / size, bilinear,
        x0, y0 = np.floor(tx), np.floor(ty)                    # clamp to edge (oracle/glsl_compat.h `texture`)
        fx, fy = (tx - x0)[..., None], (ty - y0)[..., None]
        tex = sky[..., :3].astype(np.float64) / 255
        ix0, ix1 = np.clip(x0, 0, tw - 1).astype(int), np.clip(x0 + 1, 0, tw - 1).astype(int)
        iy0, iy1 = np.clip(y0, 0, th - 1).astype(int), np.clip(y0 + 1, 0, th - 1).astype(int)
        top = tex[iy0, ix0] * (1 - fx) + tex[iy0, ix1] * fx
        bot = tex[iy1, ix0] * (1 - fx) + tex[iy1, ix1] * fx
        miss = (top * (1 - fy) + bot * fy) ** 2
        safe &= ~((rx < 0) & (np.abs(rz) < 2e-2))              # the seam of atan(z, x)

    in_gate, on_near, direct = in_gate & ~backwards & ~black, on_near & ~backwards & ~black, direct & ~backwards & ~black
    if depth_map is not None:
        # frag.glsl:80-98, 130, 457-463: depth = all_t / camera_scale of the final hit, coloured by the inferno ramp at
        # 1 - clamp((depth - min) / (max - min)); rays that end on nothing are black
        lo, hi = depth_map
        depth = np.where(in_gate, all_t, np.where(on_near, 6 * n, 30 * n))
        tt = 1 - np.clip((depth - lo) / max(1e-6, hi - lo), 0, 1)
        stops = np.array([[0.001462, 0.000466, 0.013866], [0.258234, 0.038571, 0.406485], [0.578304, 0.148039, 0.404411],
                          [0.865006, 0.316822, 0.226055], [0.987622, 0.645320, 0.039886], [0.988362, 0.998364, 0.644924]]) ** 2
        seg = np.minimum((tt / 0.2).astype(int), 4)
        for edge_t in (0.2, 0.4, 0.6, 0.8):
            safe &= ~near_boundary(tt, edge_t, 1e-4)
        f = ((tt - 0.2 * seg) / 0.2)[..., None]
        ramp = stops[seg] * (1 - f) + stops[seg + 1] * f
        near = far = far_direct = ramp
        miss = np.zeros(3)
    out[:] = miss
    out[direct] = far_direct[direct]
    out[on_near] = near[on_near]
    out[in_gate] = far[in_gate]
    out[black] = 0.0
    if projection is not None:
        safe &= ~(np.abs(dzz) < 1e-3)
    return (out if linear else np.sqrt(out)), safe, in_gate, on_near


def closed_form2(w, h):
    """tests/fixtures/analytic2.ron: sphere (centre (0,0,8), radius 2 = unit sphere under a scale-2 matrix), mirror plane
    z = 12, plain wall z = -4 behind the camera seen only in the mirror."""
    px, py = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    m = float(min(w, h))
    a, b = (px + 1 - w / 2) * 2 / m, (py + 1 - h / 2) * 2 / m
    n = np.sqrt(a * a + b * b + 1)
    d = np.stack([a / n, b / n, 1 / n], axis=-1)
    disc = 64 / (n * n) - 60                                  # (d.c)^2 - (|c|^2 - radius^2), c = (0, 0, 8)
    on_ball = disc > 0
    safe = np.abs(disc) > 2e-2
    t = 8 / n - np.sqrt(np.where(on_ball, disc, 0.0))
    nrm = (d * t[..., None] - np.array([0.0, 0.0, 8.0])) / 2  # unit; adjugate of a uniform scale keeps its direction
    cos = np.abs((d * nrm).sum(axis=-1))
    blue = np.array([0.3, 0.6, 0.9])
    ball = blue * 0.5 + blue * cos[..., None] * 0.5
    assert (t[on_ball] < 10).all()
    # mirror at z = 12: reflected direction (a, b, -1) / n, step 2e-5 along it, down to z = -4; the distance keeps adding up
    all_t = 12 * n + (16 - 2e-5 / n) * n
    wall = np.array([0.7, 0.7, 0.2])
    c = wall * 0.5 + wall * (1 / n)[..., None] * 0.5
    seen = np.array([0.9, 0.95, 1.0]) * c * ((1 - (all_t - 10) / 200) ** 4)[..., None]
    out = np.where(on_ball[..., None], ball, seen)
    return np.sqrt(out), safe, on_ball


def scene_ir(tmp_path=None, s=1.0, grid="grid"):
    from oracle import frontend
    if s == 1.0 and grid == "grid":
        return frontend.scene_ir(frontend.load_scene(SCENE), "analytic")
    text = open(SCENE, encoding="utf-8").read()
    old = '(name: "gate_b", data: Simple(offset: (100.0, 0.0, 3.0), scale: 1.0,'
    mat = 'normal_coef: 0.25, grid: true, grid_scale: 1.0, grid_coef: 0.3, grid2: false, grid3: false'
    assert old in text and mat in text
    text = text.replace(old, old.replace("scale: 1.0", f"scale: {s!r}"))
    text = text.replace(mat, mat.replace("grid2: false", "grid2: " + str(grid == "grid2").lower()).replace("grid3: false", "grid3: " + str(grid == "grid3").lower()))
    path = tmp_path / f"analytic_s{s}_{grid}.ron"
    path.write_text(text, encoding="utf-8")
    return frontend.scene_ir(frontend.load_scene(str(path)), f"analytic_s{s}_{grid}")


def oracle_frame(tmp_path=None, s=1.0):
    from oracle import runner
    ir = scene_ir(tmp_path, s)
    return ir, runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0)


def test_oracle_frame_equals_the_closed_form():
    want, safe, in_gate, on_near = closed_form(W, H)
    _, got = oracle_frame()
    assert np.all(got[..., 3] == 1.0)
    assert in_gate.sum() > 200 and on_near.sum() > 200 and (~in_gate & ~on_near).sum() > 1000 and safe.mean() > 0.9
    err = np.abs(got[..., :3].astype(np.float64) - want).max(axis=-1)
    assert err[safe].max() < 2e-5, (err[safe].max(), np.argwhere(safe & (err >= 2e-5))[:5])
    # the three regions really are different colours, and the few boundary pixels are one of the neighbouring closed forms
    assert len({tuple(np.round(want[y, x], 3)) for y, x in ((H // 2 - 1, W // 2 - 1), (H // 2 + 12, W // 2 + 14), (2, 2))}) == 3
    # centre pixel by hand: uv = 0 -> straight down the axis, through the gate, wall at distance 30 - 2e-5, grid cell
    # factor 0.7: sqrt(green * (0.7 + 0.3 * 0.7) * (1 - (20 - 2e-5) / 200)^4)
    centre = got[H // 2 - 1, W // 2 - 1, :3]
    hand = np.sqrt(np.array([0.2, 0.9, 0.5]) * (0.7 + 0.3 * 0.7) * (1 - (20 - 2e-5) / 200) ** 4)
    assert np.abs(centre - hand).max() < 2e-6


def test_oracle_depth_and_switches_on_the_analytic_scene():
    """Invariants with known answers on the same scene: depth 1 cannot finish the gate path (black behind the gate), depth 2
    can; without darkening the far wall is the undarkened closed form; without the angle term cos drops out."""
    from oracle import frontend, runner
    want, safe, in_gate, on_near = closed_form(W, H)
    ir = frontend.scene_ir(frontend.load_scene(SCENE), "analytic")
    orc = runner.Oracle(ir, "strict")
    d1 = orc.render(W, H, 1, camera=IDENTITY, camera_scale=1.0)
    assert np.all(d1[in_gate & safe][:, :3] == 0.0)                       # loop exhausted -> black (frag.glsl:158)
    assert np.abs(d1[on_near & safe][:, :3] - want[on_near & safe]).max() < 2e-5
    d2 = orc.render(W, H, 2, camera=IDENTITY, camera_scale=1.0)
    assert np.abs(d2[safe][:, :3] - want[safe]).max() < 2e-5
    nodark = orc.render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0, darken_by_distance=0)
    px, py = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    a, b = (px + 1 - W / 2) * 2 / H, (py + 1 - H / 2) * 2 / H
    n = np.sqrt(a * a + b * b + 1)
    all_t = 3 * n + (30 - (3 + 2e-5 / n)) * n
    undo = ((1 - (all_t - 10) / 200) ** 4)[..., None]
    sel = in_gate & safe
    assert np.abs(nodark[sel][:, :3].astype(np.float64) ** 2 - (want[sel] ** 2) / undo[sel]).max() < 5e-5
    flat = orc.render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0, angle_color_disable=1)
    sel = on_near & safe
    assert np.abs(flat[sel][:, :3] - np.sqrt(np.array([0.8, 0.4, 0.2]))).max() < 2e-6


@pytest.mark.parametrize("s", [2.0, 0.5])
def test_scaling_gate_carries_tmul_and_the_offset_step(s, tmp_path):
    """The gate's far side scaled by s: the jump magnifies about the gate centre, the offset step is s * offset, distance
    beyond the gate counts 1 / s (tmul) in the darkening.  Oracle and host-run kernel program against the closed form."""
    from test_program_on_host import _run_on_host
    want, safe, in_gate, _ = closed_form(W, H, s)
    base, _, _, _ = closed_form(W, H, 1.0)
    assert np.abs(want - base)[in_gate].max() > 0.01            # the scale really changes what is seen through the gate
    ir, got = oracle_frame(tmp_path, s)
    assert np.abs(got[..., :3].astype(np.float64) - want)[safe].max() < 2e-5
    prog, _ = _run_on_host(tmp_path, f"analytic_s{s}", None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": IDENTITY})
    assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32))


def test_antialiasing_averages_the_r2_samples():
    """aa_count = 4: the mean of the LINEAR colours at the four R2 sample positions, then the sqrt (frag.glsl:515-526, 550);
    aa_start = 2, aa_count = 1: the single sample R2(2) (what the motion-blur loop uses, main.rs:1797)."""
    from oracle import runner
    orc = runner.Oracle(scene_ir(), "strict")
    parts = [closed_form(W, H, sample=i, linear=True) for i in range(4)]
    want = np.sqrt(sum(p[0] for p in parts) / 4)
    safe = np.logical_and.reduce([p[1] for p in parts])
    same_region = np.logical_and.reduce([p[2] == parts[0][2] for p in parts]) & np.logical_and.reduce([p[3] == parts[0][3] for p in parts])
    assert (safe & ~same_region).sum() > 50                      # pixels whose samples straddle two regions are checked too
    got = orc.render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0, aa_count=4)
    assert np.abs(got[..., :3].astype(np.float64) - want)[safe].max() < 2e-5
    w2, safe2, _, _ = closed_form(W, H, sample=2)
    got2 = orc.render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0, aa_count=1, aa_start=2)
    assert np.abs(got2[..., :3].astype(np.float64) - w2)[safe2].max() < 2e-5
    assert np.abs(w2 - closed_form(W, H)[0]).max() > 0.05       # and it is a different frame from sample 0


def test_projections(tmp_path):
    """Other ray maps of get_color2 (frag.glsl:408-455) on the same scene: a narrower view angle, the 360 and VR180
    equirectangular cameras with their black bars; oracle and host-run kernel program."""
    from oracle import runner
    from test_program_on_host import _run_on_host
    ir = scene_ir()
    orc = runner.Oracle(ir, "strict")

    def narrow(x, y):
        hh = np.tan(np.pi / 6)                                  # view angle 60 degrees
        nn = np.sqrt((x * hh) ** 2 + (y * hh) ** 2 + 1)
        return x * hh / nn, y * hh / nn, 1 / nn, np.zeros(x.shape, dtype=bool), np.zeros(x.shape, dtype=bool)

    def vr180(x, y):
        yaw, pitch = x * np.pi / 2, y * np.pi / 2
        edge = (np.abs(np.abs(x) - 1) < 1e-6) | (np.abs(np.abs(y) - 1) < 1e-6)
        return np.sin(yaw) * np.cos(pitch), np.sin(pitch), np.cos(yaw) * np.cos(pitch), (np.abs(x) > 1) | (np.abs(y) > 1), edge

    def full360(x, y):
        ax, ay = W / min(W, H), H / min(W, H)
        rx, ry = (2 * ay, ay) if ax >= 2 * ay else (ax, ax / 2)
        yaw, pitch = x / rx * np.pi, y / ry * np.pi / 2
        edge = (np.abs(np.abs(x) - rx) < 1e-6) | (np.abs(np.abs(y) - ry) < 1e-6)
        return np.sin(yaw) * np.cos(pitch), np.sin(pitch), np.cos(yaw) * np.cos(pitch), (np.abs(x) > rx) | (np.abs(y) > ry), edge

    def panini(x, y):
        fov, dd = np.pi / 2, 0.7                                # PaniniProjection(tc, _view_angle, _panini_param), frag.glsl:304-340
        d2 = dd * dd
        fo = np.pi / 2 - fov * 0.5
        f = np.cos(fo) / np.sin(fo)
        f2 = f * f
        bb = (np.sqrt(max(0.0, (dd + d2) ** 2 * (f2 + f2 * f2))) - (dd * f + f)) / (d2 + d2 * f2 - 1.0)
        hx, v = x * bb, y * bb
        k = hx * hx / (dd + 1.0) ** 2
        discr = np.maximum(0.0, k * k * d2 - (k + 1.0) * (k * d2 - 1.0))
        cos_phi = (-k * dd + np.sqrt(discr)) / (k + 1.0)
        tan_theta = v / ((dd + 1.0) / (dd + cos_phi))
        sin_phi = np.sqrt(np.maximum(0.0, 1.0 - cos_phi ** 2)) * np.where(hx < 0, -1.0, 1.0)
        sc = 1 / np.sqrt(1.0 + tan_theta ** 2)
        return sin_phi * sc, tan_theta * sc, cos_phi * sc, np.zeros(x.shape, dtype=bool), np.abs(hx) < 1e-9

    for name, proj, kw, attrs in (("narrow", narrow, {"view_angle": np.pi / 3}, {"view_angle": np.pi / 3}),
                                  ("panini", panini, {"use_panini_projection": 1, "panini_param": 0.7},
                                   {"use_panini_projection": 1, "panini_param": 0.7}),
                                  ("vr180", vr180, {"use_180_camera": 1}, {"use_180_camera": 1}),
                                  ("full360", full360, {"use_360_camera": 1}, {"use_360_camera": 1})):
        want, safe, in_gate, on_near = closed_form(W, H, projection=proj)
        got = orc.render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0, **kw)
        assert in_gate.sum() > 20 and on_near.sum() > 100 and safe.mean() > 0.85, name
        err = np.abs(got[..., :3].astype(np.float64) - want)
        assert err[safe].max() < 3e-5, (name, err[safe].max(), np.argwhere(safe & (err.max(axis=-1) >= 3e-5))[:5])
        prog, _ = _run_on_host(tmp_path, name, None, ir=ir, tex={}, depth=DEPTH, attrs=dict(attrs, camera_matrix=IDENTITY))
        assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32)), name
    assert (closed_form(W, H, projection=full360)[0][:2] == 0).all() and (closed_form(W, H, projection=vr180)[0][:, :10] == 0).all()


@pytest.mark.parametrize("grid", ["grid2", "grid3"])
def test_other_grid_patterns(grid, tmp_path):
    """material_simple2's grid2 (circle_sdf discs) and grid3 (framed cells) variants on the far wall, at a frame size where
    the pattern is resolved."""
    from oracle import runner
    w, h = 192, 108
    ir = scene_ir(tmp_path, grid=grid)
    want, safe, in_gate, _ = closed_form(w, h, grid=grid)
    got = runner.Oracle(ir, "strict").render(w, h, DEPTH, camera=IDENTITY, camera_scale=1.0)
    err = np.abs(got[..., :3].astype(np.float64) - want)
    assert safe.mean() > 0.85 and err[safe].max() < 2e-5, (err[safe].max(), np.argwhere(safe & (err.max(axis=-1) >= 2e-5))[:5])
    assert np.abs(want - closed_form(w, h)[0])[in_gate].max() > 0.02            # not the default pattern
    assert len({round(float(x), 3) for x in (want[in_gate][:, 1] / want[in_gate][:, 1].max())}) > 3


def test_library_triangle_and_cylinder(tmp_path):
    """The predefined library's triangle() (barycentric u, v feed the grid) and cylinder() (near side, far side where the near
    one falls off the end, outward normal) against plain analytic geometry."""
    from oracle import frontend, runner
    from test_program_on_host import _run_on_host
    ir = frontend.scene_ir(frontend.load_scene(os.path.join(ROOT, "tests", "fixtures", "analytic4.ron")), "analytic4")
    px, py = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    a, b = (px + 1 - W / 2) * 2 / H, (py + 1 - H / 2) * 2 / H
    n = np.sqrt(a * a + b * b + 1)
    dx, dy, dz = a / n, b / n, 1 / n
    eps = 2e-3
    # triangle (-2,-1), (2,-1), (0,2) in the plane z = 6: P = v0 + u (4, 0) + v (2, 3)
    v = (6 * b + 1) / 3
    u = (6 * a + 2 - 2 * v) / 4
    on_tri = (u >= 0) & (v >= 0) & (u + v <= 1)
    safe = ~((np.abs(u) < eps) | (np.abs(v) < eps) | (np.abs(u + v - 1) < eps))
    fu, fv = np.mod(u * 8 * 0.25, 1.0), np.mod(v * 8 * 0.25, 1.0)
    for f in (fu, fv):
        safe &= ~(on_tri & ((np.abs(f - 0.5) < eps) | (f < eps) | (f > 1 - eps)))
    sx, sy = (fu <= 0.5).astype(np.float64), (fv <= 0.5).astype(np.float64)
    low, high = 0.7 + 0.4 * sx, 1.1 - 0.4 * sx
    factor = low + (high - low) * sy
    gold = np.array([0.9, 0.8, 0.1])
    c = gold * 0.5 + gold * dz[..., None] * 0.5
    tri = c * 0.7 + c * factor[..., None] * 0.3
    # open cylinder of radius 1 about the line x = 2.5, y = 3, for 6 < z < 12: seen partly from outside (near side) and
    # partly through its open near end (only the far side lies between the ends)
    A, Bh, Cc = np.maximum(dx * dx + dy * dy, 1e-300), 2.5 * dx + 3 * dy, 2.5 * 2.5 + 9 - 1.0
    disc = Bh * Bh - A * Cc
    root = np.sqrt(np.where(disc > 0, disc, 0.0))
    t_near, t_far = (Bh - root) / A, (Bh + root) / A
    between = lambda tt: (tt * dz > 6) & (tt * dz < 12)        # noqa: E731
    near_ok, far_ok = between(t_near), between(t_far)
    t = np.where(near_ok, t_near, t_far)
    on_rod = (disc > 0) & (near_ok | far_ok) & ~on_tri
    safe &= ~(np.abs(disc) < 5e-3)
    for tt in (t_near, t_far):
        safe &= ~((disc > 0) & ((np.abs(tt * dz - 6) < 2e-2) | (np.abs(tt * dz - 12) < 2e-2)))
    cos = np.abs(dx * (t * dx - 2.5) + dy * (t * dy - 3))
    blue = np.array([0.4, 0.7, 0.9])
    gray = np.where(t > 10, (t - 10) / 200, 0.0)
    safe &= ~(on_rod & (np.abs(t - 10) < 1e-3))
    rod = blue * cos[..., None] * ((1 - gray) ** 4)[..., None]
    want = np.sqrt(np.where(on_tri[..., None], tri, np.where(on_rod[..., None], rod, 0.36)))
    got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0)
    assert on_tri.sum() > 100 and on_rod.sum() > 80 and ((disc > 0) & ~near_ok & far_ok).sum() > 3 and safe.mean() > 0.85, (on_tri.sum(), on_rod.sum(), ((disc > 0) & ~near_ok & far_ok).sum(), safe.mean())
    err = np.abs(got[..., :3].astype(np.float64) - want)
    assert err[safe].max() < 3e-5, (err[safe].max(), np.argwhere(safe & (err.max(axis=-1) >= 3e-5))[:5])
    prog, _ = _run_on_host(tmp_path, "analytic4", None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": IDENTITY})
    assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32))


def test_which_side_of_a_plane_is_its_back(tmp_path):
    """`back` as is_inside_k receives it (scene.rs:912-927): is_collinear(hit.n, -get_normal(M)) -- true when the ray travels
    along the plane's +z.  A wall that shows one material from the back and another from the front; seen along +z it shows
    the back one, and with its matrix mirrored in z (get_normal flips) the front one."""
    from oracle import frontend, runner
    from test_program_on_host import _run_on_host
    text = open(SCENE, encoding="utf-8").read()
    old = "if (abs(x) < 4. && y > -2. && y < 3.5) return plain_M;"
    assert old in text
    text = text.replace(old, "if (abs(x) < 4. && y > -2. && y < 3.5) { if (back) return plain_M; return gridded_M; }")
    mirrored = text.replace('(name: "near", data: Simple(offset: (0.0, 0.0, 6.0), scale: 1.0, rotate: (0.0, 0.0, 0.0), mirror: (false, false, false))),',
                            '(name: "near", data: Simple(offset: (0.0, 0.0, 6.0), scale: 1.0, rotate: (0.0, 0.0, 0.0), mirror: (false, false, true))),')
    assert mirrored != text
    for name, scene_text, gridded_side in (("two_sided", text, False), ("two_sided_mirrored", mirrored, True)):
        path = tmp_path / f"{name}.ron"
        path.write_text(scene_text, encoding="utf-8")
        ir = frontend.scene_ir(frontend.load_scene(str(path)), name)
        want, safe, _, on_near = closed_form(W, H, near_is_gridded=gridded_side)
        got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0)
        assert np.abs(got[..., :3].astype(np.float64) - want)[safe].max() < 2e-5, name
        prog, _ = _run_on_host(tmp_path, name, None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": IDENTITY})
        assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32)), name
    assert np.abs(closed_form(W, H, near_is_gridded=True)[0] - closed_form(W, H)[0])[on_near].min() > 0.02


def test_teleport_codes_from_a_plain_object_are_ignored(tmp_path):
    """process_plane_intersection (library.glsl:560-572): TELEPORT / TELEPORT_SUBSPACE returned by a non-portal object is
    "wrong code, do nothing" -- the wall is simply not there."""
    from oracle import frontend, runner
    text = open(SCENE, encoding="utf-8").read()
    old = "if (abs(x) < 4. && y > -2. && y < 3.5) return plain_M;"
    want, safe, _, on_near = closed_form(W, H, no_near=True)
    assert on_near.sum() == 0 and np.abs(want - closed_form(W, H)[0]).max() > 0.1
    for code in ("TELEPORT", "TELEPORT_SUBSPACE"):
        path = tmp_path / f"ignored_{code}.ron"
        path.write_text(text.replace(old, old.replace("plain_M", code)), encoding="utf-8")
        ir = frontend.scene_ir(frontend.load_scene(str(path)), f"ignored_{code}")
        got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0)
        assert np.abs(got[..., :3].astype(np.float64) - want)[safe].max() < 2e-5, code


def test_back_flag_on_both_sides_of_a_portal(tmp_path):
    """Flat/Portal: side A is tested with normal = -get_normal(A), side B with +get_normal(B) (scene.rs:928-948), so along +z
    `back` is true at A and false at B.  A gate that teleports only when `back` is false: from the origin (facing A) it is a
    plain disc; from x = 100 (facing B) it jumps the ray back by B -> A and shows the near wall behind A."""
    from oracle import frontend, runner
    from test_program_on_host import _run_on_host
    text = open(SCENE, encoding="utf-8").read()
    old = "if (x*x + y*y < 1.) return TELEPORT;"
    assert old in text
    path = tmp_path / "analytic_back.ron"
    path.write_text(text.replace(old, "if (x*x + y*y < 1.) { if (back) return plain_M; return TELEPORT; }"), encoding="utf-8")
    ir = frontend.scene_ir(frontend.load_scene(str(path)), "analytic_back")
    orc = runner.Oracle(ir, "strict")
    px, py = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    a, b = (px + 1 - W / 2) * 2 / H, (py + 1 - H / 2) * 2 / H
    n = np.sqrt(a * a + b * b + 1)
    dz = 1 / n
    red = np.array([0.8, 0.4, 0.2])
    plain = red * 0.5 + red * dz[..., None] * 0.5
    disc = 9 * a * a + 9 * b * b < 1
    edge = np.abs(9 * a * a + 9 * b * b - 1) < 2e-3
    # facing A: the disc is the plain material at z = 3 (same shade as the near wall: no distance term below 10)
    base, safe, in_gate, _ = closed_form(W, H)
    from_a = orc.render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0)
    want_a = np.where(disc[..., None], np.sqrt(plain), base)
    assert np.abs(from_a[..., :3].astype(np.float64) - want_a)[safe].max() < 2e-5
    # facing B from (100, 0, 0): through the disc the ray comes out of A and meets the near wall at z = 6, |x| < 4, -2 < y < 3.5;
    # around the disc it sees the far wall (centre x = 100) directly
    cam = list(IDENTITY)
    cam[12] = 100.0
    from_b = orc.render(W, H, DEPTH, camera=cam, camera_scale=1.0)
    behind_a = disc & (np.abs(6 * a) < 4) & (6 * b > -2) & (6 * b < 3.5)
    assert behind_a.sum() == disc.sum()                          # the whole disc looks onto the near wall
    sel = disc & ~edge
    assert sel.sum() > 150 and np.abs(from_b[sel][:, :3].astype(np.float64) - np.sqrt(plain[sel])).max() < 2e-5
    around = ~disc & ~edge & (np.abs(30 * a) < 39) & (np.abs(30 * b) < 39)
    green = from_b[around][:, 1] > from_b[around][:, 0]
    assert around.sum() > 1000 and green.all()                   # the far wall's green, not the near wall's red
    for name, c in (("from_a", IDENTITY), ("from_b", cam)):
        prog, _ = _run_on_host(tmp_path, name, None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": c})
        ref = from_a if name == "from_a" else from_b
        assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(ref).view(np.uint32)), name


def test_camera_scale_moves_the_darkening(tmp_path):
    """A camera matrix with columns of length 2.5 sends the same rays but `_camera_scale` = 2.5 (main.rs:1325-1340): darkening
    starts at 25 instead of 10 and its ramp is 2.5 times as long (frag.glsl:130-141) -- the far wall at 30 is barely darkened."""
    from oracle import runner
    from test_program_on_host import _run_on_host
    ir = scene_ir()
    cam = [2.5 if (k % 5 == 0 and k < 15) else (1.0 if k == 15 else 0.0) for k in range(16)]
    want, safe, in_gate, _ = closed_form(W, H, cs=2.5)
    got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=cam, camera_scale=2.5)
    assert np.abs(got[..., :3].astype(np.float64) - want)[safe].max() < 2e-5
    centre = (1 - (30 - 2e-5 - 25) / 200 / 2.5) ** 4
    assert abs(want[H // 2 - 1, W // 2 - 1, 1] ** 2 / (0.9 * (0.7 + 0.3 * 0.7)) - centre) < 1e-9 and centre > 0.96
    assert np.abs(want - closed_form(W, H)[0])[in_gate].min() > 0.01
    from portal_b200.renderer import camera_scale
    assert camera_scale(cam) == 2.5
    # a small camera (scale 0.1): _t_end * 0.1 = 21 < 30 -> the far wall is clamped to fully dark, the near wall (6..) is on the ramp
    tiny = [0.1 if (k % 5 == 0 and k < 15) else (1.0 if k == 15 else 0.0) for k in range(16)]
    want2, safe2, in_gate2, on_near2 = closed_form(W, H, cs=0.1)
    got2 = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=tiny, camera_scale=0.1)
    assert np.all(want2[in_gate2] == 0.0) and np.abs(got2[..., :3].astype(np.float64) - want2)[safe2].max() < 2e-5
    assert 0.05 < want2[on_near2 & safe2][:, 0].min() < want2[on_near2 & safe2][:, 0].max() < 0.9
    prog, _ = _run_on_host(tmp_path, "cs", None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": cam})
    assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32))


def test_gate_that_turns_the_ray(tmp_path):
    """The gate's far side turned a quarter about y, and the far wall turned the same way 30 further along the new heading:
    B A^-1 now rotates positions and directions (x <- z, z <- -x), the wall's own frame undoes it, and the picture through
    the gate is the straight one with the wall at distance 33 -- rotation handedness and order, in pixels."""
    from oracle import frontend, runner
    from test_program_on_host import _run_on_host
    hp = "1.5707963267948966"
    text = open(SCENE, encoding="utf-8").read()
    for old, new in (('(name: "gate_b", data: Simple(offset: (100.0, 0.0, 3.0), scale: 1.0, rotate: (0.0, 0.0, 0.0),',
                      f'(name: "gate_b", data: Simple(offset: (100.0, 0.0, 3.0), scale: 1.0, rotate: (0.0, {hp}, 0.0),'),
                     ('(name: "far", data: Simple(offset: (100.0, 0.0, 30.0), scale: 1.0, rotate: (0.0, 0.0, 0.0),',
                      f'(name: "far", data: Simple(offset: (130.0, 0.0, 3.0), scale: 1.0, rotate: (0.0, {hp}, 0.0),')):
        assert old in text
        text = text.replace(old, new)
    path = tmp_path / "analytic_turn.ron"
    path.write_text(text, encoding="utf-8")
    ir = frontend.scene_ir(frontend.load_scene(str(path)), "analytic_turn")
    tp = np.asarray(ir["uniforms"]["gate_a_to_gate_b_mat_teleport"]["value"]).reshape(4, 4).T
    np.testing.assert_allclose(tp @ [0.5, 0.25, 3.0, 1.0], [100.0, 0.25, 2.5, 1.0], atol=1e-12)      # (x, y, 0) of the gate -> (0, y, -x)
    np.testing.assert_allclose(tp @ [0.0, 0.0, 1.0, 0.0], [1.0, 0.0, 0.0, 0.0], atol=1e-12)          # heading +z -> +x
    want, safe, in_gate, _ = closed_form(W, H, far_z=33.0, see_far_directly=False)
    got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0)
    err = np.abs(got[..., :3].astype(np.float64) - want)
    assert err[safe].max() < 3e-5, (err[safe].max(), np.argwhere(safe & (err.max(axis=-1) >= 3e-5))[:5])
    assert np.abs(want - closed_form(W, H)[0])[in_gate].max() > 0.01
    prog, _ = _run_on_host(tmp_path, "turn", None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": IDENTITY})
    assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32))


def test_time_drives_a_formula_drives_a_matrix_drives_the_pixels(tmp_path):
    """The whole uniform chain in one frame: `time` -> Formula uniform -> Parametrized matrix -> `near_mat_inv` in the block ->
    where the near wall is seen.  Both front-ends evaluate it; the frame follows the closed form with the wall moved."""
    from oracle import frontend, runner
    from portal_b200.host import HostScene
    from test_program_on_host import _run_on_host
    text = open(SCENE, encoding="utf-8").read()
    old = '(name: "near", data: Simple(offset: (0.0, 0.0, 6.0), scale: 1.0, rotate: (0.0, 0.0, 0.0), mirror: (false, false, false))),'
    new = ('(name: "near", data: Parametrized(offset: (x: Uniform(Some(Named("shift"))), y: Value(0.0), z: Value(6.0)), '
           'rotate: (x: Value(0.0), y: Value(0.0), z: Value(0.0)), mirror: (x: Value(0.0), y: Value(0.0), z: Value(0.0)), scale: Value(1.0))),')
    assert old in text and "    uniforms: ([]),\n" in text
    text = text.replace(old, new).replace("    uniforms: ([]),\n", '    uniforms: ([(name: "shift", data: Formula(("time * 2 - 1")))]),\n')
    path = tmp_path / "analytic_time.ron"
    path.write_text(text, encoding="utf-8")
    frames = []
    for tm in (0.0, 0.75, 2.0):
        shift = tm * 2 - 1
        ir = frontend.scene_ir(frontend.load_scene(str(path)), f"analytic_time_{tm}", time=tm)
        hs = HostScene.from_file(str(path))
        hs.set_time(tm)
        table = hs.uniform_table()
        assert table["shift_u"][1] == shift and table["near_mat"][1][12] == shift              # C++ front-end: same chain
        assert ir["uniforms"]["near_mat"]["value"][12] == shift
        want, safe, _, on_near = closed_form(W, H, near_shift=shift)
        got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0)
        assert on_near.sum() > 300 and np.abs(got[..., :3].astype(np.float64) - want)[safe].max() < 2e-5, tm
        frames.append(on_near)
        if tm == 0.75:
            prog, _ = _run_on_host(tmp_path, "time", None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": IDENTITY})
            assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32))
    assert (frames[0] != frames[2]).sum() > 200                    # the wall really moved across the frame


def test_spherical_portal_with_a_larger_far_side(tmp_path):
    """Complex/Portal object: a unit-sphere gate at (0,0,5) whose far side is a radius-2 sphere at (50,0,5).  The jump
    B A^-1 doubles everything about the far centre: the ray goes on from 2 (P - c) + c', its offset step and its remaining
    distance count as in the scaled flat gate, and from inside the far sphere the snippet reports no second hit."""
    from oracle import frontend, runner
    from test_program_on_host import _run_on_host
    ir = frontend.scene_ir(frontend.load_scene(os.path.join(ROOT, "tests", "fixtures", "analytic6.ron")), "analytic6")
    assert ir["material_ids"]["teleport_0_1_M"] == 12 and ir["material_ids"]["teleport_0_2_M"] == 13    # scene.rs:813-839
    px, py = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    a, b = (px + 1 - W / 2) * 2 / H, (py + 1 - H / 2) * 2 / H
    n = np.sqrt(a * a + b * b + 1)
    dx, dy, dz = a / n, b / n, 1 / n
    near_b = lambda x, e, eps=2e-3: np.abs(x - e) < eps           # noqa: E731
    disc = 25 / (n * n) - 24
    in_orb = disc > 0
    safe = ~(np.abs(disc) < 2e-2)
    t1 = 5 / n - np.sqrt(np.where(in_orb, disc, 0.0))
    pxw, pyw, pzw = dx * t1, dy * t1, dz * t1
    step = 2 * 2e-5
    z0 = 5 + 2 * (pzw - 5) + step * dz
    t2 = (30 - z0) * n
    u, v = 2 * pxw + dx * (step + t2), 2 * pyw + dy * (step + t2)
    all_t = t1 + t2 / 2

    def gridded(uu, vv, tt, where):
        nonlocal safe
        green = np.array([0.2, 0.9, 0.5])
        c = green * 0.75 + green * dz[..., None] * 0.25
        fu, fv = np.mod(uu * 0.25, 1.0), np.mod(vv * 0.25, 1.0)
        for f in (fu, fv):
            safe &= ~(where & (near_b(f, 0.5) | near_b(f, 0.0) | near_b(f, 1.0)))
        sx, sy = (fu <= 0.5).astype(np.float64), (fv <= 0.5).astype(np.float64)
        low, high = 0.7 + 0.4 * sx, 1.1 - 0.4 * sx
        factor = low + (high - low) * sy
        c = c * 0.7 + c * factor[..., None] * 0.3
        assert (tt[where] > 10).all()
        return c * ((1 - (np.minimum(tt, 210.0) - 10) / 200) ** 4)[..., None]
    through = gridded(u, v, all_t, in_orb)
    on_near = ~in_orb & (np.abs(8 * a) < 4) & (8 * b > -2) & (8 * b < 3.5)
    safe &= ~(~in_orb & (near_b(np.abs(8 * a), 4) | near_b(8 * b, -2) | near_b(8 * b, 3.5)))
    red = np.array([0.8, 0.4, 0.2])
    near = red * 0.5 + red * dz[..., None] * 0.5
    assert (8 * n[on_near] < 10).all()
    direct = ~in_orb & ~on_near & (np.abs(30 * a - 50) < 40) & (np.abs(30 * b) < 40)
    safe &= ~(~in_orb & ~on_near & (near_b(np.abs(30 * a - 50), 40, 0.05) | near_b(np.abs(30 * b), 40, 0.05)))
    far_direct = gridded(30 * a - 50, 30 * b, 30 * n, direct)
    want = np.sqrt(np.where(in_orb[..., None], through, np.where(on_near[..., None], near, np.where(direct[..., None], far_direct, 0.36))))
    got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0)
    assert in_orb.sum() > 80 and on_near.sum() > 300 and direct.sum() > 300 and safe.mean() > 0.85, (in_orb.sum(), on_near.sum(), direct.sum(), safe.mean())
    err = np.abs(got[..., :3].astype(np.float64) - want)
    assert err[safe].max() < 3e-5, (err[safe].max(), np.argwhere(safe & (err.max(axis=-1) >= 3e-5))[:5])
    prog, _ = _run_on_host(tmp_path, "analytic6", None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": IDENTITY})
    assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32))


def test_intersection_material_between_objects(tmp_path):
    """An intersection material (its own hit + its own final colour, no angle term) at z = 4, an object in front of part of
    it (z = 3: the object wins there) and a wall behind it (z = 7: the disc wins): `nearer` decides per pixel."""
    from oracle import frontend, runner
    from test_program_on_host import _run_on_host
    ir = frontend.scene_ir(frontend.load_scene(os.path.join(ROOT, "tests", "fixtures", "analytic7.ron")), "analytic7")
    px, py = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    a, b = (px + 1 - W / 2) * 2 / H, (py + 1 - H / 2) * 2 / H
    n = np.sqrt(a * a + b * b + 1)
    dz = 1 / n
    eps = 3e-3
    on_slab = (np.abs(3 * a - 1.5) < 0.5) & (np.abs(3 * b) < 2)
    safe = ~((np.abs(np.abs(3 * a - 1.5) - 0.5) < eps) | (np.abs(np.abs(3 * b) - 2) < eps))
    dist = np.hypot(4 * a - 1, 4 * b - 0.5)
    on_disc = ~on_slab & (dist < 1.5)
    for edge in (0.5, 1.0, 1.5):
        safe &= ~(np.abs(dist - edge) < eps)
    k = np.floor(dist / 0.5) + 1
    disc = np.stack([(0.2 * k) ** 2, np.full_like(k, 0.25), (1 - 0.25 * k) ** 2], axis=-1)        # color() squares
    on_wall = ~on_slab & ~on_disc & (np.abs(7 * a) < 9) & (np.abs(7 * b) < 5)
    safe &= ~(~on_slab & ~on_disc & ((np.abs(np.abs(7 * a) - 9) < eps) | (np.abs(np.abs(7 * b) - 5) < eps)))

    def simple(col, all_t):
        col = np.array(col)
        gray = np.where(all_t > 10, (all_t - 10) / 200, 0.0)                                      # frag.glsl:131-141
        return (col * 0.5 + col * dz[..., None] * 0.5) * ((1 - gray) ** 4)[..., None]
    safe &= ~(on_wall & (np.abs(7 * n - 10) < 1e-3))
    want = np.sqrt(np.where(on_slab[..., None], simple((0.8, 0.4, 0.2), 3 * n), np.where(on_disc[..., None], disc,
                            np.where(on_wall[..., None], simple((0.3, 0.3, 0.8), 7 * n), 0.36))))
    got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0)
    assert on_slab.sum() > 100 and on_disc.sum() > 150 and on_wall.sum() > 1000 and (~on_slab & ~on_disc & ~on_wall).sum() > 200
    assert len(np.unique(k[on_disc & safe])) == 3 and safe.mean() > 0.9
    err = np.abs(got[..., :3].astype(np.float64) - want)
    assert err[safe].max() < 2e-5, (err[safe].max(), np.argwhere(safe & (err.max(axis=-1) >= 2e-5))[:5])
    prog, _ = _run_on_host(tmp_path, "analytic7", None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": IDENTITY})
    assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32))


def test_debug_matrix_axes(tmp_path):
    """DebugMatrix: three capsules (radius 0.03, length 1 in the matrix's frame) along the axes, red / green / blue with the
    angle term; here scaled by 5 at z = 6.  Closed form: a capsule is a finite cylinder plus two spheres."""
    from oracle import frontend, runner
    from test_program_on_host import _run_on_host
    ir = frontend.scene_ir(frontend.load_scene(os.path.join(ROOT, "tests", "fixtures", "analytic5.ron")), "analytic5")
    w, h = 192, 108
    px, py = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    a, b = (px + 1 - w / 2) * 2 / h, (py + 1 - h / 2) * 2 / h
    n = np.sqrt(a * a + b * b + 1)
    d = np.stack([a / n, b / n, 1 / n], axis=-1)
    rad, p0 = 0.03 * 5, np.array([0.0, 0.0, 6.0])
    best_t = np.full((h, w), np.inf)
    colour = np.zeros((h, w, 3))
    safe = np.ones((h, w), dtype=bool)
    for axis, rgb in ((0, (0.9, 0.2, 0.2)), (1, (0.2, 0.9, 0.2)), (2, (0.2, 0.2, 0.9))):
        e = np.zeros(3)
        e[axis] = 1.0
        p1 = p0 + 5 * e
        # infinite cylinder about the axis line: |(t d - p0) - ((t d - p0).e) e| = rad
        de, pe = d @ e, -(p0 @ e)
        A = 1 - de * de
        Bh = -(d @ p0) - de * pe                                # half of the linear coefficient of t, negated below
        Cc = p0 @ p0 - pe * pe - rad * rad
        disc = Bh * Bh - A * Cc
        with np.errstate(invalid="ignore", divide="ignore"):
            tc = (-Bh - np.sqrt(np.where(disc > 0, disc, np.nan))) / A
        along = tc * de + pe                                    # position along the axis, 0 .. 5 on the body
        t_body = np.where((disc > 0) & (along > 0) & (along < 5), tc, np.inf)
        t_axis = t_body
        for c in (p0, p1):                                      # end spheres
            bb = d @ c
            dd = bb * bb - (c @ c - rad * rad)
            ts = np.where(dd > 0, bb - np.sqrt(np.where(dd > 0, dd, 0.0)), np.inf)
            t_axis = np.minimum(t_axis, ts)
            safe &= ~(np.abs(dd) < 2e-4)
        safe &= ~(np.abs(disc) < 2e-4) & ~((disc > 0) & ((np.abs(along) < 2e-2) | (np.abs(along - 5) < 2e-2)))
        hit = t_axis < best_t
        pos = d * np.where(np.isfinite(t_axis), t_axis, 0.0)[..., None]
        hh = np.clip((pos - p0) @ e, 0, 5)
        nrm = (pos - p0 - hh[..., None] * e) / rad
        cos = np.abs((d * nrm).sum(axis=-1))
        col = np.array(rgb) ** 2                                # color(r, g, b) squares its arguments
        shade = col * 0.5 + col * cos[..., None] * 0.5
        colour = np.where(hit[..., None], shade, colour)
        best_t = np.where(hit, t_axis, best_t)
    on_axes = np.isfinite(best_t)
    assert (best_t[on_axes] < 10).all()
    want = np.sqrt(np.where(on_axes[..., None], colour, 0.36))
    got = runner.Oracle(ir, "strict").render(w, h, DEPTH, camera=IDENTITY, camera_scale=1.0)
    err = np.abs(got[..., :3].astype(np.float64) - want)
    assert on_axes.sum() > 200 and (safe & on_axes).sum() > 100 and safe.mean() > 0.95, (on_axes.sum(), (safe & on_axes).sum(), safe.mean())
    assert err[safe].max() < 5e-5, (err[safe].max(), np.argwhere(safe & (err.max(axis=-1) >= 5e-5))[:5])
    reds, greens = want[safe & on_axes][:, 0] > 0.5, want[safe & on_axes][:, 1] > 0.5
    assert reds.sum() > 50 and greens.sum() > 50
    from test_program_on_host import W as HW, H as HH
    small = runner.Oracle(ir, "strict").render(HW, HH, DEPTH, camera=IDENTITY, camera_scale=1.0)
    prog, _ = _run_on_host(tmp_path, "analytic5", None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": IDENTITY})
    assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(small).view(np.uint32))


def test_refraction_through_a_pane(tmp_path):
    """Refract material (my_refract, library.glsl:75-92, 357-364): the pane's hit normal faces the ray, so the text takes its
    `!from_outside` branch: ri = 1 / 1.5, dir' = dir * ri + n * (ri * c - sqrt(1 - ri^2 (1 - c^2))); the wall behind is shaded
    with the refracted direction and the distance of both legs."""
    from oracle import frontend, runner
    from test_program_on_host import _run_on_host
    ir = frontend.scene_ir(frontend.load_scene(os.path.join(ROOT, "tests", "fixtures", "analytic3.ron")), "analytic3")
    px, py = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    a, b = (px + 1 - W / 2) * 2 / H, (py + 1 - H / 2) * 2 / H
    n = np.sqrt(a * a + b * b + 1)
    dz = 1 / n
    through = (np.abs(5 * a) < 3) & (np.abs(5 * b) < 2)
    safe = ~(np.abs(np.abs(5 * a) - 3) < 2e-3) & ~(np.abs(np.abs(5 * b) - 2) < 2e-3)
    ri = 1 / 1.5
    disc = 1 - ri * ri * (1 - dz * dz)                         # c = -dot(normal, dir) = d.z with normal = (0, 0, -1)
    rz = np.sqrt(disc)                                         # z of the refracted (unit) direction; x, y are d.xy * ri
    t2 = (5 - 2e-5 * rz) / rz
    wall = np.array([0.9, 0.5, 0.3])

    def shade(cos, all_t):
        c = wall * 0.5 + wall * cos[..., None] * 0.5
        gray = np.where(all_t > 10, (all_t - 10) / 200, 0.0)
        return c * ((1 - gray) ** 4)[..., None]
    seen = np.array([0.8, 0.9, 1.0]) * shade(rz, 5 * n + t2)
    plain = shade(dz, 10 * n)
    safe &= ~(~through & (np.abs(10 * n - 10) < 1e-4)) & ~(through & (np.abs(5 * n + t2 - 10) < 1e-4))     # all_t > _t_start
    want = np.sqrt(np.where(through[..., None], seen, plain))
    got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0)
    assert through.sum() > 500 and safe.mean() > 0.9
    assert np.abs(got[..., :3].astype(np.float64) - want)[safe].max() < 2e-5
    prog, _ = _run_on_host(tmp_path, "analytic3", None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": IDENTITY})
    assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32))


@pytest.mark.parametrize("name", ["analytic", "analytic2", "analytic3"])
def test_cpp_front_end_reads_the_analytic_scenes_like_the_oracle(name):
    """The product's own front-end (C++) on the same files: identical uniform tables, and a program that compiles for sm_100a."""
    from oracle import frontend
    from portal_b200.host import HostRenderer, HostScene
    path = os.path.join(ROOT, "tests", "fixtures", name + ".ron")
    want = frontend.load_scene(path).uniform_table()
    hs = HostScene.from_file(path)
    got = hs.uniform_table()
    assert list(want) == list(got)
    for k in want:
        assert np.array_equal(np.asarray(want[k][1], dtype=np.float64), np.asarray(got[k][1], dtype=np.float64)), k
    r = HostRenderer(hs, device=-1)
    assert "pe_render_kernel" in r.source()
    r.close()


def test_external_ray_probe(tmp_path):
    """teleport_external_ray (frag.glsl:205-257), the camera-teleportation probe, on segments with known answers: a step through
    the gate lands at the jumped end point (plus the offset step); with the far side scaled by 2 the rest of the step is
    stretched by 2 about the gate; a step that stops short of the gate, or passes beside it, reports no teleport; a step into
    the near wall reports the object and no position."""
    from oracle import frontend, runner
    orc = runner.Oracle(scene_ir(), "strict")

    def probe(o, a, b):
        pos, have, enc, chg = o.probe(a, b, camera=IDENTITY, camera_scale=1.0)
        return np.asarray(pos, dtype=np.float64), have, enc, chg
    pos, have, enc, chg = probe(orc, (0.2, 0.1, 2.0), (0.2, 0.1, 4.0))
    assert have and not enc and not chg and np.abs(pos - (100.2, 0.1, 4.0 + 2e-5)).max() < 1e-5
    pos, have, enc, chg = probe(orc, (0.5, -0.3, 2.5), (-0.1, 0.3, 3.5))               # oblique: crosses z = 3 at (0.2, 0, 3)
    d = np.array([-0.6, 0.6, 1.0]) / np.sqrt(0.36 + 0.36 + 1.0)
    assert have and np.abs(pos - (np.array([99.9, 0.3, 3.5]) + d * 2e-5)).max() < 1e-5
    pos, have, enc, chg = probe(orc, (0.2, 0.1, 1.0), (0.2, 0.1, 2.9))                 # stops short of the gate
    assert not have and not enc and np.all(pos == 0)
    pos, have, enc, chg = probe(orc, (1.5, 0.0, 2.0), (1.5, 0.0, 4.0))                 # beside the gate, nothing on the way
    assert not have and not enc
    pos, have, enc, chg = probe(orc, (1.5, 0.0, 5.0), (1.5, 0.0, 7.0))                 # into the near wall at z = 6
    assert not have and enc and np.all(pos == 0)
    big = runner.Oracle(scene_ir(tmp_path, 2.0), "strict")
    pos, have, enc, chg = probe(big, (0.2, 0.1, 2.0), (0.2, 0.1, 4.0))
    assert have and np.abs(pos - (100.4, 0.2, 3.0 + 4e-5 + 2.0)).max() < 1e-5         # gate centre + 2 * (offset, remaining step)
    text = open(SCENE, encoding="utf-8").read().replace("return TELEPORT;", "return TELEPORT_SUBSPACE;")
    path = tmp_path / "probe_sub.ron"
    path.write_text(text, encoding="utf-8")
    o2 = runner.Oracle(frontend.scene_ir(frontend.load_scene(str(path)), "probe_sub"), "strict")
    pos, have, enc, chg = probe(o2, (0.2, 0.1, 2.0), (0.2, 0.1, 4.0))
    assert have and chg and np.abs(pos - (100.2, 0.1, 4.0 + 2e-5)).max() < 1e-5


def test_subspace_flag(tmp_path):
    """TELEPORT_SUBSPACE flips the ray's in_subspace flag (library.glsl:575-589, frag.glsl:33-50); objects are tested only in
    their own space (scene.rs:906-910); a ray that ends on nothing inside the subspace is black (frag.glsl:150-152)."""
    from oracle import frontend, runner
    from test_program_on_host import _run_on_host
    text = open(SCENE, encoding="utf-8").read().replace("return TELEPORT;", "return TELEPORT_SUBSPACE;")
    assert text.count("in_subspace: Normal,") == 3
    head, tail = text.rsplit("in_subspace: Normal,", 1)

    def vr180(x, y):
        yaw, pitch = x * np.pi / 2, y * np.pi / 2
        edge = (np.abs(np.abs(x) - 1) < 1e-6) | (np.abs(np.abs(y) - 1) < 1e-6)
        return np.sin(yaw) * np.cos(pitch), np.sin(pitch), np.cos(yaw) * np.cos(pitch), (np.abs(x) > 1) | (np.abs(y) > 1), edge
    want, safe, in_gate, on_near = closed_form(W, H, projection=vr180)
    plain = closed_form(W, H, projection=vr180)[0]
    direct = safe & ~in_gate & ~on_near & (np.abs(plain - 0.6).max(axis=-1) > 1e-3) & (plain.max(axis=-1) > 0)
    assert direct.sum() > 20                                    # VR180 is wide enough to see the far wall directly as well
    for name, scene_text, through_gate, seen_directly in (("far_in_subspace", head + "in_subspace: Subspace," + tail, True, False),
                                                           ("far_in_normal", text, False, True)):
        path = tmp_path / f"{name}.ron"
        path.write_text(scene_text, encoding="utf-8")
        ir = frontend.scene_ir(frontend.load_scene(str(path)), name)
        got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0, use_180_camera=1)
        exp = want.copy()
        if not through_gate:
            exp[in_gate] = 0.0                                 # nothing to hit inside the subspace -> black
        if not seen_directly:
            exp[direct] = 0.6                                  # the wall lives in the subspace: ordinary rays miss it
        assert np.abs(got[..., :3].astype(np.float64) - exp)[safe].max() < 3e-5, name
        prog, _ = _run_on_host(tmp_path, name, None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": IDENTITY, "use_180_camera": 1})
        assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32)), name


def test_skybox_lookup(tmp_path):
    """A scene with a skybox: rays that end on nothing take the texture's colour at the equirectangular position of the camera
    ray's direction (scene.rs:1052-1058), sampled by the pinned rule; ordinary camera and the 360 camera (the whole sky)."""
    from oracle import frontend, runner
    from test_program_on_host import _run_on_host
    text = open(SCENE, encoding="utf-8").read()
    assert "    textures: ([]),\n" in text and "    animation_stages: ([]),\n" in text
    text = text.replace("    textures: ([]),\n", '    textures: ([(name: "sky", data: ("sky.png"))]),\n')
    text = text.replace("    animation_stages: ([]),\n", '    animation_stages: ([]),\n    skybox: Some("sky"),\n')
    path = tmp_path / "analytic_sky.ron"
    path.write_text(text, encoding="utf-8")
    ir = frontend.scene_ir(frontend.load_scene(str(path)), "analytic_sky")
    assert ir["skybox"] == "sky"
    ys, xs = np.meshgrid(np.arange(8), np.arange(16), indexing="ij")
    sky = np.stack([40 + 12 * xs, 30 + 25 * ys, 250 - 12 * xs + 3 * ys, np.full_like(xs, 255)], axis=-1).astype(np.uint8)

    def full360(x, y):
        rx, ry = W / H, W / H / 2
        yaw, pitch = x / rx * np.pi, y / ry * np.pi / 2
        edge = (np.abs(np.abs(x) - rx) < 1e-6) | (np.abs(np.abs(y) - ry) < 1e-6)
        return np.sin(yaw) * np.cos(pitch), np.sin(pitch), np.cos(yaw) * np.cos(pitch), (np.abs(x) > rx) | (np.abs(y) > ry), edge
    orc = runner.Oracle(ir, "strict", textures={"sky": sky})
    for name, proj, kw in (("plain", None, {}), ("sky360", full360, {"use_360_camera": 1})):
        want, safe, in_gate, on_near = closed_form(W, H, projection=proj, sky=sky)
        got = orc.render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0, **kw)
        err = np.abs(got[..., :3].astype(np.float64) - want)
        assert safe.mean() > 0.85 and err[safe].max() < 3e-5, (name, err[safe].max(), np.argwhere(safe & (err.max(axis=-1) >= 3e-5))[:5])
        assert np.ptp(want[safe & ~in_gate & ~on_near][:, 0]) > 0.2                # the sky really varies over the frame
        prog, _ = _run_on_host(tmp_path, name, None, ir=ir, tex={"sky": sky}, depth=DEPTH, attrs=dict(kw, camera_matrix=IDENTITY))
        assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32)), name


def test_side_by_side_stereo(tmp_path):
    """_draw_side_by_side (frag.glsl:476-497): the left half is the left eye's image, the right half the right eye's, each with
    its own image coordinates; eyes at -+0.4 on the x axis (teleport_eye_matrices with nothing between the eyes)."""
    from oracle import runner
    from test_program_on_host import _run_on_host
    ir, e = scene_ir(), 0.4
    want, safe, in_gate, on_near = closed_form(W, H, stereo=e)
    eye = lambda x: [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, x, 0, 0, 1.0]       # noqa: E731
    got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0, draw_side_by_side=1,
                                             camera_left_eye=eye(-e), camera_right_eye=eye(e))
    assert safe.mean() > 0.9 and in_gate[:, :W // 2].sum() > 50 and in_gate[:, W // 2:].sum() > 50
    assert np.abs(got[..., :3].astype(np.float64) - want)[safe].max() < 2e-5
    assert np.abs(want[:, :W // 2] - want[:, W // 2:]).max() > 0.05            # the two eyes see the gate in different places
    prog, _ = _run_on_host(tmp_path, "stereo", None, ir=ir, tex={}, depth=DEPTH,
                           attrs={"camera_matrix": IDENTITY, "draw_side_by_side": 1, "eye_distance": e})
    assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32))


def test_anaglyph_stereo(tmp_path):
    """_draw_anaglyph (frag.glsl:343-406, 467-473): every sample is traced through BOTH eyes at full resolution and the two
    linear colours are combined into red (left) / cyan (right) with the deghost compensation; both modes, closed form."""
    from oracle import runner
    from test_program_on_host import _run_on_host
    ir, e, P, Q = scene_ir(), 0.4, 0.29, 0.06
    (left, safe_l, _, _), (right, safe_r, _, _) = closed_form(W, H, linear=True, eye=-e), closed_form(W, H, linear=True, eye=e)
    safe = safe_l & safe_r
    luma = np.array([0.299, 0.587, 0.114])
    lc, rc = np.clip(left, 0, 1), np.clip(right, 0, 1)
    l, r = lc @ luma, rc @ luma
    denom = max(1e-6, 1 - P * Q)
    rout, cout = (l - P * r) / denom, (r - Q * l) / denom
    eye = lambda x: [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, x, 0, 0, 1.0]       # noqa: E731
    for mode in (0, 1):
        if mode == 0:
            lin = np.stack([rout, cout, cout], axis=-1)
        else:
            sum_gb = rc[..., 1] + rc[..., 2]
            k = np.where(sum_gb > 1e-6, 2 * cout / np.maximum(sum_gb, 1e-30), 0.0)
            lin = np.stack([rout, rc[..., 1] * k, rc[..., 2] * k], axis=-1)
        want = np.sqrt(np.clip(lin, 0, 1))
        got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0, draw_anaglyph=1, anaglyph_mode=mode,
                                                 camera_left_eye=eye(-e), camera_right_eye=eye(e))
        # sqrt near 0 amplifies the float32 rounding of a clamped-to-zero channel: compare away from it
        ok = safe & (lin.min(axis=-1) > 1e-3)
        assert ok.mean() > 0.8, ok.mean()
        assert np.abs(got[..., :3].astype(np.float64) - want)[ok].max() < 3e-5, mode
        assert np.abs(want[..., 0] - want[..., 1])[ok].max() > 0.05             # the channels really carry different eyes
        prog, _ = _run_on_host(tmp_path, f"anaglyph{mode}", None, ir=ir, tex={}, depth=DEPTH,
                               attrs={"camera_matrix": IDENTITY, "draw_anaglyph": 1, "anaglyph_mode": bool(mode), "eye_distance": e})
        assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32)), mode


def test_depth_map_colouring(tmp_path):
    """_draw_depth_map: the inferno ramp over the depth of the final hit, black where the ray ends on nothing."""
    from oracle import runner
    from test_program_on_host import _run_on_host
    ir = scene_ir()
    want, safe, in_gate, on_near = closed_form(W, H, depth_map=(2.0, 40.0))
    got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0, draw_depth_map=1, depth_map_min=2.0, depth_map_max=40.0)
    assert np.abs(got[..., :3].astype(np.float64) - want)[safe].max() < 2e-5 and safe.mean() > 0.9
    assert len({tuple(np.round(want[y, x], 2)) for y, x in ((H // 2 - 1, W // 2 - 1), (H // 2 + 12, W // 2 + 14), (2, 2))}) == 3
    prog, _ = _run_on_host(tmp_path, "depth", None, ir=ir, tex={}, depth=DEPTH,
                           attrs={"camera_matrix": IDENTITY, "draw_depth_map": 1, "depth_map_min": 2.0, "depth_map_max": 40.0})
    assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32))


def test_sphere_under_a_scaling_matrix_and_a_mirror(tmp_path):
    """Complex object path (transform by M^-1, t / len, normal through adjugate(M); scene.rs:962-976) and the Reflect material
    (my_reflect, the offset along the reflected ray, mul_to_color, distance summed over both legs; library.glsl:70-72, 347-354)
    against their closed form; oracle and host-run kernel program."""
    from oracle import frontend, runner
    from test_program_on_host import _run_on_host
    ir = frontend.scene_ir(frontend.load_scene(os.path.join(ROOT, "tests", "fixtures", "analytic2.ron")), "analytic2")
    got = runner.Oracle(ir, "strict").render(W, H, DEPTH, camera=IDENTITY, camera_scale=1.0)
    want, safe, on_ball = closed_form2(W, H)
    assert on_ball.sum() > 100 and safe.mean() > 0.95
    assert np.abs(got[..., :3].astype(np.float64) - want)[safe].max() < 2e-5
    prog, _ = _run_on_host(tmp_path, "analytic2", None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": IDENTITY})
    assert np.array_equal(np.ascontiguousarray(prog).view(np.uint32), np.ascontiguousarray(got).view(np.uint32))
    one = runner.Oracle(ir, "strict").render(W, H, 1, camera=IDENTITY, camera_scale=1.0)       # depth 1: the mirror path cannot finish
    assert np.all(one[~on_ball & safe][:, :3] == 0.0) and np.abs(one[on_ball & safe][:, :3] - want[on_ball & safe]).max() < 2e-5


def test_generated_program_on_host_equals_the_closed_form(tmp_path):
    """The sm_100a scene program for this scene, run thread by thread on the host (tests/host_harness): bit-identical to the
    oracle's frame and within rounding of the closed form -- the kernel's code is held to the shader text directly, not only
    through the oracle."""
    from test_program_on_host import H as HH, W as HW, _run_on_host
    assert (HW, HH) == (W, H)
    ir, ref = oracle_frame()
    got, _ = _run_on_host(tmp_path, "analytic", None, ir=ir, tex={}, depth=DEPTH, attrs={"camera_matrix": IDENTITY})
    assert np.array_equal(np.ascontiguousarray(got).view(np.uint32), np.ascontiguousarray(ref).view(np.uint32))
    want, safe, _, _ = closed_form(W, H)
    assert np.abs(got[..., :3].astype(np.float64) - want)[safe].max() < 2e-5
