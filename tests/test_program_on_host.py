"""The generated sm_100a scene program -- the exact text pe_scene_source returns, with the hand-written device
headers in it -- compiled as ordinary C++ and executed thread by thread on the host (tests/host_harness/), fed with
the uniform-block image the renderer would upload (pe_scene_uniform_block).  It must reproduce the oracle's frame
bit for bit, and every generator option that claims "same pixels" must keep its promise.  This is a TEST: the
product has no CPU path (tests/test_capi.py::test_render_without_gpu_fails_loudly)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import DEPTH, ROOT, SCENES, load_ir, load_tex
from portal_b200.renderer import SceneRenderer

HARNESS = os.path.join(ROOT, "tests", "host_harness")
W, H = 96, 54


def _run_on_host(tmp_path, tag, scene, options=None, uniforms=None, specialize_ints=True, ir=None, tex=None, depth=None, attrs=None):
    ir = load_ir(scene) if ir is None else ir
    r = SceneRenderer(ir, device=-1, options=options or {}, specialize_ints=specialize_ints)
    r.render_depth = DEPTH[scene] if depth is None else depth
    for k, v in (attrs or {}).items():
        setattr(r, k, v)
    for k, v in (uniforms or {}).items():
        r.set_uniform(k, v)
    block, src = r.uniform_block(W, H), r.source()
    d = tmp_path / tag
    d.mkdir()
    (d / "prog.cu").write_text(src)
    (d / "block.bin").write_bytes(block)
    cc = subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", f'-DPROGRAM_FILE="{d / "prog.cu"}"', "-I", HARNESS,
                         os.path.join(HARNESS, "run_program.cpp"), "-o", str(d / "run")], capture_output=True, text=True, timeout=900)
    assert cc.returncode == 0, cc.stderr[-3000:]
    args = [str(d / "run"), str(d / "block.bin"), str(W), str(H), str(d / "out.f32")]
    tex = (load_tex(scene) or {}) if tex is None else tex
    for t in ir["textures"]:                                   # declaration order = slot order in the block
        if t["name"] not in tex:                               # trailing samplers nobody bound (a video without frames): left
            assert all(u["name"] not in tex for u in ir["textures"][ir["textures"].index(t):])   # unbound -> texture() is black
            break
        arr = np.ascontiguousarray(tex[t["name"]], dtype=np.uint8)
        path = d / f"{t['name']}.rgba"
        path.write_bytes(arr.tobytes())
        args += [str(path), str(arr.shape[1]), str(arr.shape[0])]
    run = subprocess.run(args, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stderr[-2000:]
    return np.fromfile(d / "out.f32", dtype=np.float32).reshape(H, W, 4), src


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("scene", SCENES)
def test_generated_program_on_host_equals_oracle(scene, tmp_path):
    from oracle.runner import Oracle
    got, src = _run_on_host(tmp_path, "default", scene)
    assert "plane_intersect_lazy(" in src or scene == "mobius_monoportal" or "plane_intersect" in src
    want = Oracle(load_ir(scene), "strict", textures=load_tex(scene)).render(W, H, DEPTH[scene])
    assert np.array_equal(_bits(got), _bits(want)), f"{scene}: {(np.abs(got - want) > 0).any(axis=-1).sum()} pixels differ"


@pytest.mark.parametrize("scene", ["portal_in_portal", "triple_portal"])
def test_generator_options_do_not_change_pixels(scene, tmp_path):
    base, src = _run_on_host(tmp_path, "base", scene)
    assert "plane_intersect_lazy(" in src
    for tag, opts, kw in [("nolazy", {"lazy_planes": 0}, {}), ("nohoist", {"hoist_planes": 0}, {}),
                          ("nomat", {"specialize_matrices": 0}, {}), ("noints", {}, {"specialize_ints": False}),
                          ("rolled", {"unroll_loops": 0}, {}), ("blk128", {"block_threads": 128, "min_blocks": 4}, {}),
                          ("smem", {"uniforms_in_smem": 1}, {}), ("tile16", {"tile_w": 16}, {}),
                          ("nocanon", {"canon_rays": 0}, {}), ("nowaware", {"w_aware": 0}, {}),
                          ("tile32", {"tile_w": 32, "block_threads": 256, "min_blocks": 4}, {})]:
        got, s2 = _run_on_host(tmp_path, tag, scene, options=opts, **kw)
        assert s2 != src or tag == "rolled"          # a scene without loops has nothing to keep rolled
        assert np.array_equal(_bits(got), _bits(base)), (scene, tag)


def test_despecialised_uniform_gives_the_same_pixels(tmp_path):
    """After adaptive de-specialisation (an int that changed between renders more than four times) the program reads the
    value from the constant block: same frame as the oracle with that value."""
    from oracle.runner import Oracle
    scene = "portal_in_portal"
    ir = load_ir(scene)
    r = SceneRenderer(ir, device=-1)
    r.render_depth = DEPTH[scene]
    for v in (1, 2, 3, 4, 5, 6):
        r.set_uniform("show_teleported_u", v)
        block, src = r.uniform_block(W, H), r.source()
    assert "#define show_teleported_u (PE_C.i[" in src
    d = tmp_path / "dyn"
    d.mkdir()
    (d / "prog.cu").write_text(src)
    (d / "block.bin").write_bytes(block)
    cc = subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", f'-DPROGRAM_FILE="{d / "prog.cu"}"', "-I", HARNESS,
                         os.path.join(HARNESS, "run_program.cpp"), "-o", str(d / "run")], capture_output=True, text=True, timeout=900)
    assert cc.returncode == 0, cc.stderr[-3000:]
    assert subprocess.run([str(d / "run"), str(d / "block.bin"), str(W), str(H), str(d / "out.f32")], timeout=900).returncode == 0
    got = np.fromfile(d / "out.f32", dtype=np.float32).reshape(H, W, 4)
    orc = Oracle(ir, "strict", textures=load_tex(scene))
    orc.set_uniforms({"show_teleported_u": 6})
    assert np.array_equal(_bits(got), _bits(orc.render(W, H, DEPTH[scene])))


def test_uniform_change_reaches_the_host_run(tmp_path):
    """Same harness, different uniforms: the block image really is what drives the frame."""
    from oracle.runner import Oracle
    ov = {"teleport_light_u": 0, "show_teleported_u": 3}
    got, _ = _run_on_host(tmp_path, "ov", "portal_in_portal", uniforms=ov)
    orc = Oracle(load_ir("portal_in_portal"), "strict", textures=load_tex("portal_in_portal"))
    orc.set_uniforms(ov)
    assert np.array_equal(_bits(got), _bits(orc.render(W, H, DEPTH["portal_in_portal"])))


@pytest.mark.parametrize("scene", ["borromean_rings", "cone", "cylinder", "hopf_link", "matryoshka", "mobius", "non_linear", "recursive_room",
                                   "recursive_space", "sphere_intersection", "spherical_geometry", "time_portal_spacetime", "trefoil"])
def test_more_reference_scenes_on_host(scene, tmp_path):
    """Reference scenes beyond the five configs (tests/golden/scenes_extra): skybox sampling, Reflect / Refract,
    subspaces, the `Camera` matrix kind, and (cylinder) libm's exp/log -- on the host both sides call the same
    libm, so even that one is bit-exact here."""
    import json
    from oracle.runner import Oracle
    folder = "scenes_host" if scene == "trefoil" else "scenes_extra"     # trefoil: `TrefoilSpecial` packed uniforms
    with open(os.path.join(ROOT, "tests", "golden", folder, f"{scene}.scene.json")) as f:
        ir = json.load(f)
    mono = load_tex("monoportal")["monoportal"]
    tex = {t["name"]: mono for t in ir["textures"]}
    got, _ = _run_on_host(tmp_path, "x", scene, ir=ir, tex=tex, depth=30)
    want = Oracle(ir, "strict", textures=tex).render(W, H, 30)
    assert np.array_equal(_bits(got), _bits(want)), f"{scene}: {(np.abs(got - want) > 0).any(axis=-1).sum()} pixels differ"


def test_renderer_flags_on_host(tmp_path):
    """The renderer-side switches of SceneRenderer::set_uniforms (AA window, colouring flags, depth map, panini /
    360 / VR180 projections, side-by-side stereo): generated program on the host == oracle, bit for bit."""
    from oracle.runner import Oracle
    from portal_b200.renderer import camera_scale
    scene = "monoportal"
    orc = Oracle(load_ir(scene), "strict", textures=load_tex(scene))
    probe = SceneRenderer(load_ir(scene), device=-1)
    left, right = probe.eye_matrices()
    cases = [
        ("aa", dict(aa_count=3, aa_start=2), dict(aa_count=3, aa_start=2)),
        ("flags", dict(grid_disable=True, angle_color_disable=True, darken_by_distance=False, black_border_disable=True),
         dict(grid_disable=1, angle_color_disable=1, darken_by_distance=0, black_border_disable=1)),
        ("depthmap", dict(draw_depth_map=True), dict(draw_depth_map=1)),
        ("panini", dict(use_panini_projection=True, panini_param=0.7), dict(use_panini_projection=1, panini_param=0.7)),
        ("cam360", dict(use_360_camera=True), dict(use_360_camera=1)),
        ("cam180", dict(use_180_camera=True), dict(use_180_camera=1)),
        ("sbs", dict(draw_side_by_side=True),
         dict(draw_side_by_side=1, camera_left_eye=left, camera_right_eye=right, left_eye_scale=camera_scale(left),
              right_eye_scale=camera_scale(right))),
    ]
    plain = orc.render(W, H, DEPTH[scene])
    for tag, attrs, kw in cases:
        got, _ = _run_on_host(tmp_path, tag, scene, attrs=attrs)
        want = orc.render(W, H, DEPTH[scene], **kw)
        assert np.array_equal(_bits(got), _bits(want)), tag
        assert not np.array_equal(_bits(want), _bits(plain)), tag            # the switch really changes the frame


def test_probe_kernel_on_host_equals_oracle_probe(tmp_path):
    """The camera-teleportation probe (pe_probe_kernel, the hand-written restatement of frag.glsl:166-257) run on the host
    against the oracle's probe: the same segments tests/test_parity_gpu.py::test_external_ray_probe uses on the GPU."""
    from oracle.runner import Oracle
    cases = (("portal_in_portal", [([0, 0, -0.5], [0, 0, -1.5]), ([0, 0, 0.5], [0, 0, 0.2]), ([0.1, 0.05, -0.9], [0.12, 0.02, -1.3]),
                                   ([0.2, 0.1, 0.5], [0.1, 0.0, 1.5]), ([0, 0, -0.5], [0, 0, -5.0])]),
             ("monoportal", [([0.0, 0.1, 1.0], [0.0, 0.0, -1.0]), ([1.0, 0.3, 0.2], [-1.0, 0.1, -0.2]), ([3.0, 3.0, 3.0], [3.5, 3.0, 3.0])]))
    crossed = 0
    for scene, segs in cases:
        ir = load_ir(scene)
        r = SceneRenderer(ir, device=-1, options={"with_probe": 1})
        r.render_depth = DEPTH[scene]
        if "teleport_light_u" in ir["uniforms"]:
            r.set_uniform("teleport_light_u", 1)                  # main.rs:1367: the probe forces it on
        block, src = r.uniform_block(2, 3), r.source()
        assert "#define PE_WITH_PROBE 1" in src
        d = tmp_path / scene
        d.mkdir()
        (d / "prog.cu").write_text(src)
        (d / "block.bin").write_bytes(block)
        cc = subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", f'-DPROGRAM_FILE="{d / "prog.cu"}"', "-I", HARNESS,
                             os.path.join(HARNESS, "run_program.cpp"), "-o", str(d / "run")], capture_output=True, text=True, timeout=900)
        assert cc.returncode == 0, cc.stderr[-3000:]
        orc = Oracle(ir, "strict", textures=load_tex(scene))
        for a, b in segs:
            a32, b32 = np.asarray(a, np.float32), np.asarray(b, np.float32)
            run = subprocess.run([str(d / "run"), str(d / "block.bin"), "probe"] + [float(x).hex() for x in list(a32) + list(b32)],
                                 capture_output=True, text=True, timeout=300)
            assert run.returncode == 0, run.stderr
            out = np.asarray([float.fromhex(x) for x in run.stdout.split()], dtype=np.float32)
            rpos, rhr, reo, rcs = orc.probe(a, b)
            assert np.array_equal(out[:3].view(np.uint32), rpos.view(np.uint32)), (scene, a, b, out, rpos)
            assert (out[3] != 0, out[4] != 0, out[5] != 0) == (rhr, reo, rcs), (scene, a, b)
            crossed += int(rhr)
    assert crossed >= 3


TORTURE = """vec3 c = vec3(.5, 1., 2.e-1);
vec2 q = (hit.n.xy + vec2(1.)).yx * .5;
c.xy = q;
c.zy += vec2(1e-1, -.25) * q.yx;
c.x++;
c.x -= 1.0;
float arr[3];
arr[0] = c.x; arr[1] = c.y; arr[2] = c.z;
int k = int(floor(abs(hit.u) * 3.)) % 3;
float pick = arr[k];
mat3 m = mat3(vec3(1., 0., 0.), vec3(0., 2., 0.), vec3(0., 0., 3.));
vec3 w = (matrixCompMult(m, m) * c).zxy;
w = (k == 1) ? w.yzx : -w;
bvec3 lt = lessThan(w, vec3(0.25));
bvec2 b2 = bvec2(w.x > 0., w.y > 0.);
if (b2.x && !b2.y) { w.xz *= 2.; }
if (any(lt) && !all(lt)) { w = mix(w, w.zxy, lt); }
w = faceforward(w, r.d.xyz, hit.n);
vec4 v4 = vec4(w, 1.).wzyx;
v4.xw = v4.wx;
float s = 0.;
for (int i = 0; i < 4; i++) { if (i == 2) continue; s += float(i) * v4[i]; }
do { s *= .5; } while (s > 1.);
s = tanh(s) + 0.1 * trunc(s * 3.) + 0.01 * round(pick * 7.) + 0.001 * roundEven(pick * 5.) + 0.0001 * (cosh(q.x) - sinh(q.y));
s += pow(abs(q.x) + .1, 1.7) * exp(-q.y) + log(1. + abs(q.x)) + exp2(q.y) * log2(2. + q.x);
c = clamp(abs(vec3(s, pick, length(v4.xyz) * .1)), 0., 1.);
c = mix(c, c.zyx, step(.5, fract(hit.v)));
return material_simple(hit, r, c, 5e-1, false, 4e0, 3e-1);"""


def test_synthetic_snippet_surface(tmp_path):
    """A material snippet that leans on the GLSL surface scene authors can reach -- literal forms, r/lvalue swizzles,
    arrays, `%`, ternaries, loops with continue / do-while, mat3 products, boolean vectors and relational functions, the
    less common built-ins and the pinned elementary functions -- through BOTH independent GLSL rewriters (pe_codegen.cpp and
    oracle/gen_oracle.py): the two programs must compile and agree bit for bit."""
    from oracle import frontend
    from oracle.runner import Oracle
    ir = frontend.scene_ir(frontend.load_scene(os.path.join(ROOT, "tests", "fixtures", "two_spheres.ron")), "two_spheres")
    n = 0
    for m in ir["materials"]:
        if m["type"] == "Complex":
            m["code"] = TORTURE
            n += 1
    assert n == 1
    got, _ = _run_on_host(tmp_path, "torture", "two_spheres", ir=ir, tex={}, depth=12)
    want = Oracle(ir, "strict").render(W, H, 12)
    assert np.array_equal(_bits(got), _bits(want))
    plain = frontend.scene_ir(frontend.load_scene(os.path.join(ROOT, "tests", "fixtures", "two_spheres.ron")), "two_spheres")
    assert not np.array_equal(_bits(want), _bits(Oracle(plain, "strict").render(W, H, 12)))       # the snippet is really on the path


@pytest.mark.parametrize("seed", [0, 1, 2, 5])
def test_random_glsl_expressions_agree(seed, tmp_path):
    """Differential fuzz of the two GLSL layers (rewriter + value types + built-ins + elementary functions): a material made of
    24 random, well-typed expressions (tests/helpers/glsl_fuzz.py) through pe_codegen.cpp / pe_glsl.cuh and through
    gen_oracle.py / glsl_compat.h -- both must compile it and produce the same bits.  (Offline: 120 seeds, all bit-exact.)"""
    from helpers.glsl_fuzz import snippet
    from oracle import frontend
    from oracle.runner import Oracle
    ir = frontend.scene_ir(frontend.load_scene(os.path.join(ROOT, "tests", "fixtures", "two_spheres.ron")), "two_spheres")
    for m in ir["materials"]:
        if m["type"] == "Complex":
            m["code"] = snippet(seed)
    got, _ = _run_on_host(tmp_path, "fz", "two_spheres", ir=ir, tex={}, depth=6)
    want = Oracle(ir, "strict").render(W, H, 6)
    assert np.array_equal(_bits(got), _bits(want))
    assert len(np.unique(_bits(got)[..., 0])) > 50                  # the random material really colours the frame
