"""ctypes driver for the generated oracle libraries -- ORACLE / TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this (see oracle/README.md).  Nothing here touches a GPU.

parity unpinned: the reference holds no tests, golden vectors or fixtures for this path (oracle/README.md);
the pins are this repo's committed goldens and its second, independent implementations.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import gen_oracle

HERE = os.path.dirname(os.path.abspath(__file__))


class PeOracleFrame(C.Structure):
    _fields_ = [
        ("camera", C.c_float * 16),
        ("camera_mul_inv", C.c_float * 16),
        ("camera_left_eye", C.c_float * 16), ("camera_right_eye", C.c_float * 16),
        ("left_eye_scale", C.c_float), ("right_eye_scale", C.c_float), ("panini_param", C.c_float),
        ("left_eye_in_subspace", C.c_int), ("right_eye_in_subspace", C.c_int), ("use_panini_projection", C.c_int),
        ("use_360_camera", C.c_int), ("use_180_camera", C.c_int), ("draw_side_by_side", C.c_int),
        ("camera_scale", C.c_float), ("view_angle", C.c_float), ("t_start", C.c_float), ("t_end", C.c_float),
        ("offset_after_material", C.c_float), ("depth_map_min", C.c_float), ("depth_map_max", C.c_float),
        ("ray_tracing_depth", C.c_int), ("aa_count", C.c_int), ("aa_start", C.c_int), ("camera_in_subspace", C.c_int),
        ("darken_by_distance", C.c_int), ("angle_color_disable", C.c_int), ("grid_disable", C.c_int),
        ("black_border_disable", C.c_int), ("draw_depth_map", C.c_int),
        ("width", C.c_int), ("height", C.c_int),
        ("draw_anaglyph", C.c_int), ("anaglyph_mode", C.c_int), ("anaglyph_p", C.c_float), ("anaglyph_q", C.c_float),
    ]


def load_ir(path: str) -> dict:
    with open(path) as f:
        return json.load(f)


def load_texture_npz(path: str) -> dict:
    """tests/golden/scenes/<scene>.textures.npz: name -> uint8 [h, w, 4]."""
    if not os.path.exists(path):
        return {}
    with np.load(path) as z:
        return {k: np.ascontiguousarray(z[k]) for k in z.files}


def uniform_arrays(ir: dict, overrides: dict | None = None):
    """IR uniform table -> (mats f32 [nm,16], floats f32 [nf], ints i32 [ni]) in layout order.
    The f64 -> f32 cast is the reference's upload cast (scene.rs:587-588 `as_f32()`, :643 `as f32`)."""
    mats, floats, ints = gen_oracle.uniform_layout(ir)
    u = ir["uniforms"]
    ov = overrides or {}

    def val(n):
        return ov[n] if n in ov else u[n]["value"]

    with np.errstate(over="ignore", invalid="ignore"):
        m = np.array([val(n) for n in mats], dtype=np.float64).reshape(-1, 16).astype(np.float32)
        f = np.array([val(n) for n in floats], dtype=np.float64).astype(np.float32)
    i = np.array([val(n) for n in ints], dtype=np.int32)
    return m, f, i


def make_frame(ir: dict, width: int, height: int, depth: int, camera=None, camera_scale=None, **kw) -> PeOracleFrame:
    fr = PeOracleFrame()
    cam = ir["camera_matrix"] if camera is None else camera
    cam32 = np.asarray(cam, dtype=np.float64).astype(np.float32).reshape(16)
    for k in range(16):
        fr.camera[k] = float(cam32[k])
    cmi = np.asarray(kw.get("camera_mul_inv", np.eye(4).reshape(16)), dtype=np.float64).astype(np.float32).reshape(16)
    for k in range(16):
        fr.camera_mul_inv[k] = float(cmi[k])
    for key, dst in (("camera_left_eye", fr.camera_left_eye), ("camera_right_eye", fr.camera_right_eye)):
        e32 = np.asarray(kw.get(key, cam), dtype=np.float64).astype(np.float32).reshape(16)
        for k in range(16):
            dst[k] = float(e32[k])
    fr.left_eye_scale = float(np.float32(kw.get("left_eye_scale", 1.0)))
    fr.right_eye_scale = float(np.float32(kw.get("right_eye_scale", 1.0)))
    fr.panini_param = float(np.float32(kw.get("panini_param", 1.0)))      # RotateAroundCam::new, main.rs:112
    fr.left_eye_in_subspace = int(kw.get("left_eye_in_subspace", 0))
    fr.right_eye_in_subspace = int(kw.get("right_eye_in_subspace", 0))
    fr.use_panini_projection = int(kw.get("use_panini_projection", 0))
    fr.use_360_camera = int(kw.get("use_360_camera", 0))
    fr.use_180_camera = int(kw.get("use_180_camera", 0))
    fr.draw_side_by_side = int(kw.get("draw_side_by_side", 0))
    d = ir["renderer"]
    fr.camera_scale = float(np.float32(ir["camera_scale"] if camera_scale is None else camera_scale))
    fr.view_angle = float(np.float32(kw.get("view_angle", d["view_angle"])))
    fr.t_start = float(np.float32(kw.get("t_start", d["t_start"])))
    fr.t_end = float(np.float32(kw.get("t_end", d["t_end"])))
    fr.offset_after_material = float(np.float32(kw.get("offset_after_material", ir["cam"]["offset_after_material"])))
    fr.depth_map_min = float(kw.get("depth_map_min", 0.0))
    fr.depth_map_max = float(kw.get("depth_map_max", 10.0))
    fr.ray_tracing_depth = int(depth)
    fr.aa_count = int(kw.get("aa_count", d["aa_count"]))
    fr.aa_start = int(kw.get("aa_start", d["aa_start"]))
    fr.camera_in_subspace = int(kw.get("camera_in_subspace", d["camera_in_subspace"]))
    fr.darken_by_distance = int(kw.get("darken_by_distance", d["darken_by_distance"]))
    fr.angle_color_disable = int(kw.get("angle_color_disable", d["angle_color_disable"]))
    fr.grid_disable = int(kw.get("grid_disable", d["grid_disable"]))
    fr.black_border_disable = int(kw.get("black_border_disable", d["black_border_disable"]))
    fr.draw_depth_map = int(kw.get("draw_depth_map", 0))
    fr.width = int(width)
    fr.height = int(height)
    fr.draw_anaglyph = int(kw.get("draw_anaglyph", 0))                     # SceneRenderer::new, main.rs:1030-1033
    fr.anaglyph_mode = int(kw.get("anaglyph_mode", 0))
    fr.anaglyph_p = float(np.float32(kw.get("anaglyph_p", 0.29)))
    fr.anaglyph_q = float(np.float32(kw.get("anaglyph_q", 0.06)))
    return fr


class Oracle:
    """One compiled oracle (scene x variant)."""

    def __init__(self, ir: dict, variant: str = "fast", textures: dict | None = None):
        self.ir = ir
        self.variant = variant
        self.so_path = gen_oracle.build(ir, variant)
        self.lib = C.CDLL(self.so_path)
        self.lib.pe_oracle_render.argtypes = [C.POINTER(PeOracleFrame), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        self.lib.pe_oracle_render.restype = None
        self.lib.pe_oracle_set_uniforms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.pe_oracle_set_texture.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int]
        self._keep = []
        self.set_uniforms()
        for slot, t in enumerate(ir["textures"]):
            if textures and t["name"] in textures:
                arr = np.ascontiguousarray(textures[t["name"]], dtype=np.uint8)
                self._keep.append(arr)
                self.lib.pe_oracle_set_texture(slot, arr.ctypes.data, arr.shape[1], arr.shape[0])
            else:       # the library is loaded once per process: without this a sampler bound by an earlier Oracle stays bound
                self.lib.pe_oracle_set_texture(slot, None, 0, 0)

    def set_uniforms(self, overrides: dict | None = None):
        m, f, i = uniform_arrays(self.ir, overrides)
        m = np.ascontiguousarray(m if m.size else np.zeros((1, 16), np.float32))
        f = np.ascontiguousarray(f if f.size else np.zeros(1, np.float32))
        i = np.ascontiguousarray(i if i.size else np.zeros(1, np.int32))
        self._u = (m, f, i)
        self.lib.pe_oracle_set_uniforms(m.ctypes.data, f.ctypes.data, i.ctypes.data)

    def render(self, width, height, depth, rows=None, threads=0, want_bounces=False, **kw):
        fr = make_frame(self.ir, width, height, depth, **kw)
        r0, r1 = (0, height) if rows is None else rows
        out = np.empty((r1 - r0, width, 4), dtype=np.float32)
        bnc = np.empty((r1 - r0, width), dtype=np.int32) if want_bounces else None
        self.lib.pe_oracle_render(C.byref(fr), r0, r1, out.ctypes.data, None if bnc is None else bnc.ctypes.data, threads)
        return (out, bnc) if want_bounces else out

    def omp_threads(self, threads: int = 0) -> int:
        """Number of threads an OpenMP region of this build runs with (1 for the strict build)."""
        self.lib.pe_oracle_omp_threads.argtypes = [C.c_int]
        self.lib.pe_oracle_omp_threads.restype = C.c_int
        return int(self.lib.pe_oracle_omp_threads(int(threads)))

    def probe(self, a, b, **kw):
        """teleport_external_ray(a, b) -> (pos[3] float32, have_result, encounter_object, change_subspace).
        Like the reference (main.rs:1367) the probe runs with `teleport_light_u` forced to 1 if the scene has it."""
        had = "teleport_light_u" in self.ir["uniforms"]
        if had:
            self.set_uniforms({"teleport_light_u": 1})
        fr = make_frame(self.ir, 2, 3, 0, **kw)
        pa = (C.c_float * 3)(*[float(np.float32(x)) for x in a])
        pb = (C.c_float * 3)(*[float(np.float32(x)) for x in b])
        out, flags = (C.c_float * 3)(), (C.c_int * 3)()
        self.lib.pe_oracle_probe.argtypes = [C.POINTER(PeOracleFrame), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.pe_oracle_probe(C.byref(fr), pa, pb, out, flags)
        if had:
            self.set_uniforms()
        return np.array(list(out), dtype=np.float32), bool(flags[0]), bool(flags[1]), bool(flags[2])
