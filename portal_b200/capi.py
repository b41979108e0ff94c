"""ctypes binding of the portal_b200 C ABI (include/portal_b200.h).

This is plumbing only: every call goes straight into libportal_b200.so.  There is no Python or
CPU fallback for rendering -- if the library (or a CUDA device) is missing, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libportal_b200.so")

PE_SUBSPACE = {"Normal": 0, "Subspace": 1, "Both": 2}
PE_UNIFORM_MAT4, PE_UNIFORM_FLOAT, PE_UNIFORM_INT = 0, 1, 2


class PeTarget(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("strip_rows", C.c_int32), ("strip_first", C.c_int32),
                ("strip_step", C.c_int32), ("n_strips", C.c_int32), ("full_frame_layout", C.c_int32)]


PH_PROBE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                          C.POINTER(C.c_int32), C.POINTER(C.c_int32))


class PhFrameParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("depth", C.c_int32), ("aa_count", C.c_int32),
                ("aa_start", C.c_int32), ("use_camera", C.c_int32), ("look_at", C.c_double * 3), ("alpha", C.c_double),
                ("beta", C.c_double), ("r", C.c_double)]


class PortalB200Error(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load libportal_b200.so (built in-tree by `python -m portal_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PortalB200Error(
            f"{LIB_PATH} is missing: build it with `python -m portal_b200.build` (there is no fallback path)")
    L = C.CDLL(LIB_PATH)
    vp, cp, i32 = C.c_void_p, C.c_char_p, C.c_int
    dbl3 = C.POINTER(C.c_double)
    sig = {
        "pe_create": (vp, [i32]),
        "pe_destroy": (None, [vp]),
        "pe_last_error": (cp, [vp]),
        "pe_abi_version": (i32, []),
        "pe_scene_begin": (i32, [vp]),
        "pe_scene_add_library": (i32, [vp, cp, cp]),
        "pe_scene_add_material_simple": (i32, [vp, cp, dbl3, C.c_double, i32, C.c_double, C.c_double, i32, i32]),
        "pe_scene_add_material_reflect": (i32, [vp, cp, dbl3]),
        "pe_scene_add_material_refract": (i32, [vp, cp, dbl3, C.c_double]),
        "pe_scene_add_material_complex": (i32, [vp, cp, cp]),
        "pe_scene_add_object_flat": (i32, [vp, cp, i32, cp, cp, cp]),
        "pe_scene_add_object_complex": (i32, [vp, cp, i32, cp, cp, cp]),
        "pe_scene_add_object_debug_matrix": (i32, [vp, cp, cp]),
        "pe_scene_add_intersection_material": (i32, [vp, cp, cp]),
        "pe_scene_declare_uniform": (i32, [vp, cp, i32]),
        "pe_scene_declare_texture": (i32, [vp, cp]),
        "pe_scene_set_skybox": (i32, [vp, cp]),
        "pe_scene_compile": (i32, [vp]),
        "pe_scene_source": (cp, [vp]),
        "pe_scene_cubin": (i32, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "pe_scene_uniform_block": (i32, [vp, i32, i32, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "pe_set_option": (i32, [vp, cp, i32]),
        "pe_set_uniform_mat4": (i32, [vp, cp, C.POINTER(C.c_float)]),
        "pe_set_uniform_f32": (i32, [vp, cp, C.c_float]),
        "pe_set_uniform_i32": (i32, [vp, cp, C.c_int32]),
        "pe_set_uniforms_mat4": (i32, [vp, i32, C.POINTER(cp), C.POINTER(C.c_float)]),
        "pe_set_texture": (i32, [vp, cp, vp, i32, i32]),
        "pe_target_pixels": (C.c_size_t, [C.POINTER(PeTarget)]),
        "pe_render": (i32, [vp, C.POINTER(PeTarget), vp, vp, vp]),
        "pe_render_host": (i32, [vp, C.POINTER(PeTarget), vp]),
        "pe_render_host_rgba8": (i32, [vp, C.POINTER(PeTarget), vp]),
        "pe_render_rgba8": (i32, [vp, C.POINTER(PeTarget), vp, vp]),
        "pe_submit_host_rgba8": (i32, [vp, C.POINTER(PeTarget), vp, C.POINTER(C.c_uint64)]),
        "pe_wait_host": (i32, [vp, C.c_uint64]),
        "pe_submit_host_strips_rgba8": (i32, [vp, C.POINTER(PeTarget), vp, C.POINTER(C.c_uint64)]),
        "pe_host_register": (i32, [vp, vp, C.c_size_t]),
        "pe_host_unregister": (i32, [vp, vp]),
        "pe_host_malloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "pe_host_free": (i32, [vp, vp]),
        "pe_probe_ray": (i32, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int32),
                               C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
        "pe_autotune": (i32, [vp, C.POINTER(PeTarget), i32, cp, C.c_size_t]),
        "pe_sync": (i32, [vp]),
        "pe_launch_count": (C.c_uint64, [vp]),
        "pe_deinterleave_strips": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp]),
        "pe_device_malloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "pe_device_free": (i32, [vp, vp]),
        "pe_memcpy_d2h": (i32, [vp, vp, vp, C.c_size_t, vp]),
        "pe_scratch_buffer": (i32, [vp, i32, C.c_size_t, C.POINTER(vp)]),
        "pe_signal_u32": (i32, [vp, C.POINTER(vp), i32, C.c_uint32, vp]),
        "pe_stream_wait_geq_u32": (i32, [vp, vp, C.c_uint32, vp]),
        "pe_memset_u32": (i32, [vp, vp, C.c_uint32, C.c_size_t, vp]),
        "pe_ipc_export": (i32, [vp, vp, vp]),
        "pe_ipc_open": (i32, [vp, vp, C.POINTER(vp)]),
        "pe_ipc_close": (i32, [vp, vp]),
        "pe_average_frames_rgba8": (i32, [vp, C.POINTER(vp), i32, vp, C.c_size_t, vp]),
        "pe_quantize_rgba8": (i32, [vp, vp, vp, C.c_size_t, vp]),
        "pe_frames_differ": (i32, [vp, vp, vp, C.c_size_t, vp, C.POINTER(C.c_uint32)]),
        "pe_shard_target": (i32, [i32, i32, i32, i32, i32, i32, C.POINTER(PeTarget)]),
        "pe_sharder_create": (i32, [vp, cp, i32, i32, i32, i32, i32, i32, i32, C.POINTER(vp)]),
        "pe_sharder_destroy": (None, [vp]),
        "pe_sharder_last_error": (cp, [vp]),
        "pe_sharder_target": (i32, [vp, C.POINTER(PeTarget)]),
        "pe_sharder_render": (i32, [vp, vp, C.POINTER(vp)]),
        "pe_sharder_release": (i32, [vp, vp]),
        "pe_sharder_render_overlapped": (i32, [vp, vp, C.POINTER(vp)]),
        "pe_sharder_flush": (i32, [vp, vp, C.POINTER(vp)]),
        "pe_sharder_submit": (i32, [vp, C.POINTER(C.c_uint64)]),
        "pe_sharder_complete": (i32, [vp, C.c_uint64]),
        "pe_sharder_wait_frame": (i32, [vp, C.c_uint64, C.POINTER(vp)]),
        "pe_sharder_release_frame": (i32, [vp, C.c_uint64]),
    }
    f64p = C.POINTER(C.c_double)
    sig.update({
        "ph_scene_load_ron": (vp, [cp, C.c_size_t, C.c_char_p, C.c_size_t]),
        "ph_scene_free": (None, [vp]),
        "ph_scene_last_error": (cp, [vp]),
        "ph_scene_set_time": (i32, [vp, C.c_double, C.c_double]),
        "ph_scene_set_value": (i32, [vp, cp, C.c_double]),
        "ph_scene_set_formula_camera": (i32, [vp, f64p]),
        "ph_scene_init_stage": (i32, [vp, cp]),
        "ph_scene_stage_name": (i32, [vp, i32, C.POINTER(cp)]),
        "ph_scene_evaluate": (i32, [vp]),
        "ph_scene_uniform_get": (i32, [vp, i32, C.POINTER(cp), C.POINTER(i32), f64p]),
        "ph_scene_camera": (i32, [vp, f64p, f64p, f64p, f64p, f64p]),
        "ph_scene_texture": (i32, [vp, i32, C.POINTER(cp), C.POINTER(cp)]),
        "ph_scene_count": (i32, [vp, i32]),
        "ph_player_set_update_hook": (i32, [vp, vp, vp]),
        "ph_png_decode": (i32, [vp, C.c_size_t, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), cp, C.c_size_t]),
        "ph_png_encode_rgba8": (i32, [vp, i32, i32, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "ph_png_free": (None, [vp]),
        "ph_video_frames_dir": (i32, [cp, cp, C.c_size_t]),
        "ph_scene_video": (i32, [vp, i32, C.POINTER(cp), C.POINTER(cp), C.POINTER(cp)]),
        "ph_scene_video_frame": (i32, [vp, i32, C.c_uint64, C.POINTER(C.c_uint64)]),
        "ph_scene_build_program": (i32, [vp, vp]),
        "ph_scene_upload_uniforms": (i32, [vp, vp]),
        "ph_orbit_camera_matrix": (None, [f64p, C.c_double, C.c_double, C.c_double, f64p]),
        "ph_camera_scale": (C.c_double, [f64p]),
        "ph_render_frame": (i32, [vp, vp, C.POINTER(PhFrameParams), vp, i32]),
        "ph_frame_uniforms": (i32, [vp, vp, C.POINTER(PhFrameParams)]),
        "ph_render_target": (i32, [vp, vp, C.POINTER(PhFrameParams), C.POINTER(PeTarget), vp, vp]),
        "ph_render_motion_blur_frame": (i32, [vp, vp, C.POINTER(PhFrameParams), i32, i32, i32, C.c_double, vp]),
        "ph_player_new": (vp, [vp]),
        "ph_player_free": (None, [vp]),
        "ph_player_last_error": (cp, [vp]),
        "ph_player_attach": (i32, [vp, vp]),
        "ph_player_set_probe": (i32, [vp, vp, vp]),
        "ph_player_init_stage": (i32, [vp, cp]),
        "ph_player_init_animation": (i32, [vp, cp]),
        "ph_player_select_camera": (i32, [vp, cp]),
        "ph_player_update": (i32, [vp, C.c_double]),
        "ph_player_camera": (i32, [vp, f64p, f64p, C.POINTER(i32), f64p, f64p, f64p, C.POINTER(C.c_int64)]),
        "ph_player_set_stereo": (i32, [vp, i32, C.c_double, i32]),
        "ph_player_set_run_animations": (i32, [vp, i32]),
        "ph_player_animation_overrides": (i32, [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
        "ph_player_set_anaglyph": (i32, [vp, i32, i32, C.c_double, C.c_double]),
        "ph_player_eyes": (i32, [vp, f64p, f64p, C.POINTER(i32), C.POINTER(i32)]),
        "ph_scene_animation_count": (i32, [vp]),
        "ph_scene_animation": (i32, [vp, i32, C.POINTER(cp), f64p]),
        "ph_scene_camera_count": (i32, [vp]),
        "ph_scene_camera_name": (i32, [vp, i32, C.POINTER(cp)]),
        "ph_player_render_frame": (i32, [vp, vp, C.POINTER(PhFrameParams), vp, i32]),
        "ph_player_render_motion_blur_frame": (i32, [vp, vp, C.POINTER(PhFrameParams), i32, i32, i32, C.c_double, vp]),
    })
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def exported_symbols():
    """Names bound above (tests compare them with include/portal_b200.h)."""
    lib()
    return sorted(n for n in dir(_lib) if n.startswith(("pe_", "ph_")))


def b(s: str) -> bytes:
    return s.encode("utf-8")
