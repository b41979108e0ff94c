"""Per-scene oracle generator -- ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Turns a scene IR (oracle/frontend.py) into one C++ translation unit that restates, for that
scene, what the reference's shader generator emits
(/root/reference/src/gui/scene.rs:693-1110, `generate_shader_code`):

  * uniform declarations            scene.rs:661-691   -> fields of PE_U + name macros
  * material #defines + if-chain    scene.rs:720-842   -> material_process()
  * is_inside_N / intersect_N       scene.rs:847-883   -> translated user snippets
  * per-object intersection blocks  scene.rs:885-1009  -> scene_intersect()
  * intersection-material calls     scene.rs:1011-1035 -> scene_intersect_material_process()
  * user library                    scene.rs:1037-1044
  * marker-line filtering           scene.rs:1066-1107 with the native defaults
    for_prefer_variable = true, use_300_version = true (main.rs:935-941)

and compiles it with g++ (no GPU involved).  Snippets stay GLSL in the scene file; the only
rewriting is lexical (glsl_to_cpp): float literals get the arithmetic type's suffix, multi
-component swizzles become accessor calls, parameter qualifiers become references.

Nothing under portal_b200/ imports this module.

parity unpinned: the reference holds no tests, golden vectors or fixtures for this path (oracle/README.md);
the pins are this repo's committed goldens and its second, independent implementations.
"""
from __future__ import annotations

import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")

_TOKEN_RE = re.compile(
    r"""
    (?P<ws>[ \t\r\n]+)
  | (?P<lc>//[^\n]*)
  | (?P<bc>/\*.*?\*/)
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[fFuU]?)
  | (?P<id>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op>\+\+|--|<<=|>>=|<<|>>|<=|>=|==|!=|&&|\|\||\^\^|\+=|-=|\*=|/=|%=|&=|\|=|\^=|[-+*/%<>=!&|^~?:;,.(){}\[\]\#\\])
    """,
    re.X | re.S,
)

_SWZ_SETS = ("xyzw", "rgba", "stpq")
_CPP_RESERVED = {
    "this", "new", "delete", "class", "template", "typename", "namespace", "using", "operator", "private", "public",
    "protected", "friend", "virtual", "auto", "register", "signed", "unsigned", "long", "short", "char", "double",
    "static_cast", "union", "enum", "typedef", "extern", "goto", "try", "catch", "throw", "export", "inline", "mutable",
    "explicit", "volatile", "sizeof", "and", "or", "not", "xor", "asm", "near", "far",
}
_DROP_QUALIFIERS = {"highp", "mediump", "lowp", "in"}


def filter_marker_lines(code: str) -> str:
    """scene.rs:1066-1107 with for_prefer_variable=true: drop !FOR_NUMBER! lines (snippets carry them too)."""
    return "\n".join(line for line in code.split("\n") if "!FOR_NUMBER!" not in line)


def is_swizzle(name: str) -> bool:
    if not (2 <= len(name) <= 4):
        return False
    return any(all(c in s for c in name) for s in _SWZ_SETS)


_ASSIGN_AHEAD = re.compile(r"\s*(?:=(?!=)|\+=|-=|\*=|/=)")


def glsl_to_cpp(code: str, float_suffix: str, swizzles: set, lvalue_swizzles: set | None = None) -> str:
    """Lexical GLSL -> C++ rewrite (see module docstring)."""
    code = filter_marker_lines(code)
    out = []
    pos = 0
    prev_sig = None  # previous significant token text
    pending_ref = False
    n = len(code)
    while pos < n:
        m = _TOKEN_RE.match(code, pos)
        if not m:
            raise ValueError(f"cannot tokenize GLSL near {code[pos:pos+30]!r}")
        pos = m.end()
        kind = m.lastgroup
        text = m.group()
        if kind == "ws":
            out.append(text)
            continue
        if kind in ("lc", "bc"):
            out.append(" " if kind == "bc" else "")
            continue
        if kind == "num":
            body = text.rstrip("fF")
            if body.endswith(("u", "U")):
                out.append(body)
            elif "." in body or "e" in body.lower():
                out.append(body + float_suffix)
            else:
                out.append(body)
        elif kind == "id":
            if prev_sig == "." and is_swizzle(text) and _ASSIGN_AHEAD.match(code, pos) and lvalue_swizzles is not None:
                lvalue_swizzles.add(text)          # `v.xy += ...`: reference bundle instead of a copy
                out.append(text + "_ref()")
            elif prev_sig == "." and is_swizzle(text):
                swizzles.add(text)
                out.append(text + "()")
            elif prev_sig == ".":
                out.append(text)
            elif text in ("out", "inout"):
                pending_ref = True
                prev_sig = text
                continue
            elif text in _DROP_QUALIFIERS:
                prev_sig = text
                continue
            elif text in _CPP_RESERVED:
                out.append(text + "_")
            else:
                out.append(text)
                if pending_ref:
                    out.append("&")
                    pending_ref = False
        else:
            out.append(text)
        prev_sig = text
    return "".join(out)


def _swizzle_methods(swizzles: set, lvalue_swizzles: set = frozenset()):
    comp = {}
    for s in _SWZ_SETS:
        for i, c in enumerate(s):
            comp[c] = i
    names = "xyzw"
    per = {2: [], 3: [], 4: [], 0: []}  # receiver size -> in-class methods; 0 -> out-of-class definitions (PE_SWZ_DEFS)
    for sw in sorted(swizzles):
        idx = [comp[c] for c in sw]
        ret = f"vec{len(sw)}"
        body = ", ".join(names[i] for i in idx)
        for size in (2, 3, 4):
            if max(idx) < size:
                per[size].append(f"inline {ret} {sw}() const;")      # defined after vec4 is complete: a vec2 may widen (v.xyxy)
                per[0].append(f"inline {ret} vec{size}::{sw}() const {{ return {ret}({body}); }}")
    for sw in sorted(lvalue_swizzles):
        idx = [comp[c] for c in sw]
        if len(set(idx)) != len(idx):
            raise ValueError(f"swizzle `{sw}` repeats a component and cannot be assigned to")
        ret = f"swz{len(sw)}_ref"
        body = ", ".join(names[i] for i in idx)
        meth = f"{ret} {sw}_ref() {{ return {ret}{{{body}}}; }}"
        for size in (2, 3, 4):
            if max(idx) < size:
                per[size].append(meth)
    return per


def _lit(v: float, suffix: str) -> str:
    """A Rust `{:e}`-formatted f64 re-parsed by the GLSL compiler == shortest-repr decimal -> float."""
    r = repr(float(v))
    if "." not in r and "e" not in r and "inf" not in r and "nan" not in r:
        r += ".0"
    return r + suffix


def _bool(b) -> str:
    return "true" if b else "false"


RENDERER_FIELDS = [
    ("mat4", "_camera"), ("mat4", "_camera_mul_inv"), ("mat4", "_camera_left_eye"), ("mat4", "_camera_right_eye"),
    ("real", "_left_eye_scale"), ("real", "_right_eye_scale"), ("real", "_panini_param"),
    ("int", "_left_eye_in_subspace"), ("int", "_right_eye_in_subspace"), ("int", "_use_panini_projection"),
    ("int", "_use_360_camera"), ("int", "_use_180_camera"), ("int", "_draw_side_by_side"),
    ("real", "_camera_scale"), ("real", "_view_angle"), ("real", "_t_start"), ("real", "_t_end"),
    ("real", "_offset_after_material"), ("real", "_depth_map_min"), ("real", "_depth_map_max"),
    ("real", "_resolution_x"), ("real", "_resolution_y"),
    ("int", "_ray_tracing_depth"), ("int", "_aa_count"), ("int", "_aa_start"), ("int", "_camera_in_subspace"),
    ("int", "_darken_by_distance"), ("int", "_angle_color_disable"), ("int", "_grid_disable"),
    ("int", "_black_border_disable"), ("int", "_draw_depth_map"),
    ("int", "_draw_anaglyph"), ("int", "_anaglyph_mode"), ("real", "_anaglyph_p"), ("real", "_anaglyph_q"),
]


def uniform_layout(ir: dict):
    """name -> (kind, slot); kinds 'mat4' | 'float' | 'int' in IR order."""
    mats, floats, ints = [], [], []
    for name, u in ir["uniforms"].items():
        {"mat4": mats, "float": floats, "int": ints}[u["type"]].append(name)
    return mats, floats, ints


def generate_source(ir: dict, real: str = "float") -> str:
    suffix = "f" if real == "float" else ""
    swz: set = set()
    swz_w: set = set()

    def tr(code):
        return glsl_to_cpp(code, suffix, swz, swz_w)

    mats, floats, ints = uniform_layout(ir)
    textures = [t["name"] for t in ir["textures"]]
    objects = ir["objects"]
    mat_ids = ir["material_ids"]

    body = []
    w = body.append
    # ---- name macros (scene.rs:661-718 uniforms/textures, :728-733 + :813-825 material defines)
    for k, nme in enumerate(mats):
        w(f"#define {nme} (PE_U.m[{k}])")
    for k, nme in enumerate(floats):
        w(f"#define {nme} (PE_U.f[{k}])")
    for k, nme in enumerate(ints):
        w(f"#define {nme} (PE_U.i[{k}])")
    for k, nme in enumerate(textures):
        w(f"#define {nme}_tex (PE_U.tex[{k}])")
    for nme, mid in mat_ids.items():
        w(f"#define {nme} (USER_MATERIAL_OFFSET + {mid - 10})")
    if real != "float":
        w("#define float double")

    # ---- user library (scene.rs:1037-1044)
    for lib in ir["library"]:
        w(f"// ---- library `{lib['name']}`")
        w(tr(lib["code"]))

    # ---- intersection functions (scene.rs:847-883)
    for pos, o in enumerate(objects):
        if o["class"] == "Flat":
            extra = ", bool first" if o["kind"] == "Portal" else ""
            w(f"static int is_inside_{pos}(vec4 pos, float x, float y, bool back{extra}) {{")
            w(tr(o["code"]))
            w("}")
        elif o["class"] == "Complex":
            extra = ", bool first" if o["kind"] == "Portal" else ""
            w(f"static SceneIntersection intersect_{pos}(Ray r{extra}) {{")
            w(tr(o["code"]))
            w("}")

    # ---- intersection material functions (scene.rs:1011-1024)
    for pos, im in enumerate(ir["intersection_materials"]):
        w(f"static SceneIntersectionWithMaterial intersect_material_{pos}(Ray r) {{")
        w(tr(im["code"]))
        w("}")

    # ---- scene_intersect: frag.glsl:19-31 + scene.rs:885-1009
    w("static SceneIntersection scene_intersect(Ray r) {")
    w("    SceneIntersection i = SceneIntersection{0, intersection_none, false};")
    w("    SceneIntersection ihit = SceneIntersection{0, intersection_none, false};")
    w("    SurfaceIntersection hit = intersection_none;")
    w("    vec3 normal = vec3(real(0));")
    w("    int inside = NOT_INSIDE; (void)inside;")
    w("    real len = real(1); (void)len; (void)ihit; (void)hit; (void)normal;")
    w("    Ray transformed_ray = ray_none; (void)transformed_ray;")
    for pos, o in enumerate(objects):
        cls = o["class"]
        if cls == "DebugMatrix":
            mname = o["matrices"][0]
            if mname is None:
                raise ValueError("DebugMatrix without matrix: the reference's generator returns None")
            w(f"    transformed_ray = transform({mname}_mat_inv, r); len = length(transformed_ray.d); "
              f"transformed_ray = normalize_ray(transformed_ray);")
            w("    ihit = debug_intersect(transformed_ray); ihit.hit.t /= len;")
            w(f"    if (nearer(i, ihit)) {{ i = ihit; i.hit.n = normalize(adjugate({mname}_mat_inv) * i.hit.n); }}")
            continue
        if any(m is None for m in o["matrices"]):
            raise ValueError(f"object {o['name']} references a missing matrix")
        guard = {"Normal": "if (r.in_subspace == false) {", "Subspace": "if (r.in_subspace == true) {", "Both": "{"}[
            o["in_subspace"]]
        w("    " + guard)
        if cls == "Flat" and o["kind"] == "Simple":
            mname = o["matrices"][0]
            w(f"        normal = -get_normal({mname}_mat);")
            w(f"        hit = plane_intersect(r, {mname}_mat_inv, get_normal({mname}_mat));")
            w(f"        if (nearer(i, hit)) {{ i = process_plane_intersection(i, hit, is_inside_{pos}(r.o + r.d * hit.t, "
              f"hit.u, hit.v, is_collinear(hit.n, normal))); }}")
        elif cls == "Flat" and o["kind"] == "Portal":
            for mname, first, which in ((o["matrices"][0], True, 1), (o["matrices"][1], False, 2)):
                w(f"        normal = {'-' if first else ''}get_normal({mname}_mat);")
                w(f"        hit = plane_intersect(r, {mname}_mat_inv, normal);")
                w(f"        if (nearer(i, hit)) {{ i = process_portal_intersection(i, hit, is_inside_{pos}(r.o + r.d * hit.t, "
                  f"hit.u, hit.v, is_collinear(hit.n, normal), {_bool(first)}), teleport_{pos}_{which}_M); }}")
        elif cls == "Complex" and o["kind"] == "Simple":
            mname = o["matrices"][0]
            w(f"        transformed_ray = transform({mname}_mat_inv, r); len = length(transformed_ray.d); "
              f"transformed_ray = normalize_ray(transformed_ray);")
            w(f"        ihit = intersect_{pos}(transformed_ray); ihit.hit.t /= len;")
            w(f"        if (nearer(i, ihit)) {{ i = ihit; i.hit.n = normalize(adjugate({mname}_mat) * i.hit.n); }}")
        elif cls == "Complex" and o["kind"] == "Portal":
            for mname, first, which in ((o["matrices"][0], True, 1), (o["matrices"][1], False, 2)):
                mat = f"teleport_{pos}_{which}_M"
                w(f"        transformed_ray = transform({mname}_mat_inv, r); len = length(transformed_ray.d); "
                  f"transformed_ray = normalize_ray(transformed_ray);")
                w(f"        ihit = intersect_{pos}(transformed_ray, {_bool(first)}); ihit.hit.t /= len;")
                w(f"        if (nearer(i, ihit) && ihit.material != NOT_INSIDE) {{ if (ihit.material == TELEPORT) {{ "
                  f"ihit.material = {mat}; }} if (ihit.material == TELEPORT_SUBSPACE) {{ ihit.material = {mat}; "
                  f"ihit.in_subspace = true; }} i = ihit; i.hit.n = normalize(adjugate({mname}_mat) * i.hit.n); }}")
        else:
            raise ValueError((cls, o.get("kind")))
        w("    }")
    w("    return i;")
    w("}")

    # ---- material_process: frag.glsl:33-50 + scene.rs:720-842
    w("static MaterialProcessing material_process(Ray r, SceneIntersection i) {")
    w("    SurfaceIntersection hit = i.hit;")
    w("    if (i.in_subspace) { r.in_subspace = !r.in_subspace; }")
    w("    if (i.material == 0) {")
    for nme, rgb in (("DEBUG_RED", (0.9, 0.2, 0.2)), ("DEBUG_GREEN", (0.2, 0.9, 0.2)), ("DEBUG_BLUE", (0.2, 0.2, 0.9))):
        w(f"    }} else if (i.material == {nme}) {{")
        w(f"        return material_simple2(hit, r, color({_lit(rgb[0], suffix)}, {_lit(rgb[1], suffix)}, "
          f"{_lit(rgb[2], suffix)}), {_lit(0.5, suffix)}, false, {_lit(1.0, suffix)}, {_lit(0.0, suffix)}, false, false);")
    for m in ir["materials"]:
        w(f"    }} else if (i.material == {m['name']}_M) {{")
        t = m["type"]
        if t == "Simple":
            c = m["color"]
            w(f"        return material_simple2(hit, r, vec3({_lit(c[0], suffix)}, {_lit(c[1], suffix)}, {_lit(c[2], suffix)}), "
              f"{_lit(m['normal_coef'], suffix)}, {_bool(m['grid'])}, {_lit(m['grid_scale'], suffix)}, "
              f"{_lit(m['grid_coef'], suffix)}, {_bool(m['grid2'])}, {_bool(m['grid3'])});")
        elif t == "Reflect":
            c = m["add_to_color"]
            w(f"        return material_reflect(hit, r, vec3({_lit(c[0], suffix)}, {_lit(c[1], suffix)}, {_lit(c[2], suffix)}));")
        elif t == "Refract":
            c = m["add_to_color"]
            w(f"        return material_refract(hit, r, vec3({_lit(c[0], suffix)}, {_lit(c[1], suffix)}, {_lit(c[2], suffix)}), "
              f"{_lit(m['refractive_index'], suffix)});")
        else:
            w(tr(m["code"]))
    for pos, o in enumerate(objects):
        if o["class"] != "DebugMatrix" and o["kind"] == "Portal":
            a, b = o["matrices"]
            w(f"    }} else if (i.material == teleport_{pos}_1_M) {{")
            w(f"        return material_teleport(hit, r, {a}_to_{b}_mat_teleport);")
            w(f"    }} else if (i.material == teleport_{pos}_2_M) {{")
            w(f"        return material_teleport(hit, r, {b}_to_{a}_mat_teleport);")
    w("    }")
    w("    return material_final(vec3(real(0)));")
    w("}")

    # ---- scene_intersect_material_process: frag.glsl:52-59 + scene.rs:1026-1035
    w("static SceneIntersectionWithMaterial scene_intersect_material_process(Ray r) {")
    w("    SceneIntersectionWithMaterial result = SceneIntersectionWithMaterial{scene_intersection_none, material_empty()};")
    w("    SceneIntersectionWithMaterial hit = SceneIntersectionWithMaterial{scene_intersection_none, material_empty()};")
    w("    (void)hit; (void)r;")
    for pos, _ in enumerate(ir["intersection_materials"]):
        w(f"    hit = intersect_material_{pos}(r);")
        w("    if (nearer(result.scene.hit, hit.scene.hit)) { result = hit; }")
    w("    return result;")
    w("}")
    if real != "float":
        w("#undef float")

    per = _swizzle_methods(swz, swz_w)
    head = []
    h = head.append
    h(f"// GENERATED by oracle/gen_oracle.py for scene `{ir['scene']}` -- ORACLE / TEST INFRASTRUCTURE ONLY.")
    user_code = "\n".join([l["code"] for l in ir["library"]])
    for fn in ("transpose", "inverse", "determinant"):
        if re.search(r"\b(?:mat[234]|float)\s+" + fn + r"\s*\(", user_code):
            h(f"#define PE_NO_BUILTIN_{fn}  // the scene defines its own {fn}()")
    h(f"#define PE_REAL {real}")
    h("#define PE_L(x) " + ("x##f" if real == "float" else "x"))
    for size in (2, 3, 4):
        h(f"#define PE_SWZ_VEC{size} " + " ".join(per[size]))
    h("#define PE_SWZ_DEFS " + " ".join(per[0]))
    h('#include "glsl_compat.h"')
    h("#include <cstring>")
    h("#include <cstddef>")
    h("namespace pe_oracle {")
    h("struct RendererUniforms {")
    for typ, nme in RENDERER_FIELDS:
        h(f"    {typ} u{nme};")
    h("};")
    h("static RendererUniforms PE_R;")
    for typ, nme in RENDERER_FIELDS:
        h(f"#define {nme} (PE_R.u{nme})")
    h("struct SceneUniforms {")
    h(f"    mat4 m[{max(1, len(mats))}]; real f[{max(1, len(floats))}]; int i[{max(1, len(ints))}]; "
      f"sampler2D tex[{max(1, len(textures))}];")
    h("};")
    h("static SceneUniforms PE_U;")
    h("}")
    h('#include "portal_library.h"')
    h("namespace pe_oracle {")
    if ir.get("skybox"):
        # skybox_processing, scene.rs:1054-1058 (evaluated once, on the primary ray, before the bounce loop)
        h("#define _skybox_tex (PE_U.tex[%d])" % [t["name"] for t in ir["textures"]].index(ir["skybox"]))
        h("static inline vec3 pe_skybox_color(const Ray& r) {")
        h("    vec4 rd2 = _camera_mul_inv * r.d;")
        h("    real u = atan(rd2.z, rd2.x);")
        h("    real v = atan(sqrt(rd2.x * rd2.x + rd2.z * rd2.z), rd2.y);")
        h("    return sqrvec(vec3(texture(_skybox_tex, vec2((u / PI + PE_L(1.)) / PE_L(2.), v / PI))));")
        h("}")
        h("#define PE_NOT_FOUND_COLOR(r) pe_skybox_color(r)")
    else:
        h("// skybox_processing, scene.rs:1059-1061")
        h("#define PE_NOT_FOUND_COLOR(r) color(PE_L(0.6), PE_L(0.6), PE_L(0.6))")
    tail = []
    t_ = tail.append
    t_("}  // namespace pe_oracle")
    t_('#include "portal_frag.h"')
    t_(_ENTRY_POINTS.replace("@NM@", str(len(mats))).replace("@NF@", str(len(floats)))
       .replace("@NI@", str(len(ints))).replace("@NT@", str(len(textures))))
    return "\n".join(head) + "\n" + "\n".join(body) + "\n" + "\n".join(tail) + "\n"


_ENTRY_POINTS = r"""
#ifdef _OPENMP
#include <omp.h>
#endif
// C entry points; layouts mirrored by oracle/runner.py (ctypes).
struct PeOracleFrame {
    float camera[16];
    float camera_mul_inv[16];
    float camera_left_eye[16], camera_right_eye[16];
    float left_eye_scale, right_eye_scale, panini_param;
    int left_eye_in_subspace, right_eye_in_subspace, use_panini_projection, use_360_camera, use_180_camera, draw_side_by_side;
    float camera_scale, view_angle, t_start, t_end, offset_after_material, depth_map_min, depth_map_max;
    int ray_tracing_depth, aa_count, aa_start, camera_in_subspace, darken_by_distance, angle_color_disable,
        grid_disable, black_border_disable, draw_depth_map;
    int width, height;
    int draw_anaglyph, anaglyph_mode;
    float anaglyph_p, anaglyph_q;
};
extern "C" {
int pe_oracle_counts(int* nm, int* nf, int* ni, int* nt) { *nm = @NM@; *nf = @NF@; *ni = @NI@; *nt = @NT@; return 0; }
int pe_oracle_real_bytes() { return (int)sizeof(pe_oracle::real); }
void pe_oracle_set_uniforms(const float* mats, const float* f, const int* iv) {
    using namespace pe_oracle;
    for (int k = 0; k < @NM@; k++)
        for (int c = 0; c < 4; c++)
            PE_U.m[k].c[c] = vec4(real(mats[16 * k + 4 * c + 0]), real(mats[16 * k + 4 * c + 1]),
                                  real(mats[16 * k + 4 * c + 2]), real(mats[16 * k + 4 * c + 3]));
    for (int k = 0; k < @NF@; k++) PE_U.f[k] = real(f[k]);
    for (int k = 0; k < @NI@; k++) PE_U.i[k] = iv[k];
}
void pe_oracle_set_texture(int slot, const unsigned char* rgba, int w, int h) {
    if (slot < 0 || slot >= @NT@) return;
    pe_oracle::PE_U.tex[slot].data = rgba;  // caller keeps the buffer alive
    pe_oracle::PE_U.tex[slot].w = w;
    pe_oracle::PE_U.tex[slot].h = h;
}
static void pe_oracle_load_frame(const PeOracleFrame* fr) {
    using namespace pe_oracle;
    for (int c = 0; c < 4; c++)
        _camera.c[c] = vec4(real(fr->camera[4 * c + 0]), real(fr->camera[4 * c + 1]), real(fr->camera[4 * c + 2]),
                                 real(fr->camera[4 * c + 3]));
    for (int c = 0; c < 4; c++)
        _camera_mul_inv.c[c] = vec4(real(fr->camera_mul_inv[4 * c + 0]), real(fr->camera_mul_inv[4 * c + 1]),
                                    real(fr->camera_mul_inv[4 * c + 2]), real(fr->camera_mul_inv[4 * c + 3]));
    for (int c = 0; c < 4; c++) {
        _camera_left_eye.c[c] = vec4(real(fr->camera_left_eye[4 * c + 0]), real(fr->camera_left_eye[4 * c + 1]),
                                     real(fr->camera_left_eye[4 * c + 2]), real(fr->camera_left_eye[4 * c + 3]));
        _camera_right_eye.c[c] = vec4(real(fr->camera_right_eye[4 * c + 0]), real(fr->camera_right_eye[4 * c + 1]),
                                      real(fr->camera_right_eye[4 * c + 2]), real(fr->camera_right_eye[4 * c + 3]));
    }
    _left_eye_scale = real(fr->left_eye_scale); _right_eye_scale = real(fr->right_eye_scale);
    _panini_param = real(fr->panini_param);
    _left_eye_in_subspace = fr->left_eye_in_subspace; _right_eye_in_subspace = fr->right_eye_in_subspace;
    _use_panini_projection = fr->use_panini_projection; _use_360_camera = fr->use_360_camera;
    _use_180_camera = fr->use_180_camera; _draw_side_by_side = fr->draw_side_by_side;
    _camera_scale = real(fr->camera_scale); _view_angle = real(fr->view_angle);
    _t_start = real(fr->t_start); _t_end = real(fr->t_end);
    _offset_after_material = real(fr->offset_after_material);
    _depth_map_min = real(fr->depth_map_min); _depth_map_max = real(fr->depth_map_max);
    _resolution_x = real(fr->width); _resolution_y = real(fr->height);
    _ray_tracing_depth = fr->ray_tracing_depth; _aa_count = fr->aa_count; _aa_start = fr->aa_start;
    _camera_in_subspace = fr->camera_in_subspace; _darken_by_distance = fr->darken_by_distance;
    _angle_color_disable = fr->angle_color_disable; _grid_disable = fr->grid_disable;
    _black_border_disable = fr->black_border_disable; _draw_depth_map = fr->draw_depth_map;
    _draw_anaglyph = fr->draw_anaglyph; _anaglyph_mode = fr->anaglyph_mode;
    _anaglyph_p = real(fr->anaglyph_p); _anaglyph_q = real(fr->anaglyph_q);
}
// main()'s probe branch, frag.glsl:527-529: Ray(a, b - a, 1, _camera_in_subspace == 1); out = pos xyz + 3 flags
void pe_oracle_probe(const PeOracleFrame* fr, const float* a, const float* b, float* out_pos, int* out_flags) {
    using namespace pe_oracle;
    pe_oracle_load_frame(fr);
    Ray r = Ray{vec4(real(a[0]), real(a[1]), real(a[2]), real(1)),
                vec4(real(b[0]) - real(a[0]), real(b[1]) - real(a[1]), real(b[2]) - real(a[2]), real(0)), real(1),
                _camera_in_subspace == 1};
    ExternalRayTeleportation t = teleport_external_ray(r);
    out_pos[0] = float(t.pos.x); out_pos[1] = float(t.pos.y); out_pos[2] = float(t.pos.z);
    out_flags[0] = t.have_result; out_flags[1] = t.encounter_object; out_flags[2] = t.change_subspace;
}
// Threads an OpenMP parallel region of this library really gets (what bench.py reports as `cores`).
int pe_oracle_omp_threads(int threads) {
    int n = 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
#endif
    (void)threads;
    return n;
}
void pe_oracle_render(const PeOracleFrame* fr, int row0, int row1, float* out, int* bounces, int threads) {
    using namespace pe_oracle;
    pe_oracle_load_frame(fr);
    const int W = fr->width;
    (void)threads;
    // Work items are 64-pixel runs (not rows) so that a 16-row sample band still feeds every core.
    const long long total = (long long)(row1 - row0) * (long long)W;
    const long long CH = 64;
    const long long chunks = (total + CH - 1) / CH;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (long long ch = 0; ch < chunks; ch++) {
        const long long e = (ch + 1) * CH < total ? (ch + 1) * CH : total;
        for (long long o = ch * CH; o < e; o++) {
            const int y = row0 + int(o / W), x = int(o % W);
            int b = 0;
            vec4 c = shade_pixel(x, y, &b);
            out[4 * o + 0] = float(c.x); out[4 * o + 1] = float(c.y); out[4 * o + 2] = float(c.z); out[4 * o + 3] = float(c.w);
            if (bounces) bounces[o] = b;
        }
    }
}
}
"""


# -------------------------------------------------------------------------- building
VARIANTS = {
    # strict single-thread parity reference
    "strict": (["-O2", "-ffp-contract=off", "-mavx2", "-mfma"], "float"),
    # same arithmetic (contraction still off -> bit-identical to strict), all cores: timing baseline
    "fast": (["-O3", "-ffp-contract=off", "-mavx2", "-mfma", "-fopenmp"], "float"),
    # float64 arithmetic: flags ill-conditioned pixels only
    "f64": (["-O2", "-ffp-contract=off", "-mavx2", "-mfma", "-fopenmp"], "double"),
}


def build(ir: dict, variant: str = "fast", force: bool = False) -> str:
    """Generate + compile the oracle for one scene; returns the .so path (cached by content hash)."""
    flags, real = VARIANTS[variant]
    src = generate_source(ir, real)
    deps = ""
    for fn in ("glsl_compat.h", "portal_library.h", "portal_frag.h"):
        with open(os.path.join(HERE, fn)) as f:
            deps += f.read()
    key = hashlib.sha256((src + deps + " ".join(flags)).encode()).hexdigest()[:16]
    os.makedirs(BUILD_DIR, exist_ok=True)
    base = os.path.join(BUILD_DIR, f"{ir['scene']}_{variant}_{key}")
    so = base + ".so"
    if os.path.exists(so) and not force:
        return so
    cpp = base + ".cpp"
    with open(cpp, "w") as f:
        f.write(src)
    cmd = ["g++", "-std=c++20", "-shared", "-fPIC", "-fno-math-errno", "-fno-trapping-math", *flags, "-I", HERE, cpp, "-o",
           so + ".tmp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed for {ir['scene']}/{variant}:\n{r.stderr[:6000]}")
    os.replace(so + ".tmp", so)
    return so
