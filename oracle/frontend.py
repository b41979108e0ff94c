"""Scene front-end restatement -- ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/README.md).

CPU (float64) restatement of the host side that feeds the reference's hot path:

  * scene load / name resolution   /root/reference/src/gui/scene_serialized.rs:1102-1477
  * uniform evaluation             /root/reference/src/gui/uniform.rs:1009-1140
  * matrix DAG evaluation          /root/reference/src/gui/matrix.rs:510-631
  * uniform table (names, order)   /root/reference/src/gui/scene.rs:424-658
  * material / teleport ids        /root/reference/src/gui/scene.rs:720-842
  * orbit camera                   /root/reference/src/main.rs:278-304, 320-332, 1266-1359

glam 0.13.1 (Cargo.lock:718-719) is not vendored; its DMat4/DQuat arithmetic is
restated here operation by operation (inverse = cofactor expansion scaled by
1/det, no singularity check -> Inf/NaN for singular input, which the headline
scene relies on).

Output: a "scene IR" dict (JSON-serialisable) consumed by the oracle renderer
build (oracle/build_oracle.py), by tests, and -- as committed golden fixtures
under tests/golden/scenes/ -- by the product's Python binding on the GPU box.
Nothing under portal_b200/ imports this module.

parity unpinned: the reference holds no tests, golden vectors or fixtures for this path (oracle/README.md);
the pins are this repo's committed goldens and its second, independent implementations.
"""
from __future__ import annotations

import math
from typing import Any

from . import ron
from .formula import Evaluator, FormulaError
from .mat_sqrt import mat_sqrt
from .ron import Tagged

PI = math.pi


# =============================================================== glam restatement
# Matrices are column-major lists of 4 columns, each a list of 4 floats (DMat4).
def mat_identity():
    return [[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]]


def mat_mul_vec(m, v):
    # glam: x_axis*v.x + y_axis*v.y + z_axis*v.z + w_axis*v.w (left to right)
    out = []
    for r in range(4):
        acc = m[0][r] * v[0]
        acc = acc + m[1][r] * v[1]
        acc = acc + m[2][r] * v[2]
        acc = acc + m[3][r] * v[3]
        out.append(acc)
    return out


def mat_mul(a, b):
    """a * b (glam: result column j = a * b.column(j))."""
    return [mat_mul_vec(a, b[j]) for j in range(4)]


def _mul_ieee(a, b):
    try:
        return a * b
    except OverflowError:  # pragma: no cover
        return math.inf


def mat_inverse(m):
    """glam 0.13.1 DMat4::inverse (GLM-derived cofactor formulation)."""
    m00, m01, m02, m03 = m[0]
    m10, m11, m12, m13 = m[1]
    m20, m21, m22, m23 = m[2]
    m30, m31, m32, m33 = m[3]

    coef00 = m22 * m33 - m32 * m23
    coef02 = m12 * m33 - m32 * m13
    coef03 = m12 * m23 - m22 * m13

    coef04 = m21 * m33 - m31 * m23
    coef06 = m11 * m33 - m31 * m13
    coef07 = m11 * m23 - m21 * m13

    coef08 = m21 * m32 - m31 * m22
    coef10 = m11 * m32 - m31 * m12
    coef11 = m11 * m22 - m21 * m12

    coef12 = m20 * m33 - m30 * m23
    coef14 = m10 * m33 - m30 * m13
    coef15 = m10 * m23 - m20 * m13

    coef16 = m20 * m32 - m30 * m22
    coef18 = m10 * m32 - m30 * m12
    coef19 = m10 * m22 - m20 * m12

    coef20 = m20 * m31 - m30 * m21
    coef22 = m10 * m31 - m30 * m11
    coef23 = m10 * m21 - m20 * m11

    fac0 = (coef00, coef00, coef02, coef03)
    fac1 = (coef04, coef04, coef06, coef07)
    fac2 = (coef08, coef08, coef10, coef11)
    fac3 = (coef12, coef12, coef14, coef15)
    fac4 = (coef16, coef16, coef18, coef19)
    fac5 = (coef20, coef20, coef22, coef23)

    vec0 = (m10, m00, m00, m00)
    vec1 = (m11, m01, m01, m01)
    vec2 = (m12, m02, m02, m02)
    vec3 = (m13, m03, m03, m03)

    inv0 = [vec1[i] * fac0[i] - vec2[i] * fac1[i] + vec3[i] * fac2[i] for i in range(4)]
    inv1 = [vec0[i] * fac0[i] - vec2[i] * fac3[i] + vec3[i] * fac4[i] for i in range(4)]
    inv2 = [vec0[i] * fac1[i] - vec1[i] * fac3[i] + vec3[i] * fac5[i] for i in range(4)]
    inv3 = [vec0[i] * fac2[i] - vec1[i] * fac4[i] + vec2[i] * fac5[i] for i in range(4)]

    sign_a = (1.0, -1.0, 1.0, -1.0)
    sign_b = (-1.0, 1.0, -1.0, 1.0)
    c0 = [inv0[i] * sign_a[i] for i in range(4)]
    c1 = [inv1[i] * sign_b[i] for i in range(4)]
    c2 = [inv2[i] * sign_a[i] for i in range(4)]
    c3 = [inv3[i] * sign_b[i] for i in range(4)]

    col0 = (c0[0], c1[0], c2[0], c3[0])
    d = [m[0][i] * col0[i] for i in range(4)]
    det = d[0] + d[1] + d[2] + d[3]
    if det == 0.0:
        rcp = math.copysign(math.inf, det)  # f64::recip(+-0.0) == +-inf (glam_assert is debug-only)
    else:
        rcp = 1.0 / det

    def scale(col):
        out = []
        for x in col:
            if math.isinf(rcp):
                out.append(math.nan if (x == 0.0 or math.isnan(x)) else math.copysign(math.inf, x) * math.copysign(1.0, rcp))
            else:
                out.append(x * rcp)
        return out

    return [scale(c0), scale(c1), scale(c2), scale(c3)]


def quat_from_axis_angle_x(a):
    return (math.sin(a * 0.5), 0.0, 0.0, math.cos(a * 0.5))


def quat_from_axis_angle_y(a):
    return (0.0, math.sin(a * 0.5), 0.0, math.cos(a * 0.5))


def quat_from_axis_angle_z(a):
    return (0.0, 0.0, math.sin(a * 0.5), math.cos(a * 0.5))


def quat_mul(q0, q1):
    x0, y0, z0, w0 = q0
    x1, y1, z1, w1 = q1
    return (
        w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
        w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1,
        w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1,
        w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1,
    )


def mat_from_scale_rotation_translation(scale, q, t):
    x, y, z, w = q
    x2 = x + x
    y2 = y + y
    z2 = z + z
    xx = x * x2
    xy = x * y2
    xz = x * z2
    yy = y * y2
    yz = y * z2
    zz = z * z2
    wx = w * x2
    wy = w * y2
    wz = w * z2
    xa = [1.0 - (yy + zz), xy + wz, xz - wy, 0.0]
    ya = [xy - wz, 1.0 - (xx + zz), yz + wx, 0.0]
    za = [xz + wy, yz - wx, 1.0 - (xx + yy), 0.0]
    return [
        [c * scale[0] for c in xa],
        [c * scale[1] for c in ya],
        [c * scale[2] for c in za],
        [t[0], t[1], t[2], 1.0],
    ]


def srt_matrix(scale3, rotate3, offset3):
    """matrix.rs:537-547: from_scale_rotation_translation(scale, Rx*Ry*Rz, offset)."""
    q = quat_mul(quat_mul(quat_from_axis_angle_x(rotate3[0]), quat_from_axis_angle_y(rotate3[1])),
                 quat_from_axis_angle_z(rotate3[2]))
    return mat_from_scale_rotation_translation(scale3, q, offset3)


def mat_determinant(m):
    """glam 0.13.1 Matrix4x4::determinant (2x2 sub-determinants of rows 2,3, expansion along column 0)."""
    (m00, m01, m02, m03), (m10, m11, m12, m13), (m20, m21, m22, m23), (m30, m31, m32, m33) = m
    a2323 = m22 * m33 - m23 * m32
    a1323 = m21 * m33 - m23 * m31
    a1223 = m21 * m32 - m22 * m31
    a0323 = m20 * m33 - m23 * m30
    a0223 = m20 * m32 - m22 * m30
    a0123 = m20 * m31 - m21 * m30
    return (m00 * (m11 * a2323 - m12 * a1323 + m13 * a1223)
            - m01 * (m10 * a2323 - m12 * a0323 + m13 * a0223)
            + m02 * (m10 * a1323 - m11 * a0323 + m13 * a0123)
            - m03 * (m10 * a1223 - m11 * a0223 + m12 * a0123))


def _rcp(x):
    if x == 0.0:
        return math.copysign(math.inf, x)
    return 1.0 / x


def _sqrt(x):
    return math.sqrt(x) if x >= 0.0 else math.nan


def _half_over_sqrt(x):
    r = _sqrt(x)
    if r == 0.0:
        return math.copysign(math.inf, r)
    return 0.5 / r


def quat_from_rotation_axes(xa, ya, za):
    """glam 0.13.1 Quaternion::from_rotation_axes (the branch-on-largest-component form)."""
    m00, m01, m02 = xa
    m10, m11, m12 = ya
    m20, m21, m22 = za
    if m22 <= 0.0:
        dif10 = m11 - m00
        omm22 = 1.0 - m22
        if dif10 <= 0.0:
            four_xsq = omm22 - dif10
            inv4x = _half_over_sqrt(four_xsq)
            return (four_xsq * inv4x, (m01 + m10) * inv4x, (m02 + m20) * inv4x, (m12 - m21) * inv4x)
        four_ysq = omm22 + dif10
        inv4y = _half_over_sqrt(four_ysq)
        return ((m01 + m10) * inv4y, four_ysq * inv4y, (m12 + m21) * inv4y, (m20 - m02) * inv4y)
    sum10 = m11 + m00
    opm22 = 1.0 + m22
    if sum10 <= 0.0:
        four_zsq = opm22 - sum10
        inv4z = _half_over_sqrt(four_zsq)
        return ((m02 + m20) * inv4z, (m12 + m21) * inv4z, four_zsq * inv4z, (m01 - m10) * inv4z)
    four_wsq = opm22 + sum10
    inv4w = _half_over_sqrt(four_wsq)
    return ((m12 - m21) * inv4w, (m20 - m02) * inv4w, (m01 - m10) * inv4w, four_wsq * inv4w)


def mat_to_scale_rotation_translation(m):
    """glam 0.13.1 DMat4::to_scale_rotation_translation: scale = column lengths (x signed by det),
    rotation = quaternion of the de-scaled 3x3, translation = w column."""
    det = mat_determinant(m)
    sign = math.nan if det != det else math.copysign(1.0, det)
    length = lambda c: _sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3])  # noqa: E731
    scale = [length(m[0]) * sign, length(m[1]), length(m[2])]
    inv = [_rcp(x) for x in scale]
    axes = [[m[i][r] * inv[i] for r in range(3)] for i in range(3)]
    return scale, quat_from_rotation_axes(*axes), [m[3][0], m[3][1], m[3][2]]


def vec_lerp(a, b, s):
    return [x + ((y - x) * s) for x, y in zip(a, b)]


def quat_lerp(a, b, s):
    """glam 0.13.1 Quaternion::lerp: shortest-arc nlerp."""
    dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]
    bias = 1.0 if dot >= 0.0 else -1.0
    q = [x + (((y * bias) - x) * s) for x, y in zip(a, b)]
    r = _rcp(_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]))
    return tuple(x * r for x in q)


def mat_lerp(first, second, t):
    """matrix.rs:614-628: component-wise lerp of (scale, rotation, translation)."""
    fs, fr, ft = mat_to_scale_rotation_translation(first)
    ss, sr, st = mat_to_scale_rotation_translation(second)
    return mat_from_scale_rotation_translation(vec_lerp(fs, ss, t), quat_lerp(fr, sr, t), vec_lerp(ft, st, t))


def vec3_normalize(v):
    l = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    r = 1.0 / l if l != 0.0 else math.inf
    return [v[0] * r, v[1] * r, v[2] * r]


def vec3_cross(a, b):
    return [a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]]


# ================================================================== easing.rs:6-44
def _cos(x):       # f64::cos / sin are total: NaN for +-inf (Python raises)
    return math.nan if math.isinf(x) or math.isnan(x) else math.cos(x)


def _sin(x):
    return math.nan if math.isinf(x) or math.isnan(x) else math.sin(x)


def _pow2(y):      # 2f64.powf(y)
    if math.isnan(y):
        return math.nan
    try:
        return math.pow(2.0, y)
    except OverflowError:
        return math.inf


def easing_in(t):
    return 1.0 - _cos(t * PI * 0.5)


def easing_out(t):
    return 1.0 - easing_in(1.0 - t)


def easing_in_out(t):
    return (1.0 - _cos(t * PI)) * 0.5


def easing_in_out_fast(t):
    return easing_in_out(easing_in_out(t))


def easing_plus_minus(t):
    t *= 2.0 * PI
    t2 = 2.0 * t
    return _sin(t) * (3.0 - _cos(t) - _cos(t2) - _cos(t) * _cos(t2)) / 4.0


def easing_elastic_out(x):
    c4 = (2.0 * PI) / 3.0
    if x == 0.0:
        return 0.0
    if x == 1.0:
        return 1.0
    return _pow2(-10.0 * x) * _sin((x * 10.0 - 0.75) * c4) + 1.0


# ==================================================================== scene model
class Scene:
    """Name-resolved scene at its *effective state*: the top-level `uniforms:` /
    `matrices:` of the file (scene_serialized.rs:1114-1136); stages/animations are
    not applied (render-frame without --stage never calls init_stage,
    /root/reference/src/main.rs:2876-2929)."""

    def __init__(self, ser: dict):
        self.ser = ser
        self.cam = ser["cam"]
        self.use_time = bool(ser.get("use_time", False))
        self.skybox = ser.get("skybox")
        self.time = 0.0
        self.total_time = 0.0
        self.camera_matrix = mat_identity()  # FormulasCache.3 (uniform.rs:633)

        # uniforms: id -> AnyUniform (Tagged); named in file order
        self.uniforms: list[Tagged] = []
        self.uniform_names: list[str | None] = []
        self.uniform_by_name: dict[str, int] = {}
        for item in ser["uniforms"][0]:
            self._add_uniform(item["data"], item["name"])

        # matrices, two passes (names first so forward references resolve)
        self.matrices: list[Any] = []
        self.matrix_names: list[str | None] = []
        self.matrix_by_name: dict[str, int] = {}
        for item in ser["matrices"][0]:
            self.matrix_by_name.setdefault(item["name"], len(self.matrices))
            self.matrices.append(None)
            self.matrix_names.append(item["name"])
        for idx, item in enumerate(ser["matrices"][0]):
            self.matrices[idx] = self._matrix_from_ser(item["data"])

        # objects
        self.objects = []
        for item in ser["objects"][0]:
            self.objects.append(self._object_from_ser(item["name"], item["data"]))

        self.textures = [(it["name"], it["data"][0]) for it in ser["textures"][0]]
        self.materials = [(it["name"], it["data"]) for it in ser["materials"][0]]
        self.intersection_materials = [
            (it["name"], it["data"][0][0][0]) for it in ser.get("intersection_materials", [[]])[0]
        ]
        self.library = [(it["name"], it["data"][0][0]) for it in ser["library"][0]]
        # videos (video.rs:14-20; scene_serialized.rs:1200-1210, after the library and before the stages): a video is one more
        # sampler2D named like a texture (scene.rs:405-409, 709-711) whose image the host replaces per frame, chosen by a uniform
        self.videos = []
        for it in ser.get("videos", [[]])[0]:
            d = it["data"]
            self.videos.append((it["name"], d.get("path", ""), self._uniform_ref(d.get("uniform"))))

        # animation stages + dev stage (scene_serialized.rs:1286-1371): inline elements are inserted at load time
        self.stages = {}
        for item in ser.get("animation_stages", [[]])[0]:
            d = item["data"]
            st = {"uniforms": {}, "matrices": {}}
            for uname, anim in d.get("uniforms", {}).items():
                if uname in self.uniform_by_name:
                    st["uniforms"][uname] = self._stage_anim(anim, self._uniform_ref)
            for mname, anim in d.get("matrices", {}).items():
                if mname in self.matrix_by_name:
                    st["matrices"][mname] = self._stage_anim(anim, self._matrix_ref)
            self.stages[item["name"]] = st
        dev = ser.get("dev_stage") or {}
        self.dev_uniforms = {k: v for k, v in dev.get("uniforms", {}).items() if k in self.uniform_by_name}
        self.dev_matrices = {k: self._matrix_from_ser(v) for k, v in dev.get("matrices", {}).items() if k in self.matrix_by_name}

    @staticmethod
    def _stage_anim(anim: Tagged, resolve):
        """StageAnimSer -> ('dev',) | ('set', id or None)."""
        if anim.tag in ("ProvidedToUser", "FromDev"):
            return ("dev",)
        if anim.tag in ("Changed", "ChangedAndToUser"):
            return ("set", resolve(anim.value[0]))
        raise ValueError(anim.tag)

    def init_stage(self, name: str):
        """Scene::init_stage(CurrentStage::Animation) (scene.rs:1180-1200) via StageChanging::init_stage
        (animation.rs:171-183): `Changed(Some(x))` copies element x over the target, `FromDev` /
        `ProvidedToUser` restore the dev-stage value.  (Camera selection is the caller's business.)"""
        st = self.stages[name]
        for uname, a in st["uniforms"].items():
            uid = self.uniform_by_name[uname]
            if a[0] == "set":
                if a[1] is not None:
                    self.uniforms[uid] = self.uniforms[a[1]]
            elif uname in self.dev_uniforms:
                self.uniforms[uid] = self.dev_uniforms[uname]
        for mname, a in st["matrices"].items():
            mid = self.matrix_by_name[mname]
            if a[0] == "set":
                if a[1] is not None:
                    self.matrices[mid] = self.matrices[a[1]]
            elif mname in self.dev_matrices:
                self.matrices[mid] = self.dev_matrices[mname]

    def video_frame_index(self, k: int, frame_count: int):
        """VideoRuntime::update (main.rs:862-895): which of `frame_count` frames video k shows in the scene's current state --
        round((count - 1) * clamp(uniform, 0, 1)), Rust's round (half away from zero); None when the video has no uniform,
        the uniform has no value, or there are no frames (the texture is then left as it is)."""
        uid = self.videos[k][2]
        if uid is None or frame_count <= 0:
            return None
        v = self.get_uniform(uid)
        if v is None:
            return None
        kind, val = v
        f = (1.0 if val else 0.0) if kind == "bool" else float(val)      # From<AnyUniformResult> for f64 (uniform.rs:298-311)
        f = min(max(f, 0.0), 1.0) if f == f else f
        x = (frame_count - 1) * f
        if x != x:                                                        # NaN: `NaN as usize` is 0
            return 0
        r = math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)
        return int(min(max(r, 0.0), float(frame_count - 1)))

    # ---------------------------------------------------------------- loading
    def _add_uniform(self, data, name=None) -> int:
        uid = len(self.uniforms)
        self.uniforms.append(data)
        self.uniform_names.append(name)
        if name is not None:
            self.uniform_by_name.setdefault(name, uid)
        return uid

    def _uniform_ref(self, ref):
        if ref is None:
            return None
        if ref.tag == "Named":
            return self.uniform_by_name.get(ref.value[0])
        if ref.tag == "Inline":
            return self._add_uniform(ref.value[0])
        raise ValueError(ref)

    def _param(self, p):
        """ParametrizeOrNot -> ('v', float) | ('u', uniform id or None)."""
        if p.tag == "Value":
            return ("v", float(p.value[0]))
        if p.tag == "Uniform":
            return ("u", self._uniform_ref(p.value[0]))
        raise ValueError(p)

    def _tvec(self, d, keys="xyz"):
        return [self._param(d[k]) for k in keys]

    def _matrix_ref(self, ref):
        if ref is None:
            return None
        if ref.tag == "Named":
            return self.matrix_by_name.get(ref.value[0])
        if ref.tag == "Inline":
            mid = len(self.matrices)
            self.matrices.append(None)
            self.matrix_names.append(None)
            self.matrices[mid] = self._matrix_from_ser(ref.value[0])
            return mid
        raise ValueError(ref)

    def _matrix_from_ser(self, m: Tagged):
        t, v = m.tag, m.value
        if t == "Mul":
            return ("Mul", self._matrix_ref(v["to"]), self._matrix_ref(v["what"]))
        if t == "Teleport":
            return ("Teleport", self._matrix_ref(v["first_portal"]), self._matrix_ref(v["second_portal"]),
                    self._matrix_ref(v["what"]))
        if t == "Simple":
            return ("Simple", [float(x) for x in v["offset"]], float(v["scale"]),
                    [float(x) for x in v["rotate"]], [bool(x) for x in v["mirror"]])
        if t == "Parametrized":
            return ("Parametrized", self._tvec(v["offset"]), self._tvec(v["rotate"]),
                    self._tvec(v["mirror"]), self._param(v["scale"]))
        if t == "Exact":
            return ("Exact", self._tvec(v["i"]), self._tvec(v["j"]), self._tvec(v["k"]), self._tvec(v["pos"]))
        if t == "ExactFull":
            return ("ExactFull", *[self._tvec(v[c], "xyzw") for c in ("c0", "c1", "c2", "c3")])
        if t == "If":
            return ("If", self._param(v["condition"]), self._matrix_ref(v["then"]), self._matrix_ref(v["otherwise"]))
        if t == "Sqrt":
            return ("Sqrt", self._matrix_ref(v[0]))
        if t == "Lerp":
            return ("Lerp", self._param(v["t"]), self._matrix_ref(v["first"]), self._matrix_ref(v["second"]))
        if t == "Camera":
            return ("Camera",)
        if t == "Inv":
            return ("Inv", self._matrix_ref(v[0]))
        raise ValueError(f"unknown matrix kind {t}")

    def _object_from_ser(self, name, o: Tagged):
        if o.tag == "DebugMatrix":
            return {"name": name, "class": "DebugMatrix", "matrix": self._matrix_ref(o.value[0])}
        v = o.value
        kind = v["kind"]
        if kind.tag == "Simple":
            mats = [self._matrix_ref(kind.value[0])]
        else:
            mats = [self._matrix_ref(kind.value[0]), self._matrix_ref(kind.value[1])]
        sub = v.get("in_subspace")
        sub = "Normal" if sub is None else sub.tag
        code = v["is_inside"][0][0] if o.tag == "Flat" else v["intersect"][0][0]
        return {"name": name, "class": o.tag, "kind": kind.tag, "matrices": mats, "in_subspace": sub, "code": code}

    # -------------------------------------------------------------- evaluation
    def matrix_name(self, mid: int) -> str:
        """object.rs:188-193: named -> name, inline -> 'id<N>' (N is this loader's id)."""
        n = self.matrix_names[mid]
        return n if n is not None else f"id{mid}"

    def get_uniform(self, uid, _visited=None):
        """uniform.rs:1009-1140 -> ('bool'|'int'|'float', value) or None."""
        if uid is None:
            return None
        visited = _visited if _visited is not None else []
        if uid in visited:  # storage2.rs:158-160 cycle guard
            return None
        visited.append(uid)
        try:
            u = self.uniforms[uid]
            t = u.tag
            if t == "Bool":
                return ("bool", bool(u.value[0]))
            if t == "Int":
                return ("int", int(u.value[0]["value"]))
            if t == "Float":
                return ("float", float(u.value[0]["value"]))
            if t in ("Angle", "Progress"):
                return ("float", float(u.value[0]))
            if t in ("Formula", "FormulaInt"):
                text = u.value[0][0]

                def ns(name, args):
                    return self._formula_ns(name, args, visited)

                try:
                    val = Evaluator(ns).eval_text(text)
                except FormulaError:
                    return None
                if t == "Formula":
                    return ("float", val)
                # Rust `as i32`: saturating, NaN -> 0
                if math.isnan(val):
                    return ("int", 0)
                return ("int", int(max(-2147483648.0, min(2147483647.0, math.trunc(val)))))
            if t == "TrefoilSpecial":
                # uniform.rs:20 `TrefoilSpecial([(bool, u8, u8); 18])`; packed per element as scene.rs:644-650 uploads it
                arr = u.value[0]
                while len(arr) == 1 and isinstance(arr[0], list) and arr[0] and isinstance(arr[0][0], list):
                    arr = arr[0]
                return ("trefoil", [int(v) + int(bool(en)) * 10000 + int(col) * 1000 for en, v, col in arr])
            raise ValueError(t)
        finally:
            visited.pop()

    def _formula_ns(self, name, a, visited):
        def arg(i):
            if i >= len(a):
                raise FormulaError(f"{name}: missing argument {i}")
            return a[i]

        def truthy(x):
            return abs(x - 1.0) < 1e-6

        if name == "if":
            return arg(1) if truthy(arg(0)) else arg(2)
        if name == "and":
            return 1.0 if truthy(arg(0)) and truthy(arg(1)) else 0.0
        if name == "or":
            return 1.0 if truthy(arg(0)) or truthy(arg(1)) else 0.0
        if name == "not":
            return 0.0 if truthy(arg(0)) else 1.0
        if name == "deg2rad":
            return arg(0) / 180.0 * PI
        if name == "rad2deg":
            return arg(0) * 180.0 / PI
        if name == "switch":
            idx = arg(0)
            idx = 0 if math.isnan(idx) or idx < 0 else int(idx)  # `as usize` saturates
            return arg(idx)
        if name == "on":
            v, lo, hi = arg(0), arg(1), arg(2)
            if v < lo:
                return 0.0
            if v > hi:
                return 1.0
            return (v - lo) / (hi - lo)
        if name == "inv":
            return 1.0 - arg(0)
        if name == "sqrt":
            x = arg(0)
            return math.sqrt(x) if x >= 0 else math.nan
        if name == "atan2":
            return math.atan2(arg(0), arg(1))
        if name == "time":
            return self.time
        if name == "total_time":
            return self.total_time
        if name == "easing_linear":
            return arg(0)
        if name == "easing_in":
            return easing_in(arg(0))
        if name == "easing_out":
            return easing_out(arg(0))
        if name == "easing_in_out":
            return easing_in_out(arg(0))
        if name == "easing_in_out_fast":
            return easing_in_out_fast(arg(0))
        if name == "easing_plus_minus":
            return easing_plus_minus(arg(0))
        if name == "easing_elastic_out":
            return easing_elastic_out(arg(0))
        if name == "bump":
            x = (arg(0) - arg(1)) / arg(2)
            return 0.5 * (1.0 + math.cos(PI * x)) if abs(x) < 1.0 else 0.0
        if name == "later_start":
            t = arg(0)
            time = 1.0 - arg(1)
            return max(0.0, t / time - (1.0 - time) / time)
        if name == "early_finish":
            return min(1.0, arg(0) / arg(1))
        if name == "lerp":
            lo, hi, t = arg(0), arg(1), arg(2)
            return (1.0 - t) * lo + t * hi  # emath::lerp (egui 0.31 re-export): (1-t)*start + t*end
        uid = self.uniform_by_name.get(name)
        if uid is None:
            return None
        r = self.get_uniform(uid, visited)
        if r is None:
            return None
        return -1.0 if r[0] == "trefoil" else float(r[1])      # uniform.rs:300-311 (`From<AnyUniformResult> for f64`)

    def _param_get(self, p):
        if p[0] == "v":
            return p[1]
        r = self.get_uniform(p[1])
        return None if r is None else float(r[1])

    def get_matrix(self, mid, _visited=None):
        """matrix.rs:510-631. Returns column-major 4x4 (list of 4 columns) or None."""
        if mid is None:
            return None
        visited = _visited if _visited is not None else []
        if mid in visited:
            return None
        visited.append(mid)
        try:
            m = self.matrices[mid]
            k = m[0]
            g = lambda i: self.get_matrix(i, visited)  # noqa: E731
            if k == "Mul":
                to, what = g(m[1]), g(m[2])
                if to is None or what is None:
                    return None
                return mat_mul(what, to)
            if k == "Teleport":
                first, second, what = g(m[1]), g(m[2]), g(m[3])
                if first is None or second is None or what is None:
                    return None
                return mat_mul(mat_mul(second, mat_inverse(first)), what)
            if k == "Simple":
                _, offset, scale, rotate, mirror = m
                s3 = [scale * (-1.0 if mirror[i] else 1.0) for i in range(3)]
                return srt_matrix(s3, rotate, offset)
            if k == "Parametrized":
                _, offset, rotate, mirror, scale = m
                sc = self._param_get(scale)
                if sc is None:
                    return None
                mir = [self._param_get(p) for p in mirror]
                if any(x is None for x in mir):
                    return None
                s3 = [sc * (1.0 - 2.0 * x) for x in mir]
                rot = [self._param_get(p) for p in rotate]
                if any(x is None for x in rot):
                    return None
                off = [self._param_get(p) for p in offset]
                if any(x is None for x in off):
                    return None
                return srt_matrix(s3, rot, off)
            if k == "Exact":
                cols = []
                for vec, w in zip(m[1:5], (0.0, 0.0, 0.0, 1.0)):
                    c = [self._param_get(p) for p in vec]
                    if any(x is None for x in c):
                        return None
                    cols.append(c + [w])
                return cols
            if k == "ExactFull":
                cols = []
                for vec in m[1:5]:
                    c = [self._param_get(p) for p in vec]
                    if any(x is None for x in c):
                        return None
                    cols.append(c)
                return cols
            if k == "If":
                c = self._param_get(m[1])
                if c is None:
                    return None
                return g(m[2]) if c > 0.5 else g(m[3])
            if k == "Inv":
                x = g(m[1])
                return None if x is None else mat_inverse(x)
            if k == "Camera":
                return self.camera_matrix
            if k == "Lerp":
                t = self._param_get(m[1])
                if t is None:
                    return None
                first, second = g(m[2]), g(m[3])
                if first is None or second is None:
                    return None
                return mat_lerp(first, second, t)
            if k == "Sqrt":
                # matrix.rs:606-613: None ("Can't calculate sqrt!") when the minimisation does not reach cost < 1e-4
                x = g(m[1])
                return None if x is None else mat_sqrt(x)
            raise ValueError(k)
        finally:
            visited.pop()

    # ------------------------------------------------------------ uniform table
    def object_matrix_ids(self):
        out = []
        for o in self.objects:
            if o["class"] == "DebugMatrix":
                if o["matrix"] is not None:
                    out.append(o["matrix"])
            else:
                if all(x is not None for x in o["matrices"]):
                    out.extend(o["matrices"])
        return out

    def uniform_table(self):
        """scene.rs:545-658 -> ordered dict name -> ('mat4', 16 f64 column-major) |
        ('float', f64) | ('int', int). Values are f64; the f32 cast happens at upload."""
        table = {}
        ids = self.object_matrix_ids() + [i for i, n in enumerate(self.matrix_names) if n is not None]
        for mid in ids:
            name = self.matrix_name(mid)
            m = self.get_matrix(mid)
            if m is None:
                continue
            table[f"{name}_mat"] = ("mat4", [x for col in m for x in col])
            table[f"{name}_mat_inv"] = ("mat4", [x for col in mat_inverse(m) for x in col])
        for o in self.objects:
            if o["class"] == "DebugMatrix" or o["kind"] != "Portal":
                continue
            a, b = o["matrices"]
            if a is None or b is None:
                continue
            ma, mb = self.get_matrix(a), self.get_matrix(b)
            if ma is None or mb is None:
                continue
            na, nb = self.matrix_name(a), self.matrix_name(b)
            table[f"{na}_to_{nb}_mat_teleport"] = ("mat4", [x for col in mat_mul(mb, mat_inverse(ma)) for x in col])
            if na != nb:
                table[f"{nb}_to_{na}_mat_teleport"] = ("mat4", [x for col in mat_mul(ma, mat_inverse(mb)) for x in col])
        for uid, name in enumerate(self.uniform_names):
            if name is None:
                continue
            r = self.get_uniform(uid)
            if r is None:
                continue
            if r[0] == "trefoil":
                for i, packed in enumerate(r[1]):                # scene.rs:488-492: `ts_<i>_<name>_u`, one int each
                    table[f"ts_{i}_{name}_u"] = ("int", packed)
            elif r[0] == "bool":
                table[f"{name}_u"] = ("int", int(r[1]))
            elif r[0] == "int":
                table[f"{name}_u"] = ("int", int(r[1]))
            else:
                table[f"{name}_u"] = ("float", float(r[1]))
        return table

    def material_ids(self):
        """scene.rs:720-842: user materials in order from USER_MATERIAL_OFFSET (10),
        then teleport_<pos>_1_M / teleport_<pos>_2_M per portal object in object order."""
        out = {}
        counter = 10
        for name, _ in self.materials:
            out[f"{name}_M"] = counter
            counter += 1
        for pos, o in enumerate(self.objects):
            if o["class"] != "DebugMatrix" and o["kind"] == "Portal" and all(x is not None for x in o["matrices"]):
                out[f"teleport_{pos}_1_M"] = counter
                counter += 1
                out[f"teleport_{pos}_2_M"] = counter
                counter += 1
        return out


# ========================================================================= camera
def orbit_camera_matrix(look_at, alpha, beta, r, teleport=None, free_movement=False):
    """main.rs:278-304 RotateAroundCam::get_pos_vec / get_matrix (f64)."""
    pv = [math.sin(beta) * math.cos(alpha) * r, math.cos(beta) * r, math.sin(beta) * math.sin(alpha) * r]
    pos = [pv[i] + look_at[i] for i in range(3)]
    k = vec3_normalize([look_at[i] - pos[i] for i in range(3)])
    i_ = vec3_normalize(vec3_cross(k, [0.0, 1.0, 0.0]))
    j = vec3_normalize(vec3_cross(k, i_))
    last = look_at if free_movement else pos
    m = [[i_[0], i_[1], i_[2], 0.0], [j[0], j[1], j[2], 0.0], [k[0], k[1], k[2], 0.0],
         [last[0], last[1], last[2], 1.0]]
    if teleport is None:
        teleport = mat_identity()
    return mat_mul(teleport, m)


def camera_scale(m):
    """main.rs:1325-1333 calc_scale: mean length of the first three columns."""
    s = 0.0
    for c in m[:3]:
        s += math.sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3])
    return s / 3.0


def renderer_defaults():
    """SceneRenderer::new defaults, main.rs:1021-1047, + RotateAroundCam::new :94-130."""
    return {
        "view_angle": 90.0 / 180.0 * PI,  # deg2rad(90.)
        "t_start": 10.0,
        "t_end": 210.0,
        "darken_by_distance": 1,
        "angle_color_disable": 0,
        "grid_disable": 0,
        "black_border_disable": 0,
        "camera_in_subspace": 0,
        "aa_count": 1,
        "aa_start": 0,
    }


# ====================================================================== scene IR
def _material_ir(name, m: Tagged):
    v = m.value
    if m.tag == "Simple":
        return {"name": name, "type": "Simple", "color": [float(x) for x in v["color"]],
                "normal_coef": float(v["normal_coef"]), "grid": bool(v["grid"]),
                "grid_scale": float(v["grid_scale"]), "grid_coef": float(v["grid_coef"]),
                "grid2": bool(v.get("grid2", False)), "grid3": bool(v.get("grid3", False))}
    if m.tag == "Reflect":
        return {"name": name, "type": "Reflect", "add_to_color": [float(x) for x in v["add_to_color"]]}
    if m.tag == "Refract":
        return {"name": name, "type": "Refract", "add_to_color": [float(x) for x in v["add_to_color"]],
                "refractive_index": float(v["refractive_index"])}
    if m.tag == "Complex":
        return {"name": name, "type": "Complex", "code": v["code"][0][0]}
    raise ValueError(m.tag)


def scene_ir(scene: Scene, scene_name: str, time: float = 0.0, stage: str | None = None) -> dict:
    if stage is not None:
        scene.init_stage(stage)
    scene.time = time
    scene.total_time = time
    objects = []
    for o in scene.objects:
        if o["class"] == "DebugMatrix":
            objects.append({"name": o["name"], "class": "DebugMatrix",
                            "matrices": [None if o["matrix"] is None else scene.matrix_name(o["matrix"])]})
        else:
            objects.append({"name": o["name"], "class": o["class"], "kind": o["kind"],
                            "matrices": [None if x is None else scene.matrix_name(x) for x in o["matrices"]],
                            "in_subspace": o["in_subspace"], "code": o["code"]})
    cam = scene.cam
    cm = orbit_camera_matrix([float(x) for x in cam["look_at"]], float(cam["alpha"]), float(cam["beta"]), float(cam["r"]))
    # send_camera_object_matrix (main.rs:145, 1432-1436, 1530-1534): the `Camera` matrix kind sees the renderer's camera
    scene.camera_matrix = cm
    table = scene.uniform_table()
    d = renderer_defaults()
    return {
        "format": "portal-b200 scene IR v1",
        "scene": scene_name,
        "source": f"scenes/{scene_name}.ron of optozorax/portal @ 7234f953 (effective state, time={time})",
        "cam": {"look_at": [float(x) for x in cam["look_at"]], "alpha": float(cam["alpha"]),
                "beta": float(cam["beta"]), "r": float(cam["r"]),
                "offset_after_material": float(cam["offset_after_material"])},
        "camera_matrix": [x for col in cm for x in col],
        "camera_scale": camera_scale(cm),
        "renderer": d,
        "skybox": scene.skybox,
        # samplers = textures and videos, each name once (the BTreeSet of scene.rs:703-716); a video has no image file:
        # its frames come from the caller (`video_frame_index` says which)
        "textures": [{"name": n, "path": p} for n, p in scene.textures] +
                    [{"name": n, "path": None, "video": True} for n, _, _ in scene.videos if n not in {t for t, _ in scene.textures}],
        "videos": [{"name": n, "path": p, "uniform": (scene.uniform_names[u] if u is not None else None)} for n, p, u in scene.videos],
        "materials": [_material_ir(n, m) for n, m in scene.materials],
        "material_ids": scene.material_ids(),
        "objects": objects,
        "intersection_materials": [{"name": n, "code": c} for n, c in scene.intersection_materials],
        "library": [{"name": n, "code": c} for n, c in scene.library],
        "uniforms": {k: {"type": t, "value": v} for k, (t, v) in table.items()},
    }


def load_scene(path: str) -> Scene:
    return Scene(ron.load(path))
