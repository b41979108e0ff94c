#include "ph_scene.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <stdexcept>

#include "../../../include/portal_b200.h"

namespace ph {

// ------------------------------------------------------------------------------ glam arithmetic
// DMat4 as 16 doubles, column-major: m[4*c + r].
Mat4 mat_identity() {
    Mat4 m{};
    m[0] = m[5] = m[10] = m[15] = 1.0;
    return m;
}

static inline void mul_vec(const Mat4& m, const double v[4], double out[4]) {
    // glam: x_axis*v.x + y_axis*v.y + z_axis*v.z + w_axis*v.w, left to right
    for (int r = 0; r < 4; r++) {
        double acc = m[r] * v[0];
        acc = acc + m[4 + r] * v[1];
        acc = acc + m[8 + r] * v[2];
        acc = acc + m[12 + r] * v[3];
        out[r] = acc;
    }
}

Mat4 mat_mul(const Mat4& a, const Mat4& b) {
    Mat4 o{};
    for (int c = 0; c < 4; c++) mul_vec(a, &b[4 * c], &o[4 * c]);
    return o;
}

// glam 0.13.1 DMat4::inverse: cofactor expansion (GLM formulation) scaled by 1/det; a zero
// determinant is NOT trapped (glam_assert is a debug feature), so singular input yields Inf/NaN.
Mat4 mat_inverse(const Mat4& m) {
    const double m00 = m[0], m01 = m[1], m02 = m[2], m03 = m[3];
    const double m10 = m[4], m11 = m[5], m12 = m[6], m13 = m[7];
    const double m20 = m[8], m21 = m[9], m22 = m[10], m23 = m[11];
    const double m30 = m[12], m31 = m[13], m32 = m[14], m33 = m[15];

    const double coef00 = m22 * m33 - m32 * m23, coef02 = m12 * m33 - m32 * m13, coef03 = m12 * m23 - m22 * m13;
    const double coef04 = m21 * m33 - m31 * m23, coef06 = m11 * m33 - m31 * m13, coef07 = m11 * m23 - m21 * m13;
    const double coef08 = m21 * m32 - m31 * m22, coef10 = m11 * m32 - m31 * m12, coef11 = m11 * m22 - m21 * m12;
    const double coef12 = m20 * m33 - m30 * m23, coef14 = m10 * m33 - m30 * m13, coef15 = m10 * m23 - m20 * m13;
    const double coef16 = m20 * m32 - m30 * m22, coef18 = m10 * m32 - m30 * m12, coef19 = m10 * m22 - m20 * m12;
    const double coef20 = m20 * m31 - m30 * m21, coef22 = m10 * m31 - m30 * m11, coef23 = m10 * m21 - m20 * m11;

    const double fac0[4] = {coef00, coef00, coef02, coef03}, fac1[4] = {coef04, coef04, coef06, coef07};
    const double fac2[4] = {coef08, coef08, coef10, coef11}, fac3[4] = {coef12, coef12, coef14, coef15};
    const double fac4[4] = {coef16, coef16, coef18, coef19}, fac5[4] = {coef20, coef20, coef22, coef23};
    const double vec0[4] = {m10, m00, m00, m00}, vec1[4] = {m11, m01, m01, m01};
    const double vec2[4] = {m12, m02, m02, m02}, vec3[4] = {m13, m03, m03, m03};
    const double sign_a[4] = {1.0, -1.0, 1.0, -1.0}, sign_b[4] = {-1.0, 1.0, -1.0, 1.0};

    Mat4 inv{};
    for (int k = 0; k < 4; k++) {
        inv[k] = (vec1[k] * fac0[k] - vec2[k] * fac1[k] + vec3[k] * fac2[k]) * sign_a[k];
        inv[4 + k] = (vec0[k] * fac0[k] - vec2[k] * fac3[k] + vec3[k] * fac4[k]) * sign_b[k];
        inv[8 + k] = (vec0[k] * fac1[k] - vec1[k] * fac3[k] + vec3[k] * fac5[k]) * sign_a[k];
        inv[12 + k] = (vec0[k] * fac2[k] - vec1[k] * fac4[k] + vec2[k] * fac5[k]) * sign_b[k];
    }
    const double d0 = m00 * inv[0], d1 = m01 * inv[4], d2 = m02 * inv[8], d3 = m03 * inv[12];
    const double det = d0 + d1 + d2 + d3;
    const double rcp = 1.0 / det;
    for (double& x : inv) x = x * rcp;
    return inv;
}

static void quat_mul(const double a[4], const double b[4], double o[4]) {
    const double x0 = a[0], y0 = a[1], z0 = a[2], w0 = a[3], x1 = b[0], y1 = b[1], z1 = b[2], w1 = b[3];
    o[0] = w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1;
    o[1] = w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1;
    o[2] = w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1;
    o[3] = w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1;
}

// glam 0.13.1 DMat4::from_scale_rotation_translation
static Mat4 mat_from_srt_quat(const double scale[3], const double q[4], const double offset[3]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double x2 = x + x, y2 = y + y, z2 = z + z;
    const double xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2;
    const double wx = w * x2, wy = w * y2, wz = w * z2;
    Mat4 m{};
    m[0] = (1.0 - (yy + zz)) * scale[0]; m[1] = (xy + wz) * scale[0]; m[2] = (xz - wy) * scale[0]; m[3] = 0.0 * scale[0];
    m[4] = (xy - wz) * scale[1]; m[5] = (1.0 - (xx + zz)) * scale[1]; m[6] = (yz + wx) * scale[1]; m[7] = 0.0 * scale[1];
    m[8] = (xz + wy) * scale[2]; m[9] = (yz - wx) * scale[2]; m[10] = (1.0 - (xx + yy)) * scale[2]; m[11] = 0.0 * scale[2];
    m[12] = offset[0]; m[13] = offset[1]; m[14] = offset[2]; m[15] = 1.0;
    return m;
}

// matrix.rs:537-547: DMat4::from_scale_rotation_translation(scale, Rx * Ry * Rz, offset)
Mat4 mat_srt(const double scale[3], const double rotate[3], const double offset[3]) {
    const double qx[4] = {std::sin(rotate[0] * 0.5), 0.0, 0.0, std::cos(rotate[0] * 0.5)};
    const double qy[4] = {0.0, std::sin(rotate[1] * 0.5), 0.0, std::cos(rotate[1] * 0.5)};
    const double qz[4] = {0.0, 0.0, std::sin(rotate[2] * 0.5), std::cos(rotate[2] * 0.5)};
    double qxy[4], q[4];
    quat_mul(qx, qy, qxy);
    quat_mul(qxy, qz, q);
    return mat_from_srt_quat(scale, q, offset);
}

// glam 0.13.1 Matrix4x4::determinant
static double mat_determinant(const Mat4& m) {
    const double m00 = m[0], m01 = m[1], m02 = m[2], m03 = m[3];
    const double m10 = m[4], m11 = m[5], m12 = m[6], m13 = m[7];
    const double m20 = m[8], m21 = m[9], m22 = m[10], m23 = m[11];
    const double m30 = m[12], m31 = m[13], m32 = m[14], m33 = m[15];
    const double a2323 = m22 * m33 - m23 * m32, a1323 = m21 * m33 - m23 * m31, a1223 = m21 * m32 - m22 * m31;
    const double a0323 = m20 * m33 - m23 * m30, a0223 = m20 * m32 - m22 * m30, a0123 = m20 * m31 - m21 * m30;
    return m00 * (m11 * a2323 - m12 * a1323 + m13 * a1223) - m01 * (m10 * a2323 - m12 * a0323 + m13 * a0223) +
           m02 * (m10 * a1323 - m11 * a0323 + m13 * a0123) - m03 * (m10 * a1223 - m11 * a0223 + m12 * a0123);
}

// glam 0.13.1 Quaternion::from_rotation_axes: branch on the largest component
static void quat_from_rotation_axes(const double xa[3], const double ya[3], const double za[3], double q[4]) {
    const double m00 = xa[0], m01 = xa[1], m02 = xa[2], m10 = ya[0], m11 = ya[1], m12 = ya[2], m20 = za[0], m21 = za[1], m22 = za[2];
    if (m22 <= 0.0) {
        const double dif10 = m11 - m00, omm22 = 1.0 - m22;
        if (dif10 <= 0.0) {
            const double four_xsq = omm22 - dif10, inv4x = 0.5 / std::sqrt(four_xsq);
            q[0] = four_xsq * inv4x; q[1] = (m01 + m10) * inv4x; q[2] = (m02 + m20) * inv4x; q[3] = (m12 - m21) * inv4x;
        } else {
            const double four_ysq = omm22 + dif10, inv4y = 0.5 / std::sqrt(four_ysq);
            q[0] = (m01 + m10) * inv4y; q[1] = four_ysq * inv4y; q[2] = (m12 + m21) * inv4y; q[3] = (m20 - m02) * inv4y;
        }
    } else {
        const double sum10 = m11 + m00, opm22 = 1.0 + m22;
        if (sum10 <= 0.0) {
            const double four_zsq = opm22 - sum10, inv4z = 0.5 / std::sqrt(four_zsq);
            q[0] = (m02 + m20) * inv4z; q[1] = (m12 + m21) * inv4z; q[2] = four_zsq * inv4z; q[3] = (m01 - m10) * inv4z;
        } else {
            const double four_wsq = opm22 + sum10, inv4w = 0.5 / std::sqrt(four_wsq);
            q[0] = (m12 - m21) * inv4w; q[1] = (m20 - m02) * inv4w; q[2] = (m01 - m10) * inv4w; q[3] = four_wsq * inv4w;
        }
    }
}

// glam 0.13.1 DMat4::to_scale_rotation_translation
static void mat_to_srt(const Mat4& m, double scale[3], double q[4], double t[3]) {
    const double det = mat_determinant(m);
    const double sign = std::isnan(det) ? det : std::copysign(1.0, det);
    auto length = [&](int c) { return std::sqrt(m[4 * c] * m[4 * c] + m[4 * c + 1] * m[4 * c + 1] + m[4 * c + 2] * m[4 * c + 2] + m[4 * c + 3] * m[4 * c + 3]); };
    scale[0] = length(0) * sign; scale[1] = length(1); scale[2] = length(2);
    double ax[3][3];
    for (int c = 0; c < 3; c++) {
        const double inv = 1.0 / scale[c];
        for (int r = 0; r < 3; r++) ax[c][r] = m[4 * c + r] * inv;
    }
    quat_from_rotation_axes(ax[0], ax[1], ax[2], q);
    t[0] = m[12]; t[1] = m[13]; t[2] = m[14];
}

// matrix.rs:614-628: lerp of (scale, rotation [shortest-arc nlerp], translation)
Mat4 mat_lerp(const Mat4& first, const Mat4& second, double t) {
    double fs[3], fq[4], ft[3], ss[3], sq[4], st[3], s[3], q[4], o[3];
    mat_to_srt(first, fs, fq, ft);
    mat_to_srt(second, ss, sq, st);
    for (int i = 0; i < 3; i++) {
        s[i] = fs[i] + ((ss[i] - fs[i]) * t);
        o[i] = ft[i] + ((st[i] - ft[i]) * t);
    }
    const double dot = fq[0] * sq[0] + fq[1] * sq[1] + fq[2] * sq[2] + fq[3] * sq[3];
    const double bias = dot >= 0.0 ? 1.0 : -1.0;
    for (int i = 0; i < 4; i++) q[i] = fq[i] + (((sq[i] * bias) - fq[i]) * t);
    const double r = 1.0 / std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) q[i] = q[i] * r;
    return mat_from_srt_quat(s, q, o);
}

static void v3_normalize(double v[3]) {
    const double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const double r = 1.0 / l;
    v[0] *= r; v[1] *= r; v[2] *= r;
}
static void v3_cross(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - b[1] * a[2];
    o[1] = a[2] * b[0] - b[2] * a[0];
    o[2] = a[0] * b[1] - b[0] * a[1];
}

// RotateAroundCam::get_pos_vec / get_matrix (main.rs:278-304), teleport_matrix = identity, not free_movement
Mat4 orbit_camera_matrix(const double look_at[3], double alpha, double beta, double r) {
    const double pv[3] = {std::sin(beta) * std::cos(alpha) * r, std::cos(beta) * r, std::sin(beta) * std::sin(alpha) * r};
    const double pos[3] = {pv[0] + look_at[0], pv[1] + look_at[1], pv[2] + look_at[2]};
    double k[3] = {look_at[0] - pos[0], look_at[1] - pos[1], look_at[2] - pos[2]};
    v3_normalize(k);
    const double up[3] = {0.0, 1.0, 0.0};
    double i[3], j[3];
    v3_cross(k, up, i);
    v3_normalize(i);
    v3_cross(k, i, j);
    v3_normalize(j);
    Mat4 m{};
    m[0] = i[0]; m[1] = i[1]; m[2] = i[2];
    m[4] = j[0]; m[5] = j[1]; m[6] = j[2];
    m[8] = k[0]; m[9] = k[1]; m[10] = k[2];
    m[12] = pos[0]; m[13] = pos[1]; m[14] = pos[2]; m[15] = 1.0;
    return mat_mul(mat_identity(), m);
}

Mat4 orbit_camera_matrix(const double look_at[3], double alpha, double beta, double r, const Mat4& teleport, bool free_movement) {
    const double pv[3] = {std::sin(beta) * std::cos(alpha) * r, std::cos(beta) * r, std::sin(beta) * std::sin(alpha) * r};
    const double pos[3] = {pv[0] + look_at[0], pv[1] + look_at[1], pv[2] + look_at[2]};
    double k[3] = {look_at[0] - pos[0], look_at[1] - pos[1], look_at[2] - pos[2]};
    v3_normalize(k);
    const double up[3] = {0.0, 1.0, 0.0};
    double i[3], j[3];
    v3_cross(k, up, i);
    v3_normalize(i);
    v3_cross(k, i, j);
    v3_normalize(j);
    const double* last = free_movement ? look_at : pos;
    Mat4 m{};
    m[0] = i[0]; m[1] = i[1]; m[2] = i[2];
    m[4] = j[0]; m[5] = j[1]; m[6] = j[2];
    m[8] = k[0]; m[9] = k[1]; m[10] = k[2];
    m[12] = last[0]; m[13] = last[1]; m[14] = last[2]; m[15] = 1.0;
    return mat_mul(teleport, m);
}

void mat_mul_vec(const Mat4& m, const double v[4], double out[4]) { mul_vec(m, v, out); }

// calc_scale (main.rs:1325-1333)
double camera_scale(const Mat4& m) {
    double s = 0.0;
    for (int c = 0; c < 3; c++) s += std::sqrt(m[4 * c] * m[4 * c] + m[4 * c + 1] * m[4 * c + 1] + m[4 * c + 2] * m[4 * c + 2] + m[4 * c + 3] * m[4 * c + 3]);
    return s / 3.0;
}

// ------------------------------------------------------------------------------ loading
namespace {

struct Loader {
    Scene& sc;
    explicit Loader(Scene& s) : sc(s) {}
    bool ok = true;
    void fail(const std::string& m) {
        if (ok) sc.error = m;
        ok = false;
    }

    // SerStorage<T>(Vec<Named<T>>) -> the inner list
    const RonValue* storage(const RonValue& root, const char* field, bool required) {
        const RonValue* v = root.get(field);
        if (!v) {
            if (required) fail(std::string("scene has no `") + field + "` section");
            return nullptr;
        }
        if (v->kind == RonValue::List && v->items.size() == 1 && v->items[0]->kind == RonValue::List) return v->items[0].get();
        if (v->kind == RonValue::List && v->items.empty()) return v;
        fail(std::string("`") + field + "`: expected ([...])");
        return nullptr;
    }

    static const RonValue* unwrap1(const RonValue* v) {  // newtype struct `(x)` -> x
        while (v && v->kind == RonValue::List && v->items.size() == 1) v = v->items[0].get();
        return v;
    }

    int add_uniform(const RonValue& data, const std::string& name) {
        Uniform u;
        if (data.kind != RonValue::Tagged) {
            fail("uniform `" + name + "`: expected an enum variant");
            return -1;
        }
        const RonValue* p = data.at(0);
        if (data.s == "Bool") { u.kind = Uniform::Bool; u.b = p && p->b; }
        else if (data.s == "Int") { u.kind = Uniform::Int; const RonValue* v = unwrap1(p); const RonValue* x = v ? v->get("value") : nullptr; u.i = x ? int(x->num()) : 0; }
        else if (data.s == "Float") { u.kind = Uniform::Float; const RonValue* v = unwrap1(p); const RonValue* x = v ? v->get("value") : nullptr; u.f = x ? x->num() : 0.0; }
        else if (data.s == "Angle") { u.kind = Uniform::Angle; u.f = p ? p->num() : 0.0; }
        else if (data.s == "Progress") { u.kind = Uniform::Progress; u.f = p ? p->num() : 0.0; }
        else if (data.s == "Formula" || data.s == "FormulaInt") {
            u.kind = data.s == "Formula" ? Uniform::Formula : Uniform::FormulaInt;
            const RonValue* t = unwrap1(p);
            u.text = t && t->kind == RonValue::String ? t->s : "";
        } else if (data.s == "TrefoilSpecial") {
            u.kind = Uniform::Trefoil;
            const RonValue* arr = p;
            while (arr && arr->kind == RonValue::List && arr->items.size() == 1 && arr->items[0]->kind == RonValue::List &&
                   !arr->items[0]->items.empty() && arr->items[0]->items[0]->kind == RonValue::List)
                arr = arr->items[0].get();
            if (arr && arr->kind == RonValue::List)
                for (size_t k = 0; k < arr->items.size() && k < 18; k++) {
                    const RonValue& e = *arr->items[k];
                    if (e.kind != RonValue::List || e.items.size() != 3) continue;
                    const bool enabled = e.items[0]->kind == RonValue::Bool && e.items[0]->b;
                    u.trefoil[k] = int(e.items[1]->num()) + (enabled ? 10000 : 0) + int(e.items[2]->num()) * 1000;
                }
        } else u.kind = Uniform::Unsupported;
        int id = int(sc.uniforms.size());
        sc.uniforms.push_back(u);
        sc.uniform_names.push_back(name);
        if (!name.empty() && !sc.uniform_by_name.count(name)) sc.uniform_by_name[name] = id;
        return id;
    }

    int uniform_ref(const RonValue* ref) {
        if (!ref || ref->kind == RonValue::Null) return -1;
        if (ref->is_tag("Named")) {
            auto it = sc.uniform_by_name.find(ref->at(0) ? ref->at(0)->s : "");
            return it == sc.uniform_by_name.end() ? -1 : it->second;
        }
        if (ref->is_tag("Inline") && ref->at(0)) return add_uniform(*ref->at(0), "");
        return -1;
    }

    Param param(const RonValue* p) {
        Param out;
        if (!p) return out;
        if (p->is_tag("Value")) out.value = p->at(0) ? p->at(0)->num() : 0.0;
        else if (p->is_tag("Uniform")) { out.is_uniform = true; out.uniform = uniform_ref(p->at(0)); }
        return out;
    }

    int matrix_ref(const RonValue* ref) {
        if (!ref || ref->kind == RonValue::Null) return -1;
        if (ref->is_tag("Named")) {
            auto it = sc.matrix_by_name.find(ref->at(0) ? ref->at(0)->s : "");
            return it == sc.matrix_by_name.end() ? -1 : it->second;
        }
        if (ref->is_tag("Inline") && ref->at(0)) {
            int id = int(sc.matrices.size());
            sc.matrices.emplace_back();
            sc.matrix_names.emplace_back();
            Matrix m = matrix_from(*ref->at(0));
            sc.matrices[id] = m;
            return id;
        }
        return -1;
    }

    void vec3(const RonValue* v, double out[3]) {
        for (int k = 0; k < 3; k++) out[k] = (v && v->at(k)) ? v->at(k)->num() : 0.0;
    }
    void tvec(const RonValue* v, Param* out, int n) {
        static const char* names[4] = {"x", "y", "z", "w"};
        for (int k = 0; k < n; k++) out[k] = param(v ? v->get(names[k]) : nullptr);
    }

    Matrix matrix_from(const RonValue& m) {
        Matrix o;
        if (m.kind != RonValue::Tagged) { fail("matrix: expected an enum variant"); return o; }
        const std::string& t = m.s;
        if (t == "Mul") { o.kind = Matrix::Mul; o.a = matrix_ref(m.get("to")); o.b = matrix_ref(m.get("what")); }
        else if (t == "Teleport") { o.kind = Matrix::Teleport; o.a = matrix_ref(m.get("first_portal")); o.b = matrix_ref(m.get("second_portal")); o.c = matrix_ref(m.get("what")); }
        else if (t == "Simple") {
            o.kind = Matrix::Simple;
            vec3(m.get("offset"), o.offset);
            vec3(m.get("rotate"), o.rotate);
            o.scale = m.get("scale") ? m.get("scale")->num() : 1.0;
            const RonValue* mir = m.get("mirror");
            for (int k = 0; k < 3; k++) o.mirror[k] = mir && mir->at(k) && mir->at(k)->b;
        } else if (t == "Parametrized") {
            o.kind = Matrix::Parametrized;
            tvec(m.get("offset"), o.p + 0, 3);
            tvec(m.get("rotate"), o.p + 3, 3);
            tvec(m.get("mirror"), o.p + 6, 3);
            o.p[9] = param(m.get("scale"));
        } else if (t == "Exact") {
            o.kind = Matrix::Exact;
            tvec(m.get("i"), o.p + 0, 3); tvec(m.get("j"), o.p + 3, 3); tvec(m.get("k"), o.p + 6, 3); tvec(m.get("pos"), o.p + 9, 3);
        } else if (t == "ExactFull") {
            o.kind = Matrix::ExactFull;
            tvec(m.get("c0"), o.p + 0, 4); tvec(m.get("c1"), o.p + 4, 4); tvec(m.get("c2"), o.p + 8, 4); tvec(m.get("c3"), o.p + 12, 4);
        } else if (t == "If") { o.kind = Matrix::If; o.p[0] = param(m.get("condition")); o.a = matrix_ref(m.get("then")); o.b = matrix_ref(m.get("otherwise")); }
        else if (t == "Sqrt") { o.kind = Matrix::Sqrt; o.a = matrix_ref(m.at(0)); }
        else if (t == "Lerp") { o.kind = Matrix::Lerp; o.p[0] = param(m.get("t")); o.a = matrix_ref(m.get("first")); o.b = matrix_ref(m.get("second")); }
        else if (t == "Camera") o.kind = Matrix::Camera;
        else if (t == "Inv") { o.kind = Matrix::Inv; o.a = matrix_ref(m.at(0)); }
        else fail("unknown matrix kind `" + t + "`");
        return o;
    }

    static std::string code_of(const RonValue* v) {
        v = unwrap1(v);
        return v && v->kind == RonValue::String ? v->s : "";
    }

    void load(const RonValue& root) {
        if (root.kind != RonValue::Struct) { fail("scene file: expected a struct at top level"); return; }
        if (const RonValue* cam = root.get("cam")) {
            vec3(cam->get("look_at"), sc.look_at);
            if (cam->get("alpha")) sc.alpha = cam->get("alpha")->num();
            if (cam->get("beta")) sc.beta = cam->get("beta")->num();
            if (cam->get("r")) sc.r = cam->get("r")->num();
            if (cam->get("offset_after_material")) sc.offset_after_material = cam->get("offset_after_material")->num();
        }
        if (const RonValue* t = root.get("use_time")) sc.use_time = t->b;
        if (const RonValue* s = root.get("skybox")) sc.skybox = s->kind == RonValue::String ? s->s : "";

        // uniforms (scene_serialized.rs:1114-1117)
        if (const RonValue* us = storage(root, "uniforms", true))
            for (auto& it : us->items) {
                const RonValue* n = it->get("name");
                const RonValue* d = it->get("data");
                if (n && d) add_uniform(*d, n->s);
            }
        // matrices, two passes (:1119-1136)
        const RonValue* ms = storage(root, "matrices", true);
        if (ms) {
            for (auto& it : ms->items) {
                const RonValue* n = it->get("name");
                std::string name = n ? n->s : "";
                if (!sc.matrix_by_name.count(name)) sc.matrix_by_name[name] = int(sc.matrices.size());
                sc.matrices.emplace_back();
                sc.matrix_names.push_back(name);
            }
            for (size_t k = 0; k < ms->items.size(); k++) {
                const RonValue* d = ms->items[k]->get("data");
                if (d) {
                    Matrix m = matrix_from(*d);
                    sc.matrices[k] = m;
                }
            }
        }
        // objects (:1138-1177)
        if (const RonValue* os = storage(root, "objects", true))
            for (auto& it : os->items) {
                const RonValue* n = it->get("name");
                const RonValue* d = it->get("data");
                if (!n || !d || d->kind != RonValue::Tagged) continue;
                SceneObject o;
                o.name = n->s;
                if (d->s == "DebugMatrix") {
                    o.cls = SceneObject::DebugMatrix;
                    o.matrix_a = matrix_ref(d->at(0));
                    o.subspace = PE_SUBSPACE_BOTH;
                } else {
                    o.cls = d->s == "Flat" ? SceneObject::Flat : SceneObject::Complex;
                    const RonValue* kind = d->get("kind");
                    if (kind && kind->is_tag("Portal")) {
                        o.portal = true;
                        o.matrix_a = matrix_ref(kind->at(0));
                        o.matrix_b = matrix_ref(kind->at(1));
                    } else if (kind) {
                        o.matrix_a = matrix_ref(kind->at(0));
                    }
                    o.code = code_of(d->get(o.cls == SceneObject::Flat ? "is_inside" : "intersect"));
                    const RonValue* sub = d->get("in_subspace");
                    o.subspace = !sub ? PE_SUBSPACE_NORMAL : sub->is_tag("Subspace") ? PE_SUBSPACE_SUBSPACE : sub->is_tag("Both") ? PE_SUBSPACE_BOTH : PE_SUBSPACE_NORMAL;
                }
                sc.objects.push_back(o);
            }
        if (const RonValue* ts = storage(root, "textures", true))
            for (auto& it : ts->items)
                if (it->get("name")) sc.textures.emplace_back(it->get("name")->s, code_of(it->get("data")));
        if (const RonValue* mats = storage(root, "materials", true))
            for (auto& it : mats->items) {
                const RonValue* n = it->get("name");
                const RonValue* d = it->get("data");
                if (!n || !d || d->kind != RonValue::Tagged) continue;
                SceneMaterial m;
                m.name = n->s;
                auto num = [&](const char* f, double dflt) { return d->get(f) ? d->get(f)->num() : dflt; };
                auto flag = [&](const char* f) { return d->get(f) && d->get(f)->b; };
                if (d->s == "Simple") {
                    m.type = SceneMaterial::Simple;
                    vec3(d->get("color"), m.color);
                    m.normal_coef = num("normal_coef", 0);
                    m.grid = flag("grid");
                    m.grid_scale = num("grid_scale", 0);
                    m.grid_coef = num("grid_coef", 0);
                    m.grid2 = flag("grid2");
                    m.grid3 = flag("grid3");
                } else if (d->s == "Reflect") {
                    m.type = SceneMaterial::Reflect;
                    vec3(d->get("add_to_color"), m.color);
                } else if (d->s == "Refract") {
                    m.type = SceneMaterial::Refract;
                    vec3(d->get("add_to_color"), m.color);
                    m.refractive_index = num("refractive_index", 1);
                } else {
                    m.type = SceneMaterial::Complex;
                    m.code = code_of(d->get("code"));
                }
                sc.materials.push_back(m);
            }
        if (const RonValue* ims = storage(root, "intersection_materials", false))
            for (auto& it : ims->items)
                if (it->get("name")) sc.intersection_materials.emplace_back(it->get("name")->s, code_of(it->get("data")));
        if (const RonValue* libs = storage(root, "library", true))
            for (auto& it : libs->items)
                if (it->get("name")) sc.library.emplace_back(it->get("name")->s, code_of(it->get("data")));
        // videos (scene_serialized.rs:1200-1210: after the library, before the stages; an inline uniform gets its id here)
        if (const RonValue* vs = storage(root, "videos", false))   // #[serde(default)], scene_serialized.rs:625-626
            for (auto& it : vs->items) {
                const RonValue* n = it->get("name");
                const RonValue* d = it->get("data");
                if (!n || !d) continue;
                Scene::Video v;
                v.name = n->s;
                if (const RonValue* pth = d->get("path")) v.path = pth->s;
                v.uniform = uniform_ref(d->get("uniform"));
                sc.videos.push_back(v);
            }

        // animation stages (scene_serialized.rs:1286-1360): inline elements are inserted now
        auto stage_anim = [&](const RonValue* v, bool is_matrix) {
            StageAnim a;
            if (v && v->kind == RonValue::Tagged && (v->s == "Changed" || v->s == "ChangedAndToUser")) {
                a.from_dev = false;
                a.element = is_matrix ? matrix_ref(v->at(0)) : uniform_ref(v->at(0));
            }
            return a;
        };
        if (const RonValue* sts = storage(root, "animation_stages", false))
            for (auto& it : sts->items) {
                const RonValue* n = it->get("name");
                const RonValue* d = it->get("data");
                if (!n || !d) continue;
                Stage st;
                if (const RonValue* us = d->get("uniforms"))
                    for (auto& kv : us->map) {
                        auto f = sc.uniform_by_name.find(kv.first->s);
                        if (f != sc.uniform_by_name.end()) st.uniforms.emplace_back(f->second, stage_anim(kv.second.get(), false));
                    }
                if (const RonValue* ms2 = d->get("matrices"))
                    for (auto& kv : ms2->map) {
                        auto f = sc.matrix_by_name.find(kv.first->s);
                        if (f != sc.matrix_by_name.end()) st.matrices.emplace_back(f->second, stage_anim(kv.second.get(), true));
                    }
                sc.stages[n->s] = st;
            }
        // dev stage (scene_serialized.rs:1362-1374)
        if (const RonValue* dev = root.get("dev_stage")) {
            if (const RonValue* us = dev->get("uniforms"))
                for (auto& kv : us->map) {
                    auto f = sc.uniform_by_name.find(kv.first->s);
                    if (f == sc.uniform_by_name.end()) continue;
                    int tmp = add_uniform(*kv.second, "");        // parse through the common path, then detach
                    sc.dev_uniforms[f->second] = sc.uniforms[size_t(tmp)];
                }
            if (const RonValue* ms3 = dev->get("matrices"))
                for (auto& kv : ms3->map) {
                    auto f = sc.matrix_by_name.find(kv.first->s);
                    if (f != sc.matrix_by_name.end()) sc.dev_matrices[f->second] = matrix_from(*kv.second);
                }
        }
        load_cameras_and_animations(root);
    }

    Cam cam_from(const RonValue& d) {  // cam_from_ser (scene_serialized.rs:476-496)
        Cam c;
        if (const RonValue* la = d.get("look_at")) {
            if (la->is_tag("MatrixCenter")) {
                c.look_at_matrix = true;
                c.matrix = matrix_ref(la->at(0));
            } else if (la->is_tag("Coordinate")) {
                const RonValue* v = unwrap1(la->at(0));
                if (la->items.size() == 1 && la->items[0]->kind == RonValue::List && la->items[0]->items.size() == 3) v = la->items[0].get();
                if (v && v->kind == RonValue::List && v->items.size() == 3)
                    for (int k = 0; k < 3; k++) c.coord[k] = v->items[size_t(k)]->num();
            }
        }
        auto num = [&](const char* f, double dflt) { const RonValue* v = d.get(f); return v && v->is_num() ? v->num() : dflt; };
        auto flag = [&](const char* f) { const RonValue* v = d.get(f); return v && v->kind == RonValue::Bool && v->b; };
        c.alpha = num("alpha", 0.0);
        c.beta = num("beta", 0.0);
        c.r = num("r", 3.5);
        c.in_subspace = flag("in_subspace");
        c.free_movement = flag("free_movement");
        if (const RonValue* m = d.get("matrix"))
            if (m->kind == RonValue::List && m->items.size() == 16)
                for (int k = 0; k < 16; k++) c.teleport[size_t(k)] = m->items[size_t(k)]->num();
        return c;
    }

    int cam_ref(const RonValue* ref) {
        if (!ref || ref->kind == RonValue::Null) return -1;
        if (ref->is_tag("Named")) {
            auto it = sc.camera_by_name.find(ref->at(0) ? ref->at(0)->s : "");
            return it == sc.camera_by_name.end() ? -1 : it->second;
        }
        if (ref->is_tag("Inline") && ref->at(0)) {
            sc.cameras.push_back(cam_from(*ref->at(0)));
            return int(sc.cameras.size()) - 1;
        }
        return -1;
    }

    StageRef stage_ref(const RonValue* v) {  // scene_serialized.rs:1386-1397
        StageRef r;
        if (!v || v->kind != RonValue::Tagged) return r;
        const std::string n = v->at(0) ? v->at(0)->s : "";
        if (v->s == "Animation" && sc.stages.count(n)) {
            r.kind = StageRef::Animation;
            r.name = n;
        } else if (v->s == "RealAnimation" && sc.animation_by_name.count(n)) {
            r.kind = StageRef::Real;
            r.index = sc.animation_by_name[n];
        }
        return r;
    }

    void load_cameras_and_animations(const RonValue& root) {
        if (const RonValue* cams = storage(root, "cameras", false))
            for (auto& it : cams->items) {
                const RonValue* n = it->get("name");
                const RonValue* d = it->get("data");
                if (!n || !d) continue;
                sc.camera_by_name.emplace(n->s, int(sc.cameras.size()));
                sc.cameras.push_back(cam_from(*d));
            }
        if (const RonValue* sts = storage(root, "animation_stages", false))
            for (auto& it : sts->items) {
                const RonValue* n = it->get("name");
                const RonValue* d = it->get("data");
                if (!n || !d) continue;
                sc.stage_cam[n->s] = cam_ref(d->get("set_cam"));   // None and Some(None) both select the original camera
            }
        const RonValue* anims = storage(root, "animations", false);
        if (anims) {
            for (auto& it : anims->items)
                if (const RonValue* n = it->get("name")) {
                    sc.animation_by_name.emplace(n->s, int(sc.animations.size()));
                    RealAnimation a;
                    a.name = n->s;
                    sc.animations.push_back(a);
                }
            size_t idx = 0;
            for (auto& it : anims->items) {
                if (!it->get("name")) continue;
                RealAnimation& a = sc.animations[idx++];
                const RonValue* d = it->get("data");
                if (!d) continue;
                if (const RonValue* v = d->get("duration")) a.duration = v->num();
                a.stage = stage_ref(d->get("animation_stage"));
                auto parts = [&](const char* field, bool is_matrix, std::vector<std::pair<int, int>>& out) {
                    const RonValue* m = unwrap1(d->get(field));
                    if (!m) return;
                    for (auto& kv : m->map) {
                        auto& by = is_matrix ? sc.matrix_by_name : sc.uniform_by_name;
                        auto f = by.find(kv.first->s);
                        if (f == by.end()) continue;
                        if (kv.second->is_tag("Changed")) {
                            const int e = is_matrix ? matrix_ref(kv.second->at(0)) : uniform_ref(kv.second->at(0));
                            if (e >= 0) out.emplace_back(f->second, e);
                        }
                    }
                };
                parts("uniforms", false, a.uniforms);
                parts("matrices", true, a.matrices);
                auto flag = [&](const char* f) { const RonValue* v = d->get(f); return v && v->kind == RonValue::Bool && v->b; };
                auto opt_flag = [&](const char* f) { const RonValue* v = d->get(f); return v && v->kind == RonValue::Bool ? (v->b ? 1 : 0) : -1; };
                auto anim_by = [&](const char* f) {
                    const RonValue* v = d->get(f);
                    if (!v || v->kind != RonValue::String) return -1;
                    auto it2 = sc.animation_by_name.find(v->s);
                    return it2 == sc.animation_by_name.end() ? -1 : it2->second;
                };
                a.use_prev_cam = flag("use_prev_cam");
                a.use_start_cam_as_end = flag("use_start_cam_as_end");
                a.cam_start = cam_ref(d->get("cam_start"));
                a.cam_end = cam_ref(d->get("cam_end"));
                a.use_any_cam_as_start = opt_flag("use_any_cam_as_start");
                a.use_any_cam_as_end = opt_flag("use_any_cam_as_end");
                a.cam_any_start = anim_by("cam_any_start");
                a.cam_any_end = anim_by("cam_any_end");
                if (const RonValue* e = d->get("cam_easing"))
                    if (e->kind == RonValue::Tagged) {
                        static const char* names[] = {"Linear", "In", "Out", "InOut", "InOutFast", "ElasticOut"};
                        for (int k = 0; k < 6; k++) if (e->s == names[k]) a.cam_easing = k;
                    }
                if (const RonValue* u = d->get("cam_easing_uniform"))
                    if (u->kind != RonValue::Null) {
                        const int id = uniform_ref(u);
                        if (id >= 0) { a.has_easing_uniform = true; a.easing_uniform = id; }   // scene_serialized.rs:1468-1471
                    }
            }
        }
        sc.current_stage = stage_ref(root.get("current_stage"));
    }
};

// easing.rs:6-44
double easing_in(double t) { return 1.0 - std::cos(t * M_PI * 0.5); }
double easing_out(double t) { return 1.0 - easing_in(1.0 - t); }
double easing_in_out(double t) { return (1.0 - std::cos(t * M_PI)) * 0.5; }
double easing_in_out_fast(double t) { return easing_in_out(easing_in_out(t)); }
double easing_plus_minus(double t) {
    t *= 2.0 * M_PI;
    const double t2 = 2.0 * t;
    return std::sin(t) * (3.0 - std::cos(t) - std::cos(t2) - std::cos(t) * std::cos(t2)) / 4.0;
}
double easing_elastic_out(double x) {
    const double c4 = (2.0 * M_PI) / 3.0;
    if (x == 0.0) return 0.0;
    if (x == 1.0) return 1.0;
    return std::pow(2.0, -10.0 * x) * std::sin((x * 10.0 - 0.75) * c4) + 1.0;
}

}  // namespace

double ease(int easing, double t) {  // Easing::ease (easing.rs:87-99)
    switch (easing) {
        case 1: return easing_in(t);
        case 2: return easing_out(t);
        case 3: return easing_in_out(t);
        case 4: return easing_in_out_fast(t);
        case 5: return easing_elastic_out(t);
        default: return t;
    }
}

void Scene::init_dev_stage() {
    for (auto& kv : dev_uniforms) uniforms[size_t(kv.first)] = kv.second;
    for (auto& kv : dev_matrices) matrices[size_t(kv.first)] = kv.second;
}

bool Scene::load(const RonValue& root) {
    Loader l(*this);
    l.load(root);
    return l.ok;
}

bool Scene::init_stage(const std::string& name) {
    auto it = stages.find(name);
    if (it == stages.end()) {
        error = "scene has no stage named `" + name + "`";
        return false;
    }
    for (auto& ua : it->second.uniforms) {
        if (!ua.second.from_dev) {
            if (ua.second.element >= 0) uniforms[size_t(ua.first)] = uniforms[size_t(ua.second.element)];  // storage.set_id
        } else {
            auto d = dev_uniforms.find(ua.first);
            if (d != dev_uniforms.end()) uniforms[size_t(ua.first)] = d->second;
        }
    }
    for (auto& ma : it->second.matrices) {
        if (!ma.second.from_dev) {
            if (ma.second.element >= 0) matrices[size_t(ma.first)] = matrices[size_t(ma.second.element)];
        } else {
            auto d = dev_matrices.find(ma.first);
            if (d != dev_matrices.end()) matrices[size_t(ma.first)] = d->second;
        }
    }
    return true;
}

std::string Scene::matrix_uniform_stem(int id) const {
    return matrix_names[id].empty() ? "id" + std::to_string(id) : matrix_names[id];
}

bool Scene::video_frame(int k, size_t frame_count, size_t& index) {
    if (k < 0 || k >= int(videos.size()) || videos[size_t(k)].uniform < 0 || frame_count == 0) return false;
    int kind = 0;
    double v = 0.0;
    std::vector<int> visited;
    if (!get_uniform(videos[size_t(k)].uniform, kind, v, visited)) return false;
    if (v == v) v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);          // f64::clamp keeps NaN
    const double last = double(frame_count - 1);
    double x = std::round(last * v);                              // Rust's round: half away from zero, like std::round
    if (!(x == x)) x = 0.0;                                        // `NaN as usize` is 0
    x = x < 0.0 ? 0.0 : (x > last ? last : x);
    index = size_t(x);
    return true;
}

bool Scene::get_uniform(int id, int& kind, double& value, std::vector<int>& visited) {
    if (id < 0 || id >= int(uniforms.size())) return false;
    if (std::find(visited.begin(), visited.end(), id) != visited.end()) return false;  // storage2.rs:158-160
    Uniform& u = uniforms[id];
    switch (u.kind) {
        case Uniform::Bool: kind = 0; value = u.b ? 1.0 : 0.0; return true;
        case Uniform::Int: kind = 1; value = double(u.i); return true;
        case Uniform::Float: case Uniform::Angle: case Uniform::Progress: kind = 2; value = u.f; return true;
        case Uniform::Trefoil: kind = 3; value = -1.0; return true;   // uniform.rs:300-311: -1.0 wherever a number is wanted
        case Uniform::Unsupported: return false;
        default: break;
    }
    if (!u.parsed && !u.parse_failed) {
        std::string err;
        u.parsed = formula_parse(u.text, err);
        u.parse_failed = !u.parsed;
    }
    if (!u.parsed) return false;
    visited.push_back(id);
    // the reference's `cb` closure, uniform.rs:1014-1124
    FormulaNamespace ns = [&](const std::string& name, const std::vector<double>& a, double& out) -> bool {
        auto arg = [&](size_t k) -> double {
            if (k >= a.size()) throw std::runtime_error(name + ": missing argument");
            return a[k];
        };
        auto truthy = [](double x) { return std::fabs(x - 1.0) < 1e-6; };
        if (name == "if") out = truthy(arg(0)) ? arg(1) : arg(2);
        else if (name == "and") out = (truthy(arg(0)) && truthy(arg(1))) ? 1.0 : 0.0;
        else if (name == "or") out = (truthy(arg(0)) || truthy(arg(1))) ? 1.0 : 0.0;
        else if (name == "not") out = truthy(arg(0)) ? 0.0 : 1.0;
        else if (name == "deg2rad") out = arg(0) / 180.0 * M_PI;
        else if (name == "rad2deg") out = arg(0) * 180.0 / M_PI;
        else if (name == "switch") {
            double idx = arg(0);
            size_t k = (std::isnan(idx) || idx < 0) ? 0 : size_t(idx);
            out = arg(k);
        } else if (name == "on") {
            double v = arg(0), lo = arg(1), hi = arg(2);
            out = v < lo ? 0.0 : (v > hi ? 1.0 : (v - lo) / (hi - lo));
        } else if (name == "inv") out = 1.0 - arg(0);
        else if (name == "sqrt") out = std::sqrt(arg(0));
        else if (name == "atan2") out = std::atan2(arg(0), arg(1));
        else if (name == "time") out = time;
        else if (name == "total_time") out = total_time;
        else if (name == "easing_linear") out = arg(0);
        else if (name == "easing_in") out = easing_in(arg(0));
        else if (name == "easing_out") out = easing_out(arg(0));
        else if (name == "easing_in_out") out = easing_in_out(arg(0));
        else if (name == "easing_in_out_fast") out = easing_in_out_fast(arg(0));
        else if (name == "easing_plus_minus") out = easing_plus_minus(arg(0));
        else if (name == "easing_elastic_out") out = easing_elastic_out(arg(0));
        else if (name == "bump") {
            double x = (arg(0) - arg(1)) / arg(2);
            out = std::fabs(x) < 1.0 ? 0.5 * (1.0 + std::cos(M_PI * x)) : 0.0;
        } else if (name == "later_start") {
            double t = arg(0), tm = 1.0 - arg(1);
            out = std::fmax(0.0, t / tm - (1.0 - tm) / tm);
        } else if (name == "early_finish") out = std::fmin(1.0, arg(0) / arg(1));
        else if (name == "lerp") out = (1.0 - arg(2)) * arg(0) + arg(2) * arg(1);  // emath::lerp
        else {
            auto it = uniform_by_name.find(name);
            if (it == uniform_by_name.end()) return false;
            int k;
            double v;
            if (!get_uniform(it->second, k, v, visited)) return false;
            out = v;
        }
        return true;
    };
    double v = 0.0;
    std::string err;
    bool ok = formula_eval(u.parsed, ns, v, err);
    visited.pop_back();
    if (!ok) return false;
    if (u.kind == Uniform::Formula) {
        kind = 2;
        value = v;
    } else {  // `as i32`: saturating, NaN -> 0
        kind = 1;
        value = std::isnan(v) ? 0.0 : std::trunc(std::fmax(-2147483648.0, std::fmin(2147483647.0, v)));
    }
    return true;
}

bool Scene::get_param(const Param& p, double& out) {
    if (!p.is_uniform) {
        out = p.value;
        return true;
    }
    int k;
    std::vector<int> visited;
    return get_uniform(p.uniform, k, out, visited);
}

bool Scene::get_matrix(int id, Mat4& out, std::vector<int>& visited) {
    if (id < 0 || id >= int(matrices.size())) return false;
    if (std::find(visited.begin(), visited.end(), id) != visited.end()) return false;
    visited.push_back(id);
    const Matrix& m = matrices[id];
    bool ok = true;
    Mat4 A, B, Cc;
    switch (m.kind) {
        case Matrix::Mul:  // what * to (matrix.rs:517-521)
            ok = get_matrix(m.a, A, visited) && get_matrix(m.b, B, visited);
            if (ok) out = mat_mul(B, A);
            break;
        case Matrix::Teleport:  // second * first^-1 * what (:522-531)
            ok = get_matrix(m.a, A, visited) && get_matrix(m.b, B, visited) && get_matrix(m.c, Cc, visited);
            if (ok) out = mat_mul(mat_mul(B, mat_inverse(A)), Cc);
            break;
        case Matrix::Simple: {
            const double s[3] = {m.scale * (m.mirror[0] ? -1.0 : 1.0), m.scale * (m.mirror[1] ? -1.0 : 1.0), m.scale * (m.mirror[2] ? -1.0 : 1.0)};
            out = mat_srt(s, m.rotate, m.offset);
            break;
        }
        case Matrix::Parametrized: {
            double v[10];
            for (int k : {9, 6, 7, 8, 3, 4, 5, 0, 1, 2})  // evaluation order of matrix.rs:555-569
                if (!get_param(m.p[k], v[k])) { ok = false; break; }
            if (ok) {
                const double s[3] = {v[9] * (1.0 - 2.0 * v[6]), v[9] * (1.0 - 2.0 * v[7]), v[9] * (1.0 - 2.0 * v[8])};
                out = mat_srt(s, v + 3, v + 0);
            }
            break;
        }
        case Matrix::Exact: {
            out = Mat4{};
            for (int c = 0; c < 4 && ok; c++) {
                for (int r = 0; r < 3 && ok; r++) ok = get_param(m.p[3 * c + r], out[4 * c + r]);
                out[4 * c + 3] = c == 3 ? 1.0 : 0.0;
            }
            break;
        }
        case Matrix::ExactFull:
            for (int k = 0; k < 16 && ok; k++) ok = get_param(m.p[k], out[k]);
            break;
        case Matrix::If: {
            double c;
            ok = get_param(m.p[0], c);
            if (ok) ok = c > 0.5 ? get_matrix(m.a, out, visited) : get_matrix(m.b, out, visited);
            break;
        }
        case Matrix::Inv:
            ok = get_matrix(m.a, A, visited);
            if (ok) out = mat_inverse(A);
            break;
        case Matrix::Camera:
            out = camera_matrix_for_formulas;
            break;
        case Matrix::Lerp: {
            double t;
            ok = get_param(m.p[0], t) && get_matrix(m.a, A, visited) && get_matrix(m.b, B, visited);
            if (ok) out = mat_lerp(A, B, t);
            break;
        }
        case Matrix::Sqrt:
            // matrix.rs:606-613: the matrix is missing when the minimisation does not reach cost < 1e-4
            ok = get_matrix(m.a, A, visited) && mat_sqrt(A, out);
            break;
    }
    visited.pop_back();
    return ok;
}

uint64_t Scene::input_fingerprint() const {
    uint64_t h = 0x9e3779b97f4a7c15ull;
    auto mix_w = [&](uint64_t w) {                    // one 64-bit word per step (splitmix-style finaliser)
        h ^= w + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        h *= 0xbf58476d1ce4e5b9ull;
        h ^= h >> 31;
    };
    auto mix = [&](const void* p, size_t n) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        size_t k = 0;
        for (; k + 8 <= n; k += 8) { uint64_t w; std::memcpy(&w, b + k, 8); mix_w(w); }
        if (k < n) { uint64_t w = 0; std::memcpy(&w, b + k, n - k); mix_w(w ^ (uint64_t(n - k) << 56)); }
    };
    auto mix_d = [&](double v) { uint64_t w; std::memcpy(&w, &v, 8); mix_w(w); };
    auto mix_i = [&](long long v) { mix_w(uint64_t(v)); };
    mix_d(time); mix_d(total_time);
    bool uses_camera = false;
    mix_i((long long)uniforms.size());
    for (const Uniform& u : uniforms) {
        mix_i((long long)u.kind | ((long long)u.b << 8) | ((long long)(unsigned)u.i << 16)); mix_d(u.f);
        if (!u.text.empty()) mix(u.text.data(), u.text.size());
        if (u.kind == Uniform::Trefoil) mix(u.trefoil, sizeof u.trefoil);
    }
    mix_i((long long)matrices.size());
    for (const Matrix& m : matrices) {
        mix_i((long long)m.kind | ((long long)(m.a + 1) << 8) | ((long long)(m.b + 1) << 24) | ((long long)(m.c + 1) << 40));
        switch (m.kind) {
            case Matrix::Simple:
                mix(m.offset, sizeof m.offset); mix(m.rotate, sizeof m.rotate); mix_d(m.scale);
                mix_i(m.mirror[0] | (m.mirror[1] << 1) | (m.mirror[2] << 2));
                break;
            case Matrix::Camera: uses_camera = true; break;
            case Matrix::Mul: case Matrix::Teleport: case Matrix::Inv: case Matrix::Sqrt: break;
            default:
                for (const Param& q : m.p) { mix_i((long long)q.is_uniform | ((long long)(q.uniform + 1) << 8)); mix_d(q.value); }
        }
    }
    if (uses_camera) mix(camera_matrix_for_formulas.data(), sizeof(double) * 16);   // only the `Camera` matrix kind reads it
    return h;
}

bool Scene::uniform_table(std::vector<TableEntry>& out) {
    out.clear();
    std::map<std::string, size_t> index;
    auto put_mat = [&](const std::string& name, const Mat4& m) {
        auto it = index.find(name);
        if (it != index.end()) {
            out[it->second].m = m;
            return;
        }
        index[name] = out.size();
        TableEntry e;
        e.name = name;
        e.type = PE_UNIFORM_MAT4;
        e.m = m;
        out.push_back(e);
    };
    // passed_matrices: object-referenced, then every named matrix (scene.rs:551-592)
    std::vector<int> ids;
    for (const SceneObject& o : objects) {
        if (o.cls == SceneObject::DebugMatrix) {
            if (o.matrix_a >= 0) ids.push_back(o.matrix_a);
        } else if (o.portal) {
            if (o.matrix_a >= 0 && o.matrix_b >= 0) { ids.push_back(o.matrix_a); ids.push_back(o.matrix_b); }
        } else if (o.matrix_a >= 0) ids.push_back(o.matrix_a);
    }
    for (size_t k = 0; k < matrices.size(); k++)
        if (!matrix_names[k].empty()) ids.push_back(int(k));
    for (int id : ids) {
        Mat4 m;
        std::vector<int> visited;
        if (!get_matrix(id, m, visited)) continue;  // "matrix can't be getted": logged and skipped (scene.rs:589-591)
        const std::string stem = matrix_uniform_stem(id);
        put_mat(stem + "_mat", m);
        put_mat(stem + "_mat_inv", mat_inverse(m));
    }
    // teleport matrices (scene.rs:594-635)
    for (const SceneObject& o : objects) {
        if (o.cls == SceneObject::DebugMatrix || !o.portal || o.matrix_a < 0 || o.matrix_b < 0) continue;
        Mat4 ma, mb;
        std::vector<int> v1, v2;
        if (!get_matrix(o.matrix_a, ma, v1) || !get_matrix(o.matrix_b, mb, v2)) continue;
        const std::string na = matrix_uniform_stem(o.matrix_a), nb = matrix_uniform_stem(o.matrix_b);
        put_mat(na + "_to_" + nb + "_mat_teleport", mat_mul(mb, mat_inverse(ma)));
        if (na != nb) put_mat(nb + "_to_" + na + "_mat_teleport", mat_mul(ma, mat_inverse(mb)));
    }
    // scalars (scene.rs:637-656)
    for (size_t k = 0; k < uniforms.size(); k++) {
        if (uniform_names[k].empty()) continue;
        int kind;
        double v;
        std::vector<int> visited;
        if (!get_uniform(int(k), kind, v, visited)) continue;
        if (kind == 3) {  // TrefoilSpecial: 18 packed ints `ts_<i>_<name>_u` (scene.rs:488-492, 644-650)
            for (int i = 0; i < 18; i++) {
                TableEntry t;
                t.name = "ts_" + std::to_string(i) + "_" + uniform_names[k] + "_u";
                t.type = PE_UNIFORM_INT;
                t.i = uniforms[k].trefoil[i];
                out.push_back(t);
            }
            continue;
        }
        TableEntry e;
        e.name = uniform_names[k] + "_u";
        if (kind == 2) { e.type = PE_UNIFORM_FLOAT; e.f = v; }
        else { e.type = PE_UNIFORM_INT; e.i = int(v); }
        out.push_back(e);
    }
    return true;
}

}  // namespace ph
