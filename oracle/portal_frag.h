// CPU restatement of the reference's fragment-shader driver (the ray loop).
//
// ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// parity unpinned: the reference holds no tests, golden vectors or fixtures for this path (oracle/README.md).
// Follows /root/reference/src/frag.glsl; each function cites its lines.  Included by
// the generated per-scene translation unit AFTER scene_intersect(), material_process()
// and scene_intersect_material_process() have been emitted (frag.glsl:19-59 with the
// generator's sections, /root/reference/src/gui/scene.rs:885-1035, filled in).
//
// Variants restated: mono camera and side-by-side stereo; perspective, panini, 360 and VR180
// projections; AA loop; darken-by-distance; depth-map colouring.  Anaglyph is compiled out of the
// reference's native build (disable_anaglyph = true, main.rs:938) and the external-ray probe is a
// SURVEY.md §8(f3) "next" row; neither is restated here.
#pragma once

namespace pe_oracle {

// frag.glsl:74-78
struct RayTraceResult {
    vec3 color;
    real depth;
    bool has_depth;
};

// frag.glsl:80-84
static inline real normalize_depth_value(real depth) {
    real depth_min = min(_depth_map_min, _depth_map_max);
    real depth_max = max(_depth_map_min, _depth_map_max);
    return clamp((depth - depth_min) / max(PE_L(1e-6), depth_max - depth_min), PE_L(0.0), PE_L(1.0));
}

// frag.glsl:86-99
static inline vec3 depth_gradient_inferno(real t) {
    vec3 c0 = sqrvec(vec3(PE_L(0.001462), PE_L(0.000466), PE_L(0.013866)));
    vec3 c1 = sqrvec(vec3(PE_L(0.258234), PE_L(0.038571), PE_L(0.406485)));
    vec3 c2 = sqrvec(vec3(PE_L(0.578304), PE_L(0.148039), PE_L(0.404411)));
    vec3 c3 = sqrvec(vec3(PE_L(0.865006), PE_L(0.316822), PE_L(0.226055)));
    vec3 c4 = sqrvec(vec3(PE_L(0.987622), PE_L(0.645320), PE_L(0.039886)));
    vec3 c5 = sqrvec(vec3(PE_L(0.988362), PE_L(0.998364), PE_L(0.644924)));

    if (t < PE_L(0.2)) return mix(c0, c1, t / PE_L(0.2));
    if (t < PE_L(0.4)) return mix(c1, c2, (t - PE_L(0.2)) / PE_L(0.2));
    if (t < PE_L(0.6)) return mix(c2, c3, (t - PE_L(0.4)) / PE_L(0.2));
    if (t < PE_L(0.8)) return mix(c3, c4, (t - PE_L(0.6)) / PE_L(0.2));
    return mix(c4, c5, (t - PE_L(0.8)) / PE_L(0.2));
}

// frag.glsl:101-104
static inline vec3 sample_depth_gradient(real depth) {
    real t = PE_L(1.0) - normalize_depth_value(depth);
    return depth_gradient_inferno(t);
}

// frag.glsl:106-159 (for_prefer_variable = true: the !FOR_VARIABLE! loop header, scene.rs:1076-1077)
static inline RayTraceResult ray_tracing(Ray r, real camera_scale, int* bounces_out) {
    // skybox_processing section, scene.rs:1052-1063 (no skybox in the config scenes)
    vec3 not_found_color = PE_NOT_FOUND_COLOR(r);

    vec3 current_color = vec3(real(1));
    real all_t = real(0);
    for (int j = 0; j < _ray_tracing_depth; j++) {
        if (bounces_out) *bounces_out = j + 1;
        SceneIntersection i = scene_intersect(r);
        SceneIntersectionWithMaterial i2 = scene_intersect_material_process(r);

        MaterialProcessing m = material_empty();  // GLSL leaves `m` undefined here (frag.glsl:116)
        if (nearer(i.hit, i2.scene.hit)) {
            r.o += r.d * i2.scene.hit.t;
            all_t += i2.scene.hit.t * r.tmul;
            if (i2.scene.material == CUSTOM_MATERIAL) {
                m = i2.material;
            } else {
                m = material_process(r, i2.scene);
            }
        } else if (i.hit.hit) {
            r.o += r.d * i.hit.t;
            all_t += i.hit.t * r.tmul;
            m = material_process(r, i);
        }

        if (i.hit.hit || i2.scene.hit.hit) {
            current_color *= m.mul_to_color;
            if (m.is_final) {
                real depth = all_t / max(camera_scale, PE_L(1e-6));
                if (all_t > _t_start * camera_scale && _darken_by_distance == 1) {
                    if (all_t > _t_end * camera_scale) all_t = _t_end * camera_scale;
                    real gray_t = (all_t - _t_start * camera_scale) / (_t_end - _t_start) / camera_scale;
                    return RayTraceResult{
                        color(real(0), real(0), real(0)) * sqr(sqr(gray_t)) + current_color * sqr(sqr(PE_L(1.0) - gray_t)), depth,
                        true};
                } else {
                    return RayTraceResult{current_color, depth, true};
                }
            } else {
                r = m.new_ray;
            }
        } else {
            if (r.in_subspace) {
                return RayTraceResult{color(real(0), real(0), real(0)), PE_L(0.0), false};
            } else {
                return RayTraceResult{current_color * not_found_color, PE_L(0.0), false};
            }
        }
    }
    return RayTraceResult{color(real(0), real(0), real(0)), PE_L(0.0), false};
}

// frag.glsl:199-203
struct ExternalRayTeleportation {
    vec3 pos;
    bool encounter_object;
    bool change_subspace;
    bool have_result;   // not in the GLSL struct: the reference's host infers it from pos != 0 (main.rs:1400-1408)
};

// frag.glsl:209-257: camera teleportation probe -- follow the segment a->b (r.d = b - a, so t in [0,1]
// spans the segment) through at most 10 portals and report where its end point lands.
static inline ExternalRayTeleportation teleport_external_ray(Ray r) {
    r = normalize_ray(r);
    bool have_result = false;
    bool stop_at_object = false;
    real all_t = real(0);
    int max_camera_teleports = 10;
    for (int j = 0; j < max_camera_teleports; j++) {
        SceneIntersection i = scene_intersect(r);
        SceneIntersectionWithMaterial i2 = scene_intersect_material_process(r);

        bool continue_intersect = false;
        MaterialProcessing m = material_empty();
        if (nearer(i.hit, i2.scene.hit)) {
            if (i2.scene.hit.t * r.tmul + all_t < PE_L(1.0)) {
                r.o += r.d * i2.scene.hit.t;
                all_t += i2.scene.hit.t * r.tmul;
                if (i2.scene.material == CUSTOM_MATERIAL) {
                    m = i2.material;
                } else {
                    m = material_process(r, i2.scene);
                }
                continue_intersect = !m.is_final;
                stop_at_object = stop_at_object || m.is_final;
            }
        } else if (i.hit.hit) {
            if (i.hit.t * r.tmul + all_t < PE_L(1.0)) {
                r.o += r.d * i.hit.t;
                all_t += i.hit.t * r.tmul;
                m = material_process(r, i);
                continue_intersect = !m.is_final;
                stop_at_object = stop_at_object || m.is_final;
            }
        }
        if (continue_intersect) {
            r = m.new_ray;
            have_result = true;
        } else {
            break;
        }
    }
    if (have_result) {
        r.o += r.d * (PE_L(1.0) - all_t) / r.tmul;
        return ExternalRayTeleportation{vec3(r.o), stop_at_object, int(r.in_subspace) != _camera_in_subspace, true};
    } else {
        return ExternalRayTeleportation{vec3(real(0)), stop_at_object, int(r.in_subspace) != _camera_in_subspace, false};
    }
}

// frag.glsl:297-301
static inline real Pow2(real x) { return x * x; }

// frag.glsl:305-342
static inline vec3 PaniniProjection(vec2 tc, real fov, real d) {
    const real Pi = PE_L(3.14159265359);
    const real Pi05 = Pi * PE_L(0.5);
    real d2 = d * d;
    {
        real fo = Pi05 - fov * PE_L(0.5);
        real f = cos(fo) / sin(fo);
        real f2 = f * f;
        real b = (sqrt(max(PE_L(0.0), Pow2(d + d2) * (f2 + f2 * f2))) - (d * f + f)) / (d2 + d2 * f2 - PE_L(1.0));
        tc *= b;
    }
    real h = tc.x;
    real v = tc.y;
    real h2 = h * h;
    real k = h2 / Pow2(d + PE_L(1.0));
    real k2 = k * k;
    real discr = max(PE_L(0.0), k2 * d2 - (k + PE_L(1.0)) * (k * d2 - PE_L(1.0)));
    real cosPhi = (-k * d + sqrt(discr)) / (k + PE_L(1.0));
    real S = (d + PE_L(1.0)) / (d + cosPhi);
    real tanTheta = v / S;
    real sinPhi = sqrt(max(PE_L(0.0), PE_L(1.0) - Pow2(cosPhi)));
    if (tc.x < PE_L(0.0)) sinPhi *= PE_L(-1.0);
    real s = inversesqrt(PE_L(1.0) + Pow2(tanTheta));
    return vec3(sinPhi, tanTheta, cosPhi) * s;
}

// frag.glsl:408-464: panini :411-412, 360 :413-437, VR180 :438-448, perspective :449-452
static inline vec3 get_color2(vec2 image_position, const mat4& camera_matrix, bool in_subspace, real camera_scale,
                              vec2 resolution, int* bounces_out) {
    const real Pi = PE_L(3.14159265359);
    const real Pi05 = Pi * PE_L(0.5);
    vec4 o = camera_matrix * vec4(real(0), real(0), real(0), real(1));
    vec4 d;
    if (_use_panini_projection == 1) {
        d = normalize(camera_matrix * vec4(PaniniProjection(vec2(image_position.x, image_position.y), _view_angle, _panini_param), real(0)));
    } else if (_use_360_camera == 1) {
        real coef = min(resolution.x, resolution.y);
        real ax = resolution.x / coef;
        real ay = resolution.y / coef;
        real rx;
        real ry;
        if (ax >= PE_L(2.0) * ay) {
            ry = ay;
            rx = PE_L(2.0) * ay;
        } else {
            rx = ax;
            ry = ax / PE_L(2.0);
        }
        if (abs(image_position.x) > rx || abs(image_position.y) > ry) {
            return vec3(PE_L(0.0));
        }
        real yaw = (image_position.x / rx) * Pi;
        real pitch = (image_position.y / ry) * Pi05;
        vec3 dir_local = vec3(sin(yaw) * cos(pitch), sin(pitch), cos(yaw) * cos(pitch));
        d = normalize(camera_matrix * vec4(dir_local, real(0)));
    } else if (_use_180_camera == 1) {
        if (abs(image_position.x) > PE_L(1.0) || abs(image_position.y) > PE_L(1.0)) {
            return vec3(PE_L(0.0));
        }
        real yaw = image_position.x * Pi05;
        real pitch = image_position.y * Pi05;
        vec3 dir_local = vec3(sin(yaw) * cos(pitch), sin(pitch), cos(yaw) * cos(pitch));
        d = normalize(camera_matrix * vec4(dir_local, real(0)));
    } else {
        real h = tan(_view_angle / PE_L(2.));
        d = normalize(camera_matrix * vec4(image_position.x * h, image_position.y * h, PE_L(1.0), real(0)));
    }

    Ray r = Ray{o, d, PE_L(1.0), in_subspace};
    RayTraceResult trace = ray_tracing(r, camera_scale, bounces_out);
    if (_draw_depth_map == 1) {
        if (trace.has_depth) {
            return sample_depth_gradient(trace.depth);
        } else {
            return vec3(PE_L(0.0));
        }
    }
    return trace.color;
}

// frag.glsl:343-406
static inline vec3 anaglyphCombineLinear(vec3 leftLin, vec3 rightLin, int mode) {
    leftLin = clamp(leftLin, PE_L(0.0), PE_L(1.0));
    rightLin = clamp(rightLin, PE_L(0.0), PE_L(1.0));
    if (mode == 0) {
        const vec3 LUMA = vec3(PE_L(0.299), PE_L(0.587), PE_L(0.114));
        real P = _anaglyph_p;
        real Q = _anaglyph_q;
        real l = dot(leftLin, LUMA);
        real r = dot(rightLin, LUMA);
        real denom = max(PE_L(1e-6), PE_L(1.0) - P * Q);
        real Rout = (l - P * r) / denom;
        real Cout = (r - Q * l) / denom;
        return clamp(vec3(Rout, Cout, Cout), PE_L(0.0), PE_L(1.0));
    } else {
        real P = _anaglyph_p;
        real Q = _anaglyph_q;
        const vec3 LUMA = vec3(PE_L(0.299), PE_L(0.587), PE_L(0.114));
        real l = dot(leftLin, LUMA);
        real r = dot(rightLin, LUMA);
        real denom = max(PE_L(1e-6), PE_L(1.0) - P * Q);
        real Rout = (l - P * r) / denom;
        real Cout = (r - Q * l) / denom;
        real sumGB = rightLin.g + rightLin.b;
        real k = (sumGB > PE_L(1e-6)) ? (PE_L(2.0) * Cout / sumGB) : PE_L(0.0);
        vec3 outLin = vec3(Rout, rightLin.g * k, rightLin.b * k);
        return clamp(outLin, PE_L(0.0), PE_L(1.0));
    }
}

// frag.glsl:466-503.  The native build starts with disable_anaglyph = true (main.rs:938), which drops the
// !ANAGLYPH! lines (scene.rs:1079) until the user enables them (main.rs:1692); with them the shader reads as below.
static inline vec3 get_color(vec2 image_position, int* bounces_out) {
    if (_draw_anaglyph == 1) {
        vec2 full = vec2(_resolution_x, _resolution_y);
        int bl = 0, br = 0;
        vec3 left = get_color2(image_position, _camera_left_eye, _left_eye_in_subspace == 1, _left_eye_scale, full, &bl);
        vec3 right = get_color2(image_position, _camera_right_eye, _right_eye_in_subspace == 1, _right_eye_scale, full, &br);
        if (bounces_out) *bounces_out = bl > br ? bl : br;
        return anaglyphCombineLinear(left, right, _anaglyph_mode);
    }
    mat4 final_matrix = _camera;
    bool final_in_subspace = _camera_in_subspace == 1;
    real final_scale = _camera_scale;
    vec2 _resolution = vec2(_resolution_x, _resolution_y);
    vec2 final_resolution = _resolution;
    if (_draw_side_by_side == 1) {
        real coef = min(_resolution.x, _resolution.y);
        vec2 position = image_position / PE_L(2.) * coef + _resolution / PE_L(2.);
        vec2 resolution = vec2(_resolution.x / PE_L(2.), _resolution.y);
        real coef2 = min(resolution.x, resolution.y);
        if (position.x < resolution.x) {
            image_position = (position - resolution / PE_L(2.)) / coef2 * PE_L(2.);
            final_matrix = _camera_left_eye;
            final_in_subspace = _left_eye_in_subspace == 1;
            final_scale = _left_eye_scale;
            final_resolution = resolution;
        } else {
            image_position = (position - vec2(resolution.x, PE_L(0.)) - resolution / PE_L(2.)) / coef2 * PE_L(2.);
            final_matrix = _camera_right_eye;
            final_in_subspace = _right_eye_in_subspace == 1;
            final_scale = _right_eye_scale;
            final_resolution = resolution;
        }
    }
    return get_color2(image_position, final_matrix, final_in_subspace, final_scale, final_resolution, bounces_out);
}

// frag.glsl:506-513
static inline vec2 quasi_random(int i) {
    real a1 = PE_L(0.7548776662466927600500267982588025643670318456949186300834636687);
    real a2 = PE_L(0.5698402909980532659121818632752155853637566123932930564053138358);
    return vec2(mod(PE_L(0.5) + a1 * real(i), PE_L(1.0)), mod(PE_L(0.5) + a2 * real(i), PE_L(1.0)));
}

// Vertex shader /root/reference/src/gui/scene.rs:1688-1693 (uv_screen) followed by
// frag.glsl:515-526, 550-551 (main, _teleport_external_ray == 0).
// Pixel (px, py) is the fragment whose centre is position.xy = (px + 0.5, py + 0.5);
// row py = 0 is position.y = 0.5 (framebuffer row convention).
static inline vec4 shade_pixel(int px, int py, int* bounces_out) {
    vec2 resolution = vec2(_resolution_x, _resolution_y);
    vec2 position = vec2(real(px) + PE_L(0.5), real(py) + PE_L(0.5));
    real coef = min(resolution.x, resolution.y);
    vec2 uv_screen = (position - resolution / PE_L(2.)) / coef * PE_L(2.);

    vec3 result = vec3(real(0));
    real pixel_size = PE_L(1.) / min(resolution.x, resolution.y);
    int worst = 0;
    for (int a = _aa_start; a < _aa_count + _aa_start; a++) {
        vec2 offset = quasi_random(a);
        int b = 0;
        result += get_color(uv_screen + offset * pixel_size * PE_L(2.), &b);
        if (b > worst) worst = b;
    }
    if (bounces_out) *bounces_out = worst;
    result = sqrt(result / real(_aa_count));
    return vec4(result, PE_L(1.));
}

}  // namespace pe_oracle
