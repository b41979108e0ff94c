#include "ph_anim.h"

#include <cmath>

namespace ph {

// ------------------------------------------------------------------------------------ OrbitCam
void OrbitCam::get_pos_vec(double out[3]) const {
    out[0] = std::sin(beta) * std::cos(alpha) * r;
    out[1] = std::cos(beta) * r;
    out[2] = std::sin(beta) * std::sin(alpha) * r;
}

Mat4 OrbitCam::get_matrix() const { return orbit_camera_matrix(look_at, alpha, beta, r, teleport_matrix, free_movement); }

void OrbitCam::get_cam_pos(double out[3]) const {
    const double origin[4] = {0.0, 0.0, 0.0, 1.0};
    double p[4];
    mat_mul_vec(get_matrix(), origin, p);
    out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
}

CalculatedCam OrbitCam::get_calculated_cam() const {
    CalculatedCam c;
    for (int k = 0; k < 3; k++) c.look_at[k] = look_at[k];
    c.alpha = alpha; c.beta = beta; c.r = r;
    c.in_subspace = in_subspace;
    c.free_movement = free_movement;
    c.matrix = teleport_matrix;
    c.override_matrix = true;
    return c;
}

// -------------------------------------------------------------------------------------- Player
Player::Player(Scene& scene) : sc(scene) {
    // result.cam.set_cam(&result.scene.cam); prev_cam_pos = get_cam_pos()  (main.rs:1057-1058)
    for (int k = 0; k < 3; k++) cam.look_at[k] = sc.look_at[k];
    cam.alpha = sc.alpha; cam.beta = sc.beta; cam.r = sc.r;
    cam.get_cam_pos(cam.prev_cam_pos);
    original_cam = cam.get_calculated_cam();
}

bool Player::cam_get(int cam_id, CalculatedCam& out) {
    if (cam_id < 0 || cam_id >= int(sc.cameras.size())) return false;
    const Cam& c = sc.cameras[size_t(cam_id)];
    if (!c.look_at_matrix) {
        for (int k = 0; k < 3; k++) out.look_at[k] = c.coord[k];
    } else {
        Mat4 m;
        std::vector<int> visited;
        if (!sc.get_matrix(c.matrix, m, visited)) return false;
        // glam 0.13.1 DMat4::project_point3(DVec3::ZERO) + 0.001
        double res[4];
        for (int k = 0; k < 4; k++) res[k] = m[size_t(k)] * 0.0;
        for (int k = 0; k < 4; k++) res[k] = m[size_t(4 + k)] * 0.0 + res[k];
        for (int k = 0; k < 4; k++) res[k] = m[size_t(8 + k)] * 0.0 + res[k];
        for (int k = 0; k < 4; k++) res[k] = m[size_t(12 + k)] + res[k];
        const double rw = 1.0 / res[3];
        for (int k = 0; k < 3; k++) out.look_at[k] = res[k] * rw + 0.001;
    }
    out.alpha = c.alpha; out.beta = c.beta; out.r = c.r;
    out.in_subspace = c.in_subspace;
    out.free_movement = c.free_movement;
    out.matrix = c.teleport;
    out.override_matrix = true;
    return true;
}

int Player::get_start_cam(int idx, int depth) const {  // scene.rs:1291-1317
    if (idx < 0 || idx >= int(sc.animations.size()) || depth > 256) return -1;   // (the reference would overflow its stack on a cycle)
    const RealAnimation& a = sc.animations[size_t(idx)];
    if (a.use_prev_cam) return idx >= 1 ? get_end_cam(idx - 1, depth + 1) : -1;
    if (a.use_any_cam_as_start >= 0) {
        if (a.cam_any_start < 0) return -1;
        return a.use_any_cam_as_start ? get_end_cam(a.cam_any_start, depth + 1) : get_start_cam(a.cam_any_start, depth + 1);
    }
    return a.cam_start;
}

int Player::get_end_cam(int idx, int depth) const {  // scene.rs:1319-1335
    if (idx < 0 || idx >= int(sc.animations.size()) || depth > 256) return -1;
    const RealAnimation& a = sc.animations[size_t(idx)];
    if (a.use_start_cam_as_end) return get_start_cam(idx, depth + 1);
    if (a.use_any_cam_as_end >= 0) {
        if (a.cam_any_end < 0) return -1;
        return a.use_any_cam_as_end ? get_end_cam(a.cam_any_end, depth + 1) : get_start_cam(a.cam_any_end, depth + 1);
    }
    return a.cam_end;
}

double Player::total_animation_duration() const {
    double s = 0.0;
    for (auto& a : sc.animations) s += a.duration;
    return s;
}

bool Player::init_stage(const StageRef& stage) {  // scene.rs:1180-1236
    if (stage.kind == StageRef::Animation) {
        if (!sc.init_stage(stage.name)) { error = sc.error; return false; }
        auto it = sc.stage_cam.find(stage.name);
        current_cam = it == sc.stage_cam.end() ? -1 : it->second;
    } else if (stage.kind == StageRef::Dev) {
        sc.init_dev_stage();
        current_cam = -1;
    } else {
        if (stage.index < 0 || stage.index >= int(sc.animations.size())) { error = "bad animation index"; return false; }
        const RealAnimation a = sc.animations[size_t(stage.index)];
        if (a.stage != stage) {
            if (++depth_guard > 64) { depth_guard = 0; error = "animation stages refer to each other in a cycle"; return false; }
            const bool ok = init_stage(a.stage);
            --depth_guard;
            if (!ok) return false;
        }
        for (auto& p : a.uniforms) sc.uniforms[size_t(p.first)] = sc.uniforms[size_t(p.second)];   // storage.set_id
        for (auto& p : a.matrices) sc.matrices[size_t(p.first)] = sc.matrices[size_t(p.second)];
        const int cs = get_start_cam(stage.index);
        if (cs >= 0) current_cam = cs;
    }
    sc.current_stage = stage;
    return true;
}

bool Player::init_stage_by_name(const std::string& name) {
    if (!sc.stages.count(name)) { error = "scene has no stage named `" + name + "`"; return false; }
    StageRef r;
    r.kind = StageRef::Animation;
    r.name = name;
    return init_stage(r);
}

bool Player::init_animation_by_name(const std::string& name) {
    auto it = sc.animation_by_name.find(name);
    if (it == sc.animation_by_name.end()) { error = "scene has no animation named `" + name + "`"; return false; }
    StageRef r;
    r.kind = StageRef::Real;
    r.index = it->second;
    if (!init_stage(r)) return false;
    update_inner_variables(name);
    return true;
}

void Player::update_inner_variables(const std::string& name) {
    static const char* const k500[] = {"v2.face.2", "v2.face.3", "v2.face.4", "v2.face.5", "v2.inside.1", "v2.inside.3", "v2.intro.1",
                                       "v2.normal.2", "v2.normal.3", "v2.rod.2", "v2.rod.3", "v2.spiral.3", "v2.spiral.4", "v2.spiral.5",
                                       "v2.spiral.6", "v2.spiral.7", "v2.spiral.9", "v2.spaaaace.0", "v4.golden.0", "v4.golden.1",
                                       "v4.golden.2", "v4.thumbnail.2"};
    static const char* const k1000[] = {"v2.spiral.4", "v2.spiral.5", "v2.spiral.6"};
    static const char* const kDepth100[] = {"v2.rotated.0", "v2.spiral.0", "v2.screenshot.5", "v2.screenshot.6"};
    auto in = [&](const char* const* list, size_t n) {
        for (size_t k = 0; k < n; k++)
            if (name == list[k]) return true;
        return false;
    };
    render_depth_override = fps_override = 0;
    // `find_id("subspace_degree")?`: a scene without that uniform leaves the function at once -- also before the depth / fps
    // overrides, exactly as the `?` does in the reference
    auto set_degree = [&](int v) {
        auto it = sc.uniform_by_name.find("subspace_degree");
        if (it == sc.uniform_by_name.end()) return false;
        Uniform u;
        u.kind = Uniform::Int;
        u.i = v;
        sc.uniforms[size_t(it->second)] = u;      // get_original_mut: the stored element itself
        return true;
    };
    if (in(k500, sizeof k500 / sizeof *k500) && !set_degree(500)) return;
    if (in(k1000, sizeof k1000 / sizeof *k1000) && !set_degree(1000)) return;
    if (in(kDepth100, sizeof kDepth100 / sizeof *kDepth100)) render_depth_override = 100;
    if (name == "v2.screenshot.3") fps_override = 600;
}

bool Player::select_camera(const std::string& name) {
    auto it = sc.camera_by_name.find(name);
    if (it == sc.camera_by_name.end()) { error = "scene has no camera named `" + name + "`"; return false; }
    current_cam = it->second;
    return true;
}

static double lerp_f(double a, double b, double t) { return (1.0 - t) * a + t * b; }  // emath::lerp

bool Player::scene_update(double time) {  // Scene::update, scene.rs:1353-1496
    double total_time;
    if (run_animations) {
        const double total = total_animation_duration();
        if (total > 0.0) {
            time = std::fmod(time, total);
            total_time = time;
        } else {
            time = 0.0;
            total_time = 0.0;
        }
        for (size_t idx = 0; idx < sc.animations.size(); idx++) {
            const double duration = sc.animations[idx].duration;
            if (time < duration) {
                StageRef r;
                r.kind = StageRef::Real;
                r.index = int(idx);
                if (sc.current_stage != r && !init_stage(r)) return false;
                time /= duration;
                break;
            }
            time -= duration;
        }
    } else if (sc.current_stage.kind == StageRef::Real) {
        const int idx = sc.current_stage.index;
        const double duration = sc.animations[size_t(idx)].duration;
        if (duration > 0.0) {
            const double local_seconds = std::fmod(time, duration);
            time = local_seconds / duration;
            double prefix = 0.0;
            for (int k = 0; k < idx; k++) prefix += sc.animations[size_t(k)].duration;
            total_time = prefix + local_seconds;
        } else {
            time = 0.0;
            total_time = 0.0;
        }
    } else {
        total_time = time;
    }
    sc.time = time;
    sc.total_time = total_time;

    if (sc.current_stage.kind == StageRef::Real) {
        const int idx = sc.current_stage.index;
        const RealAnimation& a = sc.animations[size_t(idx)];
        const int c1 = get_start_cam(idx), c2 = get_end_cam(idx);
        if (c1 >= 0 && c2 >= 0) {
            CalculatedCam cam1, cam2;
            if (!cam_get(c1, cam1) || !cam_get(c2, cam2)) { error = "an animation camera looks at a missing matrix"; return false; }
            const double t_raw = std::fmod(sc.time, 1.0);
            double t = 0.0;
            bool have_t = false;
            if (a.has_easing_uniform) {
                int kind;
                double v;
                std::vector<int> visited;
                if (sc.get_uniform(a.easing_uniform, kind, v, visited)) {
                    if (!std::isfinite(v)) v = 0.0;
                    t = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
                    have_t = true;
                }
            }
            if (!have_t) t = ease(a.cam_easing, t_raw);
            override_cam = CalculatedCam();
            for (int k = 0; k < 3; k++) override_cam.look_at[k] = cam1.look_at[k] + ((cam2.look_at[k] - cam1.look_at[k]) * t);  // DVec3::lerp
            override_cam.alpha = lerp_f(cam1.alpha, cam2.alpha, t);
            override_cam.beta = lerp_f(cam1.beta, cam2.beta, t);
            override_cam.r = lerp_f(cam1.r, cam2.r, t);
            override_cam.in_subspace = cam1.in_subspace;
            override_cam.free_movement = cam1.free_movement;
            override_cam.matrix = cam1.matrix;
            override_cam.override_matrix = t_raw < prev_t_raw || t_raw == 0.0;
            has_override = true;
            prev_t_raw = t_raw;
        }
    }
    return true;
}

bool Player::run_probe(const double a[3], const double b[3], double pos[3], bool& have, bool& enc, bool& chg) {
    n_probes++;
    if (!probe(a, b, pos, have, enc, chg)) {
        if (error.empty()) error = "external-ray probe failed";
        return false;
    }
    return true;
}

// main.rs:1174-1215.  `ok` = Some(..); the return value is false only on a device error.
bool Player::teleport_matrix(const Mat4& matrix, const double start_pos[3], const double direction_pos[3], const double actual[3],
                             double dx, Mat4& out, bool& ok) {
    ok = false;
    Mat4 nm{};
    for (int axis = 0; axis < 3; axis++) {
        double e[4] = {0.0, 0.0, 0.0, 0.0}, v[4];
        e[axis] = 1.0;
        mat_mul_vec(matrix, e, v);
        const double off[3] = {v[0] * dx, v[1] * dx, v[2] * dx};
        const double a[3] = {start_pos[0] + off[0], start_pos[1] + off[1], start_pos[2] + off[2]};
        const double b[3] = {direction_pos[0] + off[0], direction_pos[1] + off[1], direction_pos[2] + off[2]};
        double pos[3];
        bool have, enc, chg;
        if (!run_probe(a, b, pos, have, enc, chg)) return false;
        if (!have) return true;  // None
        const double rcp = 1.0 / dx;  // glam: DVec4 / f64 multiplies by the reciprocal
        for (int k = 0; k < 3; k++) nm[size_t(4 * axis + k)] = (pos[k] - actual[k]) * rcp;
        nm[size_t(4 * axis + 3)] = 0.0 * rcp;
    }
    nm[12] = 0.0; nm[13] = 0.0; nm[14] = 0.0; nm[15] = 1.0;
    const double dp[4] = {direction_pos[0], direction_pos[1], direction_pos[2], 1.0};
    double p[4];
    mat_mul_vec(mat_mul(nm, mat_inverse(matrix)), dp, p);
    out = nm;
    out[12] = actual[0] - p[0];
    out[13] = actual[1] - p[1];
    out[14] = actual[2] - p[2];
    out[15] = 1.0;
    ok = true;
    return true;
}

bool Player::teleport_camera(const OrbitCam& prev) {  // main.rs:1217-1264
    if (cam.do_not_teleport_one_frame) {
        cam.do_not_teleport_one_frame = false;
        cam.get_cam_pos(cam.prev_cam_pos);
        return true;
    }
    if (!(cam.allow_teleport || cam.stop_at_objects) || !probe) return true;
    double cam_pos[3], teleported[3];
    cam.get_cam_pos(cam_pos);
    bool have, enc, chg;
    if (!run_probe(cam.prev_cam_pos, cam_pos, teleported, have, enc, chg)) return false;
    if (cam.stop_at_objects && enc) {
        cam = prev;
        return true;
    }
    if (have) {
        if (!cam.allow_teleport) return true;
        for (double dx : {0.001, 0.0001, 0.00001, 0.000001}) {
            Mat4 m;
            bool ok;
            if (!teleport_matrix(cam.teleport_matrix, cam.prev_cam_pos, cam_pos, teleported, dx, m, ok)) return false;
            if (!ok) continue;
            cam.teleport_matrix = m;
            if (chg) cam.in_subspace = !cam.in_subspace;
            cam.get_cam_pos(cam.prev_cam_pos);
            return true;
        }
        cam = prev;
    } else {
        for (int k = 0; k < 3; k++) cam.prev_cam_pos[k] = cam_pos[k];
    }
    return true;
}

bool Player::teleport_eye_matrices() {  // main.rs:1121-1172
    if (!((draw_anaglyph || draw_side_by_side) && cam.allow_teleport)) return true;   // main.rs:1122
    const double ed = swap_eyes ? -eye_distance : eye_distance;
    auto one = [&](double eye_x, Mat4& matrix, bool& sub) -> bool {
        double start[3], p[4];
        cam.get_cam_pos(start);
        const Mat4 m = cam.get_matrix();
        const double ev[4] = {eye_x, 0.0, 0.0, 1.0};
        mat_mul_vec(m, ev, p);
        const double dp[3] = {p[0], p[1], p[2]};
        Mat4 tr = mat_identity();
        tr[12] = dp[0] - start[0]; tr[13] = dp[1] - start[1]; tr[14] = dp[2] - start[2];
        matrix = mat_mul(tr, m);
        sub = cam.in_subspace;
        if (!probe) return true;
        double teleported[3];
        bool have, enc, chg;
        if (!run_probe(start, dp, teleported, have, enc, chg)) return false;
        if (have) {
            for (double dx : {0.001, 0.0001, 0.00001, 0.000001}) {
                Mat4 mm;
                bool ok;
                if (!teleport_matrix(matrix, start, dp, teleported, dx, mm, ok)) return false;
                if (!ok) continue;
                matrix = mm;
                if (chg) sub = !cam.in_subspace;
                break;
            }
        }
        return true;
    };
    return one(-ed, left_eye_matrix, left_eye_in_subspace) && one(ed, right_eye_matrix, right_eye_in_subspace);
}

bool Player::update(double time) {  // SceneRenderer::update, main.rs:1430-1543
    if (!scene_update(time)) return false;
    sc.camera_matrix_for_formulas = cam.get_matrix();  // send_camera_object_matrix (default true)
    if (cam.from != current_cam) {
        CalculatedCam cc;
        if (current_cam >= 0) {
            if (cam.from < 0) original_cam = cam.get_calculated_cam();
            if (!cam_get(current_cam, cc)) { error = "the selected camera looks at a missing matrix"; return false; }
        } else {
            cc = original_cam;
        }
        cam.from = current_cam;
        cam.alpha = cc.alpha; cam.beta = cc.beta; cam.r = cc.r;
        for (int k = 0; k < 3; k++) cam.look_at[k] = cc.look_at[k];
        cam.teleport_matrix = cc.matrix;
        cam.in_subspace = cc.in_subspace;
        cam.free_movement = cc.free_movement;
        if (cam.free_movement) {
            double pv[3];
            cam.get_pos_vec(pv);
            for (int k = 0; k < 3; k++) cam.look_at[k] = pv[k] + cam.look_at[k];
        }
        cam.do_not_teleport_one_frame = true;
    } else if (cam.from >= 0) {
        CalculatedCam cc;
        if (!cam_get(cam.from, cc)) { error = "the selected camera looks at a missing matrix"; return false; }
        if (!cam.free_movement)
            for (int k = 0; k < 3; k++) cam.look_at[k] = cc.look_at[k];
    }
    if (has_override) {
        cam.alpha = override_cam.alpha; cam.beta = override_cam.beta; cam.r = override_cam.r;
        for (int k = 0; k < 3; k++) cam.look_at[k] = override_cam.look_at[k];
        cam.free_movement = override_cam.free_movement;
        if (override_cam.override_matrix) {
            cam.teleport_matrix = override_cam.matrix;
            cam.in_subspace = override_cam.in_subspace;
            cam.do_not_teleport_one_frame = true;
        }
        has_override = false;
    }
    if (cam.get_matrix() != prev_cam.get_matrix()) {
        const OrbitCam prev = prev_cam;
        if (!teleport_camera(prev)) return false;
    }
    if (!teleport_eye_matrices()) return false;
    prev_cam = cam;
    sc.camera_matrix_for_formulas = cam.get_matrix();
    return true;
}

}  // namespace ph
