// The animation-facing half of the reference's SceneRenderer, on the host, in float64:
//   Cam::get                         /root/reference/src/gui/camera.rs:111-140
//   Scene::init_stage                src/gui/scene.rs:1180-1236 (stages, dev stage, real animations)
//   get_start_cam / get_end_cam      src/gui/scene.rs:1291-1343
//   Scene::update                    src/gui/scene.rs:1353-1496 (time mapping, camera interpolation)
//   SceneRenderer::new / update      src/main.rs:1014-1062, 1430-1543
//   teleport_camera, teleport_matrix src/main.rs:1217-1264, 1174-1215 (on top of the external-ray probe)
// The egui `Memory` entries the reference threads through these functions (CurrentCam, OriginalCam,
// OverrideCam) are plain members here.
#pragma once
#include <functional>
#include <string>

#include "ph_scene.h"

namespace ph {

struct CalculatedCam {  // camera.rs:44-60
    double look_at[3] = {0, 0, 0};
    double alpha = 0, beta = 0, r = 0;
    bool in_subspace = false, free_movement = false;
    Mat4 matrix = mat_identity();
    bool override_matrix = true;
};

struct OrbitCam {  // the parts of RotateAroundCam (main.rs:40-160) the offline paths touch
    double look_at[3] = {0, 0, 0};
    double alpha = 81.0 / 180.0 * 3.14159265358979323846, beta = 64.0 / 180.0 * 3.14159265358979323846, r = 3.5;
    int from = -1;  // CameraId or -1 (None)
    Mat4 teleport_matrix = mat_identity();
    bool allow_teleport = true, stop_at_objects = false;
    double prev_cam_pos[3] = {0, 0, 0};
    bool in_subspace = false, free_movement = false, do_not_teleport_one_frame = false;

    void get_pos_vec(double out[3]) const;
    Mat4 get_matrix() const;
    void get_cam_pos(double out[3]) const;
    CalculatedCam get_calculated_cam() const;
};

// teleport_external_ray (main.rs:1361-1409) with the player's current uniforms: returns false on a
// device error; `have` = a portal was crossed and `pos` is where b lands.
using ProbeFn = std::function<bool(const double a[3], const double b[3], double pos[3], bool& have, bool& encounter_object,
                                   bool& change_subspace)>;

struct Player {
    explicit Player(Scene& scene);  // SceneRenderer::new: camera from the scene's CamSettings
    Scene& sc;
    std::string error;
    ProbeFn probe;  // empty: the camera never teleports
    long long n_probes = 0;

    // memory
    int current_cam = -1;
    CalculatedCam original_cam;
    bool has_override = false;
    CalculatedCam override_cam;

    bool run_animations = false;
    double prev_t_raw = 0.0;
    OrbitCam cam, prev_cam;
    // stereo (main.rs:1027-1030, teleport_eye_matrices :1121-1172)
    bool draw_side_by_side = false, swap_eyes = false;
    bool draw_anaglyph = false, anaglyph_mode = false;   // main.rs:1030-1033
    double anaglyph_p = 0.29, anaglyph_q = 0.06;
    double eye_distance = 0.07;
    Mat4 left_eye_matrix = mat_identity(), right_eye_matrix = mat_identity();
    bool left_eye_in_subspace = false, right_eye_in_subspace = false;
    bool teleport_eye_matrices();

    bool init_stage(const StageRef& stage);
    bool init_stage_by_name(const std::string& name);
    bool init_animation_by_name(const std::string& name);
    // SceneRenderer::update_inner_variables (main.rs:1696-1755), which render-frame and the `render` loop call right after
    // init_animation_by_name: overrides hard-coded by animation name -- the `subspace_degree` uniform (500 / 1000), the render
    // depth (100) and the frame rate (600).  0 = no override.
    void update_inner_variables(const std::string& animation_name);
    int render_depth_override = 0, fps_override = 0;
    bool select_camera(const std::string& name);
    bool update(double time);  // SceneRenderer::update

    bool cam_get(int cam_id, CalculatedCam& out);
    int get_start_cam(int anim, int depth = 0) const;   // depth: guards against animations whose cameras refer to each other in a cycle
    int get_end_cam(int anim, int depth = 0) const;
    double total_animation_duration() const;

  private:
    bool scene_update(double time);
    bool teleport_camera(const OrbitCam& prev);
    bool teleport_matrix(const Mat4& matrix, const double start_pos[3], const double direction_pos[3], const double actual[3],
                         double dx, Mat4& out, bool& ok);
    bool run_probe(const double a[3], const double b[3], double pos[3], bool& have, bool& enc, bool& chg);
    int depth_guard = 0;
};

}  // namespace ph
