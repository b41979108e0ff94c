// Host-side scene model: loader, uniform / matrix evaluation, camera.
//
// Mirrors the parts of the reference's Rust host that feed the ray loop every frame
// (SURVEY.md §8a rows U, M, C):
//   scene load + name resolution   /root/reference/src/gui/scene_serialized.rs:1102-1477
//   AnyUniform evaluation          /root/reference/src/gui/uniform.rs:268-296, :1009-1140
//   Matrix DAG evaluation (f64)    /root/reference/src/gui/matrix.rs:510-631
//   uniform table + upload order   /root/reference/src/gui/scene.rs:424-658
//   orbit camera                   /root/reference/src/main.rs:278-304, :1325-1333
// glam 0.13.1's DMat4 / DQuat arithmetic (un-vendored) is implemented operation by operation
// in ph_scene.cpp so that singular matrices give the same Inf/NaN patterns.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "ph_formula.h"
#include "ph_ron.h"

namespace ph {

using Mat4 = std::array<double, 16>;  // column-major

Mat4 mat_identity();
Mat4 mat_mul(const Mat4& a, const Mat4& b);
Mat4 mat_inverse(const Mat4& m);
Mat4 mat_srt(const double scale[3], const double rotate[3], const double offset[3]);
Mat4 mat_lerp(const Mat4& first, const Mat4& second, double t);
// matrix kind Sqrt (matrix.rs:909-985): BFGS minimisation of |X*X - m|^2 from X0 = m; false = "Can't calculate sqrt!"
bool mat_sqrt(const Mat4& m, Mat4& out);
Mat4 orbit_camera_matrix(const double look_at[3], double alpha, double beta, double r);
// RotateAroundCam::get_matrix with a teleport matrix and the free-movement flag (main.rs:286-304)
Mat4 orbit_camera_matrix(const double look_at[3], double alpha, double beta, double r, const Mat4& teleport, bool free_movement);
void mat_mul_vec(const Mat4& m, const double v[4], double out[4]);
double camera_scale(const Mat4& m);

struct Uniform {
    enum Kind { Bool, Int, Float, Angle, Progress, Formula, FormulaInt, Trefoil, Unsupported } kind = Float;
    int trefoil[18] = {0};  // TrefoilSpecial([(bool, u8, u8); 18]) (uniform.rs:20), packed as scene.rs:644-650 uploads it
    bool b = false;
    int i = 0;
    double f = 0.0;
    std::string text;
    FormulaPtr parsed;  // lazily parsed
    bool parse_failed = false;
};

struct Param {  // ParametrizeOrNot (uniform.rs:487-497)
    bool is_uniform = false;
    double value = 0.0;
    int uniform = -1;  // index into Scene::uniforms, -1 = None
};

struct Matrix {
    enum Kind { Mul, Teleport, Simple, Parametrized, Exact, ExactFull, If, Sqrt, Lerp, Camera, Inv } kind = Simple;
    int a = -1, b = -1, c = -1;       // matrix refs (Mul: to, what; Teleport: first, second, what; If: then, otherwise; Inv/Sqrt: a)
    double offset[3] = {0, 0, 0}, rotate[3] = {0, 0, 0}, scale = 1.0;
    bool mirror[3] = {false, false, false};
    Param p[16];                       // Parametrized: offset[0..2] rotate[3..5] mirror[6..8] scale[9]; Exact*: columns; If/Lerp: p[0]
};

struct SceneObject {
    std::string name;
    enum Class { Flat, Complex, DebugMatrix } cls = Flat;
    bool portal = false;
    int subspace = 0;
    int matrix_a = -1, matrix_b = -1;
    std::string code;
};

struct SceneMaterial {
    std::string name;
    enum Type { Simple, Reflect, Refract, Complex } type = Simple;
    double color[3] = {0, 0, 0};
    double normal_coef = 0, grid_scale = 0, grid_coef = 0, refractive_index = 0;
    bool grid = false, grid2 = false, grid3 = false;
    std::string code;
};

struct TableEntry {
    std::string name;
    int type = 0;  // PE_UNIFORM_*
    Mat4 m{};
    double f = 0.0;
    int i = 0;
};

struct StageAnim {      // StageAnimSer (scene_serialized.rs:508-513) after name resolution
    bool from_dev = true;   // FromDev / ProvidedToUser: restore the dev-stage value
    int element = -1;       // Changed / ChangedAndToUser(Some(x)): copy element x over the target; -1 = None
};
struct Stage {
    std::vector<std::pair<int, StageAnim>> uniforms, matrices;  // (target id, what to do)
};

// Cam (camera.rs:62-76)
struct Cam {
    bool look_at_matrix = false;  // CamLookAt::MatrixCenter(id) / Coordinate(pos)
    double coord[3] = {0, 0, 0};
    int matrix = -1;
    double alpha = 0, beta = 0, r = 3.5;
    bool in_subspace = false, free_movement = false;
    Mat4 teleport = mat_identity();
};

struct StageRef {  // CurrentStage (scene.rs) after name resolution
    enum Kind { Dev, Animation, Real } kind = Dev;
    std::string name;   // Animation: stage name
    int index = -1;     // Real: index into Scene::animations
    bool operator==(const StageRef& o) const { return kind == o.kind && name == o.name && index == o.index; }
    bool operator!=(const StageRef& o) const { return !(*this == o); }
};

struct RealAnimation {  // animation.rs:1015-1045 after name resolution
    std::string name;
    double duration = 0.0;
    StageRef stage;
    std::vector<std::pair<int, int>> uniforms, matrices;  // Changed(Some(x)) parts only: (target id, element x)
    bool use_prev_cam = false, use_start_cam_as_end = false;
    int cam_start = -1, cam_end = -1;
    int use_any_cam_as_start = -1, use_any_cam_as_end = -1;  // -1 None, 0 Some(false), 1 Some(true)
    int cam_any_start = -1, cam_any_end = -1;
    int cam_easing = 0;              // Easing: 0 Linear 1 In 2 Out 3 InOut 4 InOutFast 5 ElasticOut
    bool has_easing_uniform = false;
    int easing_uniform = -1;
};

double ease(int easing, double t);

struct Scene {
    // saved camera (CamSettings, scene.rs) + per-scene `_offset_after_material`
    double look_at[3] = {0, 0, 0}, alpha = 0, beta = 0, r = 1, offset_after_material = 0.005;
    bool use_time = false;
    double time = 0.0, total_time = 0.0;
    Mat4 camera_matrix_for_formulas = mat_identity();

    std::vector<Uniform> uniforms;
    std::vector<std::string> uniform_names;  // "" for inline
    std::map<std::string, int> uniform_by_name;
    std::vector<Matrix> matrices;
    std::vector<std::string> matrix_names;   // "" for inline
    std::map<std::string, int> matrix_by_name;
    std::vector<SceneObject> objects;
    std::vector<SceneMaterial> materials;
    std::vector<std::pair<std::string, std::string>> intersection_materials, library, textures;  // (name, code|path)
    std::string skybox;  // texture name, empty = none
    // Videos (video.rs:14-20): one more sampler2D each, named like a texture (scene.rs:405-409, 709-711); the host replaces its
    // image per frame with the frame the uniform selects (main.rs:862-914).
    struct Video { std::string name, path; int uniform = -1; };
    std::vector<Video> videos;
    // VideoRuntime::update (main.rs:862-895): round((count - 1) * clamp(uniform, 0, 1)); false = no uniform / no value / no frames
    bool video_frame(int k, size_t frame_count, size_t& index);
    std::string error;
    std::map<std::string, Stage> stages;                 // animation stages by name
    std::map<int, Uniform> dev_uniforms;                 // dev stage: target id -> value
    std::map<int, Matrix> dev_matrices;
    // Scene::init_stage(CurrentStage::Animation(id)) (scene.rs:1180-1200, animation.rs:171-183)
    bool init_stage(const std::string& name);
    // cameras, stage cameras, real animations (scene_serialized.rs:1286-1473)
    std::vector<Cam> cameras;
    std::map<std::string, int> camera_by_name;
    std::map<std::string, int> stage_cam;                // stage name -> camera id (-1 = original camera)
    std::vector<RealAnimation> animations;               // in file (= visible) order
    std::map<std::string, int> animation_by_name;
    StageRef current_stage;
    void init_dev_stage();                               // DevStageChanging::init_stage (animation.rs:223-227)

    // deserialize_scene_new_format
    bool load(const RonValue& root);
    std::string matrix_uniform_stem(int id) const;  // object.rs:188-193: name or "id<N>"
    // AnyUniform::get -> (kind 0 bool / 1 int / 2 float, value); false = None
    bool get_uniform(int id, int& kind, double& value, std::vector<int>& visited);
    bool get_param(const Param& p, double& out);
    // Matrix::get; false = None
    bool get_matrix(int id, Mat4& out, std::vector<int>& visited);
    // Scene::uniforms + Scene::set_uniforms: the evaluated table in upload order
    bool uniform_table(std::vector<TableEntry>& out);
    // Hash of everything uniform_table() reads (time, formula camera, every uniform's and matrix's fields): when it has not
    // changed since the last evaluation the table has not either -- a frame loop that only moves the camera skips the
    // float64 evaluation of the whole matrix DAG.
    uint64_t input_fingerprint() const;
};

}  // namespace ph
