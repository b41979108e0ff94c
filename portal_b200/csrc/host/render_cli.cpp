// portal_b200_render -- command-line mirror of the reference's `portal render-frame`
// (/root/reference/src/main.rs:2736-2760 options, :2876-2946 render_frame), C++ host + C ABI only:
//
//   portal_b200_render render-frame <scene.ron> [--width W] [--height H] [--render-depth D] [--aa-count N]
//                      [--time T] [--stage NAME] [--animation NAME] [--camera NAME] [--device K]
//                      [--texture name=file.png | name=file.rgba:WxH ...] [--assets DIR] [--output out.png|.ppm|.rgba]
//   portal_b200_render render <scene.ron> [--animations a,b,... | --starts-with PREFIX] [--fps N] [--motion-blur-frames M] [--width W]
//                      [--height H] [--render-depth D] [--aa-count N] [--stereo-image] [--no-skip-existing] [--out-dir DIR] [--max-frames K]
//                      [--frame-format ppm|png|rgba] [--assets DIR]
//                      (`portal render`, main.rs:2808-2873 -> render_named_animations :1876-1930 ->
//                       render_animation :1757-1830; frames are written as DIR/<animation>/frame_<i>.<format>, the
//                       ffmpeg step is out of scope)
//
// Output: the RGBA8 frame the reference hands to export_png (main.rs:2939-2943) as a PNG when the name ends in .png
// (ph_png_encode_rgba8), as raw RGBA8 for .rgba, else as a binary PPM (P6, alpha dropped).  Textures: every texture of the
// scene is loaded from its stored path (a PNG, relative to --assets, default the working directory) like reload_textures
// (main.rs:1066-1085) -- a file that cannot be read is reported and the sampler stays unbound, as there; --texture overrides one
// by name (PNG, or raw RGBA8 with :WxH).  Videos: the frames of video_png/<stem>/*.png, sorted, the one the video's uniform
// selects swapped in after every update (VideoRuntime, main.rs:771-930).
#include <dirent.h>
#include <sys/stat.h>
#include <sys/types.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/portal_b200_host.h"

static std::string slurp(const std::string& path, bool& ok) {
    std::ifstream f(path, std::ios::binary);
    ok = bool(f);
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

static bool ends_with(const std::string& s, const char* suffix) {
    const size_t n = std::strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

static bool write_image(const std::string& output, const std::vector<uint8_t>& px, int width, int height) {
    std::ofstream out(output, std::ios::binary);
    if (!out) return false;
    if (ends_with(output, ".png")) {             // Image::export_png (main.rs:2939-2943)
        uint8_t* png = nullptr;
        size_t n = 0;
        if (ph_png_encode_rgba8(px.data(), width, height, &png, &n)) return false;
        out.write(reinterpret_cast<const char*>(png), std::streamsize(n));
        ph_png_free(png);
    } else if (ends_with(output, ".rgba")) {
        out.write(reinterpret_cast<const char*>(px.data()), std::streamsize(px.size()));
    } else {
        out << "P6\n" << width << " " << height << "\n255\n";
        for (size_t i = 0; i < size_t(width) * size_t(height); i++) out.write(reinterpret_cast<const char*>(&px[4 * i]), 3);
    }
    return bool(out);
}

// One PNG file -> the sampler `name` (Texture2D::from_file_with_format + material.set_texture, main.rs:1072-1078)
static bool set_texture_from_png(pe_ctx* ctx, const std::string& name, const std::string& path, std::string& why) {
    bool ok = false;
    const std::string bytes = slurp(path, ok);
    if (!ok) { why = "cannot read " + path; return false; }
    uint8_t* px = nullptr;
    int32_t w = 0, h = 0;
    char err[256] = {0};
    if (ph_png_decode(reinterpret_cast<const uint8_t*>(bytes.data()), bytes.size(), &px, &w, &h, err, sizeof err)) { why = path + ": " + err; return false; }
    const int rc = pe_set_texture(ctx, name.c_str(), px, w, h);
    ph_png_free(px);
    if (rc) { why = pe_last_error(ctx); return false; }
    return true;
}

// VideoRuntime (main.rs:771-930): the sorted PNG frames of one video and the frame last uploaded
struct VideoRuntime {
    int video = 0;
    std::string name;
    std::vector<std::string> frames;
    long last = -1;
};
struct VideoState {
    ph_scene* scene = nullptr;
    pe_ctx* ctx = nullptr;
    std::vector<VideoRuntime> videos;
};

// video_collect_frame_files (main.rs:795-817): every *.png of video_png/<stem>, sorted by path
static std::vector<std::string> video_frame_files(const std::string& assets, const char* video_path) {
    std::vector<std::string> files;
    char rel[1024];
    if (!ph_video_frames_dir(video_path, rel, sizeof rel)) return files;
    const std::string dir = assets.empty() ? std::string(rel) : assets + "/" + rel;
    if (DIR* d = ::opendir(dir.c_str())) {
        while (const dirent* e = ::readdir(d)) {
            const std::string f = e->d_name;
            if (ends_with(f, ".png")) files.push_back(dir + "/" + f);
        }
        ::closedir(d);
    }
    std::sort(files.begin(), files.end());
    return files;
}

// VideoRuntime::update for every video (main.rs:862-914); the player calls it after each update (ph_player_set_update_hook)
static int update_videos(void* user) {
    VideoState& st = *static_cast<VideoState*>(user);
    for (VideoRuntime& v : st.videos) {
        uint64_t idx = 0;
        if (ph_scene_video_frame(st.scene, v.video, v.frames.size(), &idx)) continue;
        if (long(idx) == v.last) continue;                     // avoid a reload when the frame did not change
        std::string why;
        if (!set_texture_from_png(st.ctx, v.name, v.frames[size_t(idx)], why)) continue;   // IO errors are ignored there too: the texture just does not update
        v.last = long(idx);
    }
    return 0;
}

static bool file_exists(const std::string& path) {
    struct stat st;
    return ::stat(path.c_str(), &st) == 0;
}

// std::fs::create_dir_all (main.rs:1781-1783)
static bool make_dirs(const std::string& path) {
    for (size_t i = 1; i <= path.size(); i++) {
        if (i != path.size() && path[i] != '/') continue;
        const std::string part = path.substr(0, i);
        if (::mkdir(part.c_str(), 0755) != 0 && errno != EEXIST) return false;
    }
    struct stat st;
    return ::stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

int main(int argc, char** argv) {
    const bool frame_cmd = argc >= 3 && std::strcmp(argv[1], "render-frame") == 0;
    const bool anim_cmd = argc >= 3 && std::strcmp(argv[1], "render") == 0;
    if (!frame_cmd && !anim_cmd) {
        std::fprintf(stderr, "usage: %s render-frame <scene.ron> [--width W] [--height H] [--render-depth D] [--aa-count N] [--time T] "
                             "[--stage NAME] [--animation NAME] [--camera NAME] [--device K] [--texture name=file.png|name=file.rgba:WxH] [--assets DIR] [--output out.png|.ppm|.rgba]\n"
                             "       %s render <scene.ron> [a,b | --animations a,b | --starts-with PREFIX] [--fps N] [--motion-blur-frames M] [--width W] [--height H] "
                             "[--render-depth D] [--aa-count N] [--stereo-image] [--no-skip-existing] [--out-dir DIR] [--max-frames K]\n", argv[0], argv[0]);
        return 2;
    }
    std::string scene_path = argv[2], output = "frame.ppm", out_dir = "video";
    // defaults of RenderFrameCliOptions (main.rs:2744-2754) / RenderCliOptions (main.rs:2764-2803)
    int width = frame_cmd ? 1920 : 3840, height = frame_cmd ? 1080 : 2160, depth = frame_cmd ? 100 : 150, aa = frame_cmd ? 1 : 4, device = 0;
    int fps = 60, motion_blur = 1, max_frames = -1;
    bool stereo = false, skip_existing = true;
    double time = 0.0;
    std::vector<std::string> textures;
    std::string stage, animation, camera, animations, starts_with, assets, frame_ext = ".ppm";
    int first_option = 3;
    if (anim_cmd && argc > 3 && std::strncmp(argv[3], "--", 2) != 0) animations = argv[first_option++];   // `render <scene> [animations]`
    for (int i = first_option; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--width") width = std::atoi(next());
        else if (a == "--height") height = std::atoi(next());
        else if (a == "--render-depth" || a == "--render_depth") depth = std::atoi(next());
        else if (a == "--aa-count" || a == "--aa_count") aa = std::atoi(next());
        else if (a == "--time") time = std::atof(next());
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--stage") stage = next();
        else if (a == "--animation") animation = next();
        else if (a == "--camera") camera = next();
        else if (a == "--animations") animations = next();
        else if (a == "--starts-with" || a == "--filter-starts-with" || a == "--starts_with" || a == "--filter_starts_with") starts_with = next();
        else if (a == "--fps") fps = std::atoi(next());
        else if (a == "--motion-blur-frames" || a == "--motion_blur_frames") motion_blur = std::atoi(next());
        else if (a == "--out-dir") out_dir = next();
        else if (a == "--max-frames") max_frames = std::atoi(next());
        else if (a == "--stereo-image" || a == "--stereoimage" || a == "--stereo_image") stereo = true;
        else if (a == "--no-skip-existing" || a == "--no_skip_existing") skip_existing = false;
        else if (a == "--output") output = next();
        else if (a == "--texture") textures.push_back(next());
        else if (a == "--assets") assets = next();
        else if (a == "--frame-format") frame_ext = std::string(".") + next();          // render: ppm (default) | png | rgba
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    bool ok = false;
    std::string text = slurp(scene_path, ok);
    if (!ok) { std::fprintf(stderr, "cannot read %s\n", scene_path.c_str()); return 1; }
    char err[2048] = {0};
    ph_scene* scene = ph_scene_load_ron(text.data(), text.size(), err, sizeof err);
    if (!scene) { std::fprintf(stderr, "Failed to parse scene `%s`: %s\n", scene_path.c_str(), err); return 1; }
    pe_ctx* ctx = pe_create(device);
    if (!ctx) { std::fprintf(stderr, "pe_create: %s\n", pe_last_error(nullptr)); return 1; }
    ph_player* player = ph_player_new(scene);
    if (frame_cmd) {  // main.rs:2900-2926
        if (!stage.empty() && ph_player_init_stage(player, stage.c_str())) {
            std::fprintf(stderr, "Scene `%s` has no stage named `%s`\n", scene_path.c_str(), stage.c_str());
            return 1;
        }
        if (!animation.empty() && ph_player_init_animation(player, animation.c_str())) {
            std::fprintf(stderr, "Scene `%s` has no animation named `%s`\n", scene_path.c_str(), animation.c_str());
            return 1;
        }
        if (!camera.empty() && ph_player_select_camera(player, camera.c_str())) {
            std::fprintf(stderr, "Scene `%s` has no camera named `%s`\n", scene_path.c_str(), camera.c_str());
            return 1;
        }
    }
    if (ph_scene_build_program(scene, ctx) || ph_scene_upload_uniforms(scene, ctx)) {
        std::fprintf(stderr, "%s\n", ph_scene_last_error(scene));
        return 1;
    }
    std::map<std::string, bool> given;
    for (const std::string& t : textures) {  // name=file.png | name=file.rgba:WxH
        size_t eq = t.find('='), colon = t.rfind(':'), x = t.rfind('x');
        if (eq != std::string::npos && ends_with(t, ".png")) {
            std::string why;
            if (!set_texture_from_png(ctx, t.substr(0, eq), t.substr(eq + 1), why)) { std::fprintf(stderr, "texture %s: %s\n", t.c_str(), why.c_str()); return 1; }
            given[t.substr(0, eq)] = true;
            continue;
        }
        if (eq == std::string::npos || colon == std::string::npos || x == std::string::npos || x < colon) { std::fprintf(stderr, "bad --texture %s\n", t.c_str()); return 2; }
        given[t.substr(0, eq)] = true;
        int tw = std::atoi(t.substr(colon + 1, x - colon - 1).c_str()), th = std::atoi(t.substr(x + 1).c_str());
        bool tok = false;
        std::string bytes = slurp(t.substr(eq + 1, colon - eq - 1), tok);
        if (!tok || bytes.size() != size_t(tw) * size_t(th) * 4) { std::fprintf(stderr, "texture %s: cannot read / wrong size\n", t.c_str()); return 1; }
        if (pe_set_texture(ctx, t.substr(0, eq).c_str(), reinterpret_cast<const uint8_t*>(bytes.data()), tw, th)) { std::fprintf(stderr, "%s\n", pe_last_error(ctx)); return 1; }
    }
    // reload_textures (main.rs:1066-1085): every texture of the scene from its stored path; a failure is reported, the render goes on
    for (int k = 0;; k++) {
        const char *name = nullptr, *path = nullptr;
        if (ph_scene_texture(scene, k, &name, &path)) break;
        if (given.count(name)) continue;
        const std::string file = assets.empty() || (path[0] == '/') ? std::string(path) : assets + "/" + path;
        std::string why;
        if (!set_texture_from_png(ctx, name, file, why)) std::fprintf(stderr, "texture `%s`: %s (the sampler stays unbound)\n", name, why.c_str());
    }
    if (pe_scene_compile(ctx)) { std::fprintf(stderr, "%s\n", pe_last_error(ctx)); return 1; }
    ph_player_attach(player, ctx);  // camera teleportation goes through pe_probe_ray
    // VideoRuntime::from_scene (main.rs:820-858): videos that have a path, a uniform and frames on disk
    VideoState video_state;
    video_state.scene = scene;
    video_state.ctx = ctx;
    for (int k = 0;; k++) {
        const char *name = nullptr, *path = nullptr, *uniform = nullptr;
        if (ph_scene_video(scene, k, &name, &path, &uniform)) break;
        if (!path[0] || !uniform[0]) continue;
        VideoRuntime v;
        v.video = k;
        v.name = name;
        v.frames = video_frame_files(assets, path);
        if (!v.frames.empty()) video_state.videos.push_back(v);
    }
    if (!video_state.videos.empty()) ph_player_set_update_hook(player, update_videos, &video_state);
    if (stereo) {                   // `render --stereo-image`: side-by-side eyes, doubled width (main.rs:2809-2816, 2842)
        width *= 2;
        ph_player_set_stereo(player, 1, 0.07, 0);
    }
    std::vector<uint8_t> px(size_t(width) * size_t(height) * 4);
    ph_frame_params p = {width, height, depth, aa, 0, 0, {0, 0, 0}, 0, 0, 0};
    if (frame_cmd) {
        if (ph_player_update(player, time) || ph_player_render_frame(player, ctx, &p, px.data(), 1)) {   // main.rs:2928-2929
            std::fprintf(stderr, "%s\n", ph_player_last_error(player));
            return 1;
        }
        if (!write_image(output, px, width, height)) { std::fprintf(stderr, "cannot write %s\n", output.c_str()); return 1; }
        std::printf("Rendered `%s` to `%s`\n", scene_path.c_str(), output.c_str());
    } else {
        // --animations a,b: render_named_animations (main.rs:1876-1930); otherwise render_all_animations (:1932-1974): every
        // animation in file order is initialised and updated once, and rendered unless --starts-with filters it out
        std::vector<std::string> names;
        const bool all = animations.empty();
        if (all) {
            for (int a = 0; a < ph_scene_animation_count(scene); a++) {
                const char* nm = nullptr;
                if (ph_scene_animation(scene, a, &nm, nullptr) == 0) names.push_back(nm);
            }
        } else {
            std::stringstream ss(animations);
            for (std::string item; std::getline(ss, item, ',');) if (!item.empty()) names.push_back(item);
        }
        for (size_t k = 0; k < names.size(); k++) {
            if (ph_player_init_animation(player, names[k].c_str())) {
                std::fprintf(stderr, "Scene `%s` has no animation named `%s`\n", scene_path.c_str(), names[k].c_str());
                return 1;
            }
            if (ph_player_update(player, 0.0)) { std::fprintf(stderr, "%s\n", ph_player_last_error(player)); return 1; }
            double duration = 0.0;
            for (int a = 0; a < ph_scene_animation_count(scene); a++) {
                const char* nm = nullptr;
                double d = 0.0;
                if (ph_scene_animation(scene, a, &nm, &d) == 0 && names[k] == nm) duration = d;
            }
            if (all && !starts_with.empty() && names[k].compare(0, starts_with.size(), starts_with) != 0) continue;
            std::printf("Rendering animation %s, %zu/%zu\n", names[k].c_str(), k + 1, names.size());
            const std::string dir = out_dir + "/" + names[k];
            if (!make_dirs(dir)) { std::fprintf(stderr, "cannot create %s\n", dir.c_str()); return 1; }
            const float duration32 = float(duration);                                          // render_animation takes f32
            int count = int(duration32 * float(fps));                                           // main.rs:1785
            if (count < 1) count = 1;
            const int todo = max_frames >= 0 && max_frames < count ? max_frames : count;
            for (int i = 0; i < todo; i++) {
                if (skip_existing && file_exists(dir + "/frame_" + std::to_string(i) + frame_ext)) continue;   // resumes like main.rs:1789-1793
                if (ph_player_render_motion_blur_frame(player, ctx, &p, i, count, motion_blur, double(duration32), px.data())) {
                    std::fprintf(stderr, "%s\n", ph_player_last_error(player));
                    return 1;
                }
                const std::string name = dir + "/frame_" + std::to_string(i) + frame_ext;
                if (!write_image(name, px, width, height)) { std::fprintf(stderr, "cannot write %s\n", name.c_str()); return 1; }
                std::printf("\r%d/%d done      ", i + 1, count);
                std::fflush(stdout);
            }
            std::printf("\n");
        }
    }
    ph_player_free(player);
    pe_destroy(ctx);
    ph_scene_free(scene);
    return 0;
}
