// portal_b200_render -- command-line mirror of the reference's `portal render-frame`
// (/root/reference/src/main.rs:2736-2760 options, :2876-2946 render_frame), C++ host + C ABI only:
//
//   portal_b200_render render-frame <scene.ron> [--width W] [--height H] [--render-depth D] [--aa-count N]
//                      [--time T] [--stage NAME] [--device K] [--texture name=file.rgba:WxH ...] [--output out.ppm]
//
// Output: binary PPM (P6) of the RGBA8 frame the reference would hand to export_png (alpha dropped), or
// raw RGBA8 with a .rgba extension.  PNG encode/decode is out of scope: textures
// are passed as raw RGBA8 files.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
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

int main(int argc, char** argv) {
    if (argc < 3 || std::strcmp(argv[1], "render-frame") != 0) {
        std::fprintf(stderr, "usage: %s render-frame <scene.ron> [--width W] [--height H] [--render-depth D] [--aa-count N] "
                             "[--time T] [--device K] [--texture name=file.rgba:WxH] [--output out.ppm]\n", argv[0]);
        return 2;
    }
    std::string scene_path = argv[2], output = "frame.ppm";
    int width = 1920, height = 1080, depth = 100, aa = 1, device = 0;  // defaults of RenderFrameCliOptions, main.rs:2744-2754
    double time = 0.0;
    std::vector<std::string> textures;
    std::string stage;
    for (int i = 3; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--width") width = std::atoi(next());
        else if (a == "--height") height = std::atoi(next());
        else if (a == "--render-depth") depth = std::atoi(next());
        else if (a == "--aa-count") aa = std::atoi(next());
        else if (a == "--time") time = std::atof(next());
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--stage") stage = next();
        else if (a == "--output") output = next();
        else if (a == "--texture") textures.push_back(next());
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
    if (!stage.empty() && ph_scene_init_stage(scene, stage.c_str())) {   // main.rs:2900-2906
        std::fprintf(stderr, "Scene `%s` has no stage named `%s`\n", scene_path.c_str(), stage.c_str());
        return 1;
    }
    ph_scene_set_time(scene, time, time);
    if (ph_scene_build_program(scene, ctx) || ph_scene_upload_uniforms(scene, ctx)) {
        std::fprintf(stderr, "%s\n", ph_scene_last_error(scene));
        return 1;
    }
    for (const std::string& t : textures) {  // name=file.rgba:WxH
        size_t eq = t.find('='), colon = t.rfind(':'), x = t.rfind('x');
        if (eq == std::string::npos || colon == std::string::npos || x == std::string::npos || x < colon) { std::fprintf(stderr, "bad --texture %s\n", t.c_str()); return 2; }
        int tw = std::atoi(t.substr(colon + 1, x - colon - 1).c_str()), th = std::atoi(t.substr(x + 1).c_str());
        bool tok = false;
        std::string bytes = slurp(t.substr(eq + 1, colon - eq - 1), tok);
        if (!tok || bytes.size() != size_t(tw) * size_t(th) * 4) { std::fprintf(stderr, "texture %s: cannot read / wrong size\n", t.c_str()); return 1; }
        if (pe_set_texture(ctx, t.substr(0, eq).c_str(), reinterpret_cast<const uint8_t*>(bytes.data()), tw, th)) { std::fprintf(stderr, "%s\n", pe_last_error(ctx)); return 1; }
    }
    if (pe_scene_compile(ctx)) { std::fprintf(stderr, "%s\n", pe_last_error(ctx)); return 1; }
    std::vector<uint8_t> px(size_t(width) * size_t(height) * 4);
    ph_frame_params p = {width, height, depth, aa, 0, 0, {0, 0, 0}, 0, 0, 0};
    if (ph_render_frame(scene, ctx, &p, px.data(), 1)) { std::fprintf(stderr, "%s\n", ph_scene_last_error(scene)); return 1; }
    std::ofstream out(output, std::ios::binary);
    if (output.size() > 5 && output.substr(output.size() - 5) == ".rgba") {
        out.write(reinterpret_cast<const char*>(px.data()), std::streamsize(px.size()));
    } else {
        out << "P6\n" << width << " " << height << "\n255\n";
        for (size_t i = 0; i < size_t(width) * size_t(height); i++) out.write(reinterpret_cast<const char*>(&px[4 * i]), 3);
    }
    std::printf("Rendered `%s` to `%s`\n", scene_path.c_str(), output.c_str());
    pe_destroy(ctx);
    ph_scene_free(scene);
    return 0;
}
