/* Plain C11 caller of the portal_b200 C ABI: what a non-C++ host (the reference's Rust via FFI)
 * does, without any of this repo's Python.  Loads a `.ron` scene with the host front-end, hands the scene program
 * to the renderer, compiles it for sm_100a, renders one RGBA8 frame into host memory and writes a PPM.
 *
 *   gcc -std=c11 -I include examples/render_frame.c -L portal_b200 -lportal_b200 -Wl,-rpath,$PWD/portal_b200 -o render_frame
 *   ./render_frame tests/fixtures/two_spheres.ron out.ppm [width height depth device]
 *
 * With device = -1 (no GPU) everything up to and including compilation works and the render call fails loudly:
 * there is no CPU rendering path. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "portal_b200_host.h"

static char* slurp(const char* path, size_t* len) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* buf = (char*)malloc((size_t)n + 1);
    if (buf && fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); buf = NULL; }
    fclose(f);
    if (buf) { buf[n] = 0; *len = (size_t)n; }
    return buf;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s scene.ron out.ppm [width height depth device]\n", argv[0]);
        return 2;
    }
    const int width = argc > 3 ? atoi(argv[3]) : 320, height = argc > 4 ? atoi(argv[4]) : 180;
    const int depth = argc > 5 ? atoi(argv[5]) : 20, device = argc > 6 ? atoi(argv[6]) : 0;
    size_t len = 0;
    char* text = slurp(argv[1], &len);
    if (!text) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    char err[1024] = {0};
    ph_scene* scene = ph_scene_load_ron(text, len, err, sizeof err);
    free(text);
    if (!scene) { fprintf(stderr, "scene: %s\n", err); return 1; }
    pe_ctx* ctx = pe_create(device);
    if (!ctx) { fprintf(stderr, "pe_create: %s\n", pe_last_error(NULL)); return 1; }
    if (ph_scene_build_program(scene, ctx) || ph_scene_upload_uniforms(scene, ctx)) {
        fprintf(stderr, "scene program: %s\n", ph_scene_last_error(scene));
        return 1;
    }
    if (pe_scene_compile(ctx)) { fprintf(stderr, "compile: %s\n", pe_last_error(ctx)); return 1; }
    const void* cubin = NULL;
    size_t cubin_size = 0;
    pe_scene_cubin(ctx, &cubin, &cubin_size);
    printf("compiled %zu bytes of sm_100a code for %d objects\n", cubin_size, ph_scene_count(scene, 0));

    unsigned char* pixels = (unsigned char*)malloc((size_t)width * (size_t)height * 4);
    ph_frame_params p;
    memset(&p, 0, sizeof p);
    p.width = width; p.height = height; p.depth = depth; p.aa_count = 1;
    if (ph_render_frame(scene, ctx, &p, pixels, 1)) {
        fprintf(stderr, "render: %s\n", ph_scene_last_error(scene));
        return 3;
    }
    FILE* out = fopen(argv[2], "wb");
    if (!out) return 1;
    fprintf(out, "P6\n%d %d\n255\n", width, height);
    for (size_t i = 0; i < (size_t)width * (size_t)height; i++) fwrite(pixels + 4 * i, 1, 3, out);
    fclose(out);
    printf("wrote %s (%dx%d, depth %d, %llu kernel launch(es))\n", argv[2], width, height, depth, (unsigned long long)pe_launch_count(ctx));
    free(pixels);
    pe_destroy(ctx);
    ph_scene_free(scene);
    return 0;
}
