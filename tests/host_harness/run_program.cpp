// Runs a generated sm_100a scene program (the text pe_scene_source returns) on the HOST, thread by thread.
// TEST INFRASTRUCTURE (tests/test_program_on_host.py): the product has no CPU path; this exists so that the
// generated program + the hand-written device headers can be compared with the oracle without a GPU, and so
// that generator options (lazy_planes, hoist_planes, specialize_matrices ...) can be proven frame-identical.
//
//   g++ -std=c++20 -O1 -ffp-contract=off -DPROGRAM_FILE='"prog.cu"' run_program.cpp -o run
//   ./run block.bin W H out.f32 [tex0.rgba W0 H0 [tex1.rgba W1 H1 ...]]
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "cuda_on_host.h"
#include PROGRAM_FILE

static std::vector<unsigned char> slurp(const char* path) {
    std::vector<unsigned char> b;
    if (FILE* f = std::fopen(path, "rb")) {
        std::fseek(f, 0, SEEK_END);
        long n = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        b.resize(size_t(n));
        if (std::fread(b.data(), 1, size_t(n), f) != size_t(n)) b.clear();
        std::fclose(f);
    }
    return b;
}

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    const std::vector<unsigned char> block = slurp(argv[1]);
#if PE_WITH_PROBE
    if (std::string(argv[2]) == "probe") {   // ./run block.bin probe ax ay az bx by bz: teleport_external_ray on the host
        if (argc < 9 || block.size() != sizeof(pe::PeConstBlock)) return 2;
        std::memcpy(&PE_C, block.data(), sizeof PE_C);
#if PE_UNIFORMS_SMEM
        std::memcpy(&PE_C_UPLOAD, block.data(), sizeof PE_C_UPLOAD);
#endif
        float out6[6] = {0, 0, 0, 0, 0, 0};
        PeProbe P{float(std::atof(argv[3])), float(std::atof(argv[4])), float(std::atof(argv[5])), float(std::atof(argv[6])),
                  float(std::atof(argv[7])), float(std::atof(argv[8])), out6};
        blockDim = pe_uint3{32, 1, 1};
        gridDim = pe_uint3{1, 1, 1};
        blockIdx = pe_uint3{0, 0, 0};
        threadIdx = pe_uint3{0, 0, 0};
        pe_probe_kernel(P);
        std::printf("%a %a %a %a %a %a\n", out6[0], out6[1], out6[2], out6[3], out6[4], out6[5]);
        return 0;
    }
#endif
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
    if (block.size() != sizeof(pe::PeConstBlock)) {
        std::fprintf(stderr, "uniform block is %zu bytes, program expects %zu\n", block.size(), sizeof(pe::PeConstBlock));
        return 1;
    }
    std::memcpy(&PE_C, block.data(), sizeof PE_C);
    std::vector<std::vector<unsigned char>> textures;
    const int n_tex = int(sizeof(PE_C.tex) / sizeof(PE_C.tex[0]));
    for (int k = 0; k < n_tex; k++) {  // device pointers in the block are meaningless here: bind host texels (or nothing)
        PE_C.tex[k].data = nullptr;
        PE_C.tex[k].w = PE_C.tex[k].h = 0;
        const int a = 5 + 3 * k;
        if (a + 2 < argc) {
            textures.push_back(slurp(argv[a]));
            PE_C.tex[k].data = reinterpret_cast<const uchar4*>(textures.back().data());
            PE_C.tex[k].w = std::atoi(argv[a + 1]);
            PE_C.tex[k].h = std::atoi(argv[a + 2]);
        }
    }
#if PE_UNIFORMS_SMEM
    // the "shared" image is an ordinary global here and the threads run one after another: give the upload symbol the
    // same bytes, so that each thread's share of the block copy rewrites what is already there
    std::memcpy(&PE_C_UPLOAD, &PE_C, sizeof PE_C);
#endif
    std::vector<float4> out(size_t(w) * size_t(h));
    PeLaunch L{};
    L.out = out.data();
    L.bounces = nullptr;
    L.width = w;
    L.height = h;
    L.strip_rows = h;
    L.strip_first = 0;
    L.strip_step = 1;
    L.n_strips = 1;
    L.out_full_frame = 1;
    L.tiles_x = (w + 7) / 8;
    L.tiles_y = (h + 3) / 4;
    L.out_rgba8 = 0;
    L.queue = nullptr;
    L.strip_shift = -1;
    const unsigned rows_per_block = PE_BLOCK_ROWS;
    blockDim = pe_uint3{unsigned(PE_BLOCK_THREADS), 1, 1};
    gridDim = pe_uint3{unsigned((w + PE_BLOCK_W - 1) / PE_BLOCK_W), unsigned((h + rows_per_block - 1) / rows_per_block), 1};
    const pe_uint3 bd = blockDim, gd = gridDim;
#pragma omp parallel for schedule(dynamic, 1)   // (only with -fopenmp: the built-in indices are thread_local)
    for (int by = 0; by < int(gd.y); by++)
        for (unsigned bx = 0; bx < gd.x; bx++)
            for (unsigned t = 0; t < bd.x; t++) {
                blockDim = bd;
                gridDim = gd;
                blockIdx = pe_uint3{bx, unsigned(by), 0};
                threadIdx = pe_uint3{t, 0, 0};
                pe_render_kernel(L);
            }
    FILE* f = std::fopen(argv[4], "wb");
    if (!f) return 1;
    std::fwrite(out.data(), sizeof(float4), out.size(), f);
    std::fclose(f);
    return 0;
}
