// Static (nvcc-compiled, sm_100a) helper kernels of portal_b200: the HBM-bound passes around the
// ray loop.  All are pure streaming kernels: 16-byte vector accesses, grid sized to the SM count,
// grid-stride loops.
//
//   pe_k_quantize_rgba8     float RGBA -> RGBA8: what the reference's RGBA8 render target does to
//                           FragColor before get_texture_data() (/root/reference/src/main.rs:2939-2943)
//   pe_k_deinterleave       rank-major gathered strips -> row-major frame
//   pe_k_average_rgba8      motion-blur average in gamma-2 space = average_images
//                           (/root/reference/src/main.rs:640-722): square, integer mean, rounded sqrt
#include <cuda_runtime.h>
#include <stdint.h>

#include "pe_kernels.h"

namespace {

__device__ __forceinline__ unsigned quant8(float v) {
    // GL unorm8 conversion: clamp to [0,1], scale by 255, round to nearest
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    return (unsigned)__float2int_rn(v * 255.0f);
}

__global__ void __launch_bounds__(256) pe_k_quantize_rgba8(const float4* __restrict__ in, uchar4* __restrict__ out, size_t n) {
    size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        float4 p = in[i];
        // NaN -> 0 (fmaxf(NaN, 0) == 0), the behaviour of GL's float->unorm conversion on NVIDIA
        out[i] = make_uchar4((unsigned char)quant8(p.x), (unsigned char)quant8(p.y), (unsigned char)quant8(p.z),
                             (unsigned char)quant8(p.w));
    }
}

// One thread per float4 pixel.  gathered: [rank][strips_per_rank][strip_rows][width]
__global__ void __launch_bounds__(256) pe_k_deinterleave(const float4* __restrict__ gathered, float4* __restrict__ frame,
                                                         int width, int height, int strip_rows, int n_ranks,
                                                         int strips_per_rank) {
    size_t n = size_t(width) * size_t(height);
    size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        int y = int(i / size_t(width));
        int x = int(i - size_t(y) * size_t(width));
        int gstrip = y / strip_rows;
        int rank = gstrip % n_ranks;
        int lstrip = gstrip / n_ranks;
        size_t src = ((size_t(rank) * strips_per_rank + lstrip) * strip_rows + size_t(y - gstrip * strip_rows)) * size_t(width) + x;
        frame[i] = gathered[src];
    }
}

#define PE_MAX_AVG_FRAMES 64
struct FramePtrs {
    const uchar4* p[PE_MAX_AVG_FRAMES];
};

__device__ __forceinline__ unsigned l_to_s(unsigned l) {
    // L_TO_S[i] = ((i as f32).sqrt() + 0.5) as u8   (main.rs:645-651)
    return (unsigned)(sqrtf((float)l) + 0.5f);
}

__global__ void __launch_bounds__(256) pe_k_average_rgba8(FramePtrs frames, int n_frames, uchar4* __restrict__ out, size_t n) {
    size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned r = 0, g = 0, b = 0;
        for (int f = 0; f < n_frames; f++) {
            uchar4 p = frames.p[f][i];
            r += unsigned(p.x) * unsigned(p.x);
            g += unsigned(p.y) * unsigned(p.y);
            b += unsigned(p.z) * unsigned(p.z);
        }
        unsigned nn = (unsigned)n_frames;
        out[i] = make_uchar4((unsigned char)l_to_s(r / nn), (unsigned char)l_to_s(g / nn), (unsigned char)l_to_s(b / nn), 255);
    }
}

// Cross-GPU signal: after everything earlier in the stream (the render kernel's remote stores
// included) one thread publishes `value` to up to 8 flag words, which may live in peer memory.
struct SignalPtrs {
    unsigned int* p[8];
};
__global__ void pe_k_signal(SignalPtrs ptrs, int n, unsigned int value) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        __threadfence_system();
        for (int i = 0; i < n; i++) *((volatile unsigned int*)ptrs.p[i]) = value;
        __threadfence_system();
    }
}

int grid_for(size_t n, int sms) {
    size_t blocks = (n + 255) / 256;
    size_t cap = size_t(sms) * 8;  // 8 x 256-thread blocks per SM = full occupancy
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return int(blocks);
}

}  // namespace

namespace pe_host {

int launch_quantize_rgba8(const void* in, void* out, size_t n, int sms, cudaStream_t s) {
    pe_k_quantize_rgba8<<<grid_for(n, sms), 256, 0, s>>>((const float4*)in, (uchar4*)out, n);
    return (int)cudaGetLastError();
}

int launch_deinterleave(const void* gathered, void* frame, int width, int height, int strip_rows, int n_ranks,
                        int strips_per_rank, int sms, cudaStream_t s) {
    size_t n = size_t(width) * size_t(height);
    pe_k_deinterleave<<<grid_for(n, sms), 256, 0, s>>>((const float4*)gathered, (float4*)frame, width, height, strip_rows,
                                                     n_ranks, strips_per_rank);
    return (int)cudaGetLastError();
}

int launch_signal(void* const* ptrs, int n, unsigned int value, cudaStream_t s) {
    if (n < 1 || n > 8) return (int)cudaErrorInvalidValue;
    SignalPtrs sp;
    for (int i = 0; i < n; i++) sp.p[i] = (unsigned int*)ptrs[i];
    pe_k_signal<<<1, 32, 0, s>>>(sp, n, value);
    return (int)cudaGetLastError();
}

int launch_average_rgba8(const void* const* frames, int n_frames, void* out, size_t n, int sms, cudaStream_t s) {
    if (n_frames < 1 || n_frames > PE_MAX_AVG_FRAMES) return (int)cudaErrorInvalidValue;
    FramePtrs fp;
    for (int i = 0; i < n_frames; i++) fp.p[i] = (const uchar4*)frames[i];
    pe_k_average_rgba8<<<grid_for(n, sms), 256, 0, s>>>(fp, n_frames, (uchar4*)out, n);
    return (int)cudaGetLastError();
}

}  // namespace pe_host
