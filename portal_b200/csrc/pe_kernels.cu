// Static (nvcc-compiled, sm_100a) helper kernels of portal_b200: the HBM-bound passes around the
// ray loop.  All are pure streaming kernels: 16-byte vector accesses with several independent loads in
// flight per thread, streaming cache hints (every byte is touched once), grid sized to the SM count.
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
__device__ __forceinline__ unsigned pack8(float4 p) {
    // NaN -> 0 (fmaxf(NaN, 0) == 0), the behaviour of GL's float->unorm conversion on NVIDIA
    return quant8(p.x) | (quant8(p.y) << 8) | (quant8(p.z) << 16) | (quant8(p.w) << 24);
}

// Streaming loads / stores that do not pollute L1 and are evicted first from L2: every byte here is touched once.
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ uint4 ld_stream(const uint4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(uint4* p, uint4 v) { __stcs(p, v); }

// float RGBA -> RGBA8.  A thread owns 4 CONSECUTIVE pixels: four 16-byte loads (all issued before the first use, so 64 B
// per thread are in flight: 2048 threads/SM x 64 B = 128 KB per SM, enough to cover the HBM latency-bandwidth product)
// and ONE 16-byte store.  A warp's four load instructions together cover 2 KB contiguous; each 32-byte sector is
// fetched once (the second half of a sector is an L1 hit of the neighbouring load instruction).
#define PE_Q_UNROLL 2
__global__ void __launch_bounds__(256) pe_k_quantize_rgba8(const float4* __restrict__ in, uchar4* __restrict__ out, size_t n) {
    const size_t n4 = n >> 2;                                   // groups of 4 pixels
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; g + (PE_Q_UNROLL - 1) * stride < n4; g += PE_Q_UNROLL * stride) {
        float4 p[PE_Q_UNROLL][4];
#pragma unroll
        for (int u = 0; u < PE_Q_UNROLL; u++)
#pragma unroll
            for (int k = 0; k < 4; k++) p[u][k] = ld_stream(in + 4 * (g + u * stride) + k);
#pragma unroll
        for (int u = 0; u < PE_Q_UNROLL; u++)
            st_stream(reinterpret_cast<uint4*>(out) + (g + u * stride),
                      make_uint4(pack8(p[u][0]), pack8(p[u][1]), pack8(p[u][2]), pack8(p[u][3])));
    }
    for (; g < n4; g += stride) {
        float4 p[4];
#pragma unroll
        for (int k = 0; k < 4; k++) p[k] = ld_stream(in + 4 * g + k);
        st_stream(reinterpret_cast<uint4*>(out) + g, make_uint4(pack8(p[0]), pack8(p[1]), pack8(p[2]), pack8(p[3])));
    }
    // ragged tail (n not a multiple of 4): at most 3 pixels
    const size_t i = (n4 << 2) + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) reinterpret_cast<unsigned*>(out)[i] = pack8(in[i]);
}

// Rank-major gathered strips -> row-major frame.  Work unit = one 16-byte pixel; the index map is per ROW (a row of the
// frame is contiguous in both layouts), so the kernel walks rows with blockIdx.y and a thread copies 4 pixels of the
// row, all four loads in flight before the stores.  gathered: [rank][strips_per_rank][strip_rows][width]
__global__ void __launch_bounds__(256) pe_k_deinterleave(const float4* __restrict__ gathered, float4* __restrict__ frame,
                                                         int width, int height, int strip_rows, int n_ranks,
                                                         int strips_per_rank) {
    auto row_src = [&](int y) {
        const int gstrip = y / strip_rows;
        const int rank = gstrip % n_ranks;
        const int lstrip = gstrip / n_ranks;
        return gathered + ((size_t(rank) * strips_per_rank + lstrip) * strip_rows + size_t(y - gstrip * strip_rows)) * size_t(width);
    };
    const int step = blockDim.x * gridDim.x;
    // two rows per pass: eight 16-byte loads in flight per thread before the first store
    int y = blockIdx.y;
    for (; y + int(gridDim.y) < height; y += 2 * gridDim.y) {
        const float4 *s0 = row_src(y), *s1 = row_src(y + gridDim.y);
        float4 *d0 = frame + size_t(y) * size_t(width), *d1 = frame + size_t(y + gridDim.y) * size_t(width);
        int x = blockIdx.x * blockDim.x + threadIdx.x;
        for (; x + 3 * step < width; x += 4 * step) {
            float4 a = ld_stream(s0 + x), b = ld_stream(s0 + x + step), c = ld_stream(s0 + x + 2 * step), d = ld_stream(s0 + x + 3 * step);
            float4 e = ld_stream(s1 + x), f = ld_stream(s1 + x + step), g = ld_stream(s1 + x + 2 * step), h = ld_stream(s1 + x + 3 * step);
            st_stream(d0 + x, a); st_stream(d0 + x + step, b); st_stream(d0 + x + 2 * step, c); st_stream(d0 + x + 3 * step, d);
            st_stream(d1 + x, e); st_stream(d1 + x + step, f); st_stream(d1 + x + 2 * step, g); st_stream(d1 + x + 3 * step, h);
        }
        for (; x < width; x += step) { st_stream(d0 + x, ld_stream(s0 + x)); st_stream(d1 + x, ld_stream(s1 + x)); }
    }
    for (; y < height; y += gridDim.y) {
        const float4* src = row_src(y);
        float4* dst = frame + size_t(y) * size_t(width);
        for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < width; x += step) st_stream(dst + x, ld_stream(src + x));
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
__device__ __forceinline__ void acc_sq(unsigned w, unsigned& r, unsigned& g, unsigned& b) {
    const unsigned x = w & 0xffu, y = (w >> 8) & 0xffu, z = (w >> 16) & 0xffu;
    r += x * x; g += y * y; b += z * z;
}

// Motion-blur mean in gamma-2 space.  A thread owns 4 consecutive pixels (one 16-byte load per sub-frame, one 16-byte
// store); the loads of 4 sub-frames are issued together.
// `magic` = ceil(2^32 / n_frames): for sums below 2^23 (64 frames x 255^2 < 2^22) __umulhi(sum, magic) == sum / n_frames
// exactly (the error term sum * (magic - 2^32 / n) / 2^32 < 2^-9 cannot carry past the next integer: the fractional part of
// sum / n is at most 1 - 1/64) -- one IMAD.HI instead of a ~20-instruction unsigned division per channel.
__device__ __forceinline__ unsigned div_n(unsigned sum, unsigned magic, unsigned nn) { return nn == 1u ? sum : __umulhi(sum, magic); }

__global__ void __launch_bounds__(256) pe_k_average_rgba8(FramePtrs frames, int n_frames, unsigned magic, uchar4* __restrict__ out, size_t n) {
    const size_t n4 = n >> 2;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    const unsigned nn = (unsigned)n_frames;
    for (size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x; g < n4; g += stride) {
        unsigned r[4] = {0, 0, 0, 0}, gg[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
        int f = 0;
        for (; f + 4 <= n_frames; f += 4) {
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = ld_stream(reinterpret_cast<const uint4*>(frames.p[f + k]) + g);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                acc_sq(v[k].x, r[0], gg[0], b[0]); acc_sq(v[k].y, r[1], gg[1], b[1]);
                acc_sq(v[k].z, r[2], gg[2], b[2]); acc_sq(v[k].w, r[3], gg[3], b[3]);
            }
        }
        for (; f < n_frames; f++) {
            const uint4 v = ld_stream(reinterpret_cast<const uint4*>(frames.p[f]) + g);
            acc_sq(v.x, r[0], gg[0], b[0]); acc_sq(v.y, r[1], gg[1], b[1]); acc_sq(v.z, r[2], gg[2], b[2]); acc_sq(v.w, r[3], gg[3], b[3]);
        }
        unsigned o[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            o[k] = l_to_s(div_n(r[k], magic, nn)) | (l_to_s(div_n(gg[k], magic, nn)) << 8) | (l_to_s(div_n(b[k], magic, nn)) << 16) | 0xff000000u;
        st_stream(reinterpret_cast<uint4*>(out) + g, make_uint4(o[0], o[1], o[2], o[3]));
    }
    const size_t i = (n4 << 2) + size_t(blockIdx.x) * blockDim.x + threadIdx.x;   // ragged tail: at most 3 pixels
    if (i < n) {
        unsigned r = 0, g = 0, b = 0;
        for (int f = 0; f < n_frames; f++) acc_sq(reinterpret_cast<const unsigned*>(frames.p[f])[i], r, g, b);
        reinterpret_cast<unsigned*>(out)[i] = l_to_s(r / nn) | (l_to_s(g / nn) << 8) | (l_to_s(b / nn) << 16) | 0xff000000u;
    }
}

// Scalar forms for pixel buffers that are not 16-byte aligned (a caller's sub-rectangle): one pixel per thread.
__global__ void __launch_bounds__(256) pe_k_quantize_rgba8_px(const float4* __restrict__ in, unsigned* __restrict__ out, size_t n) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = pack8(in[i]);
}
__global__ void __launch_bounds__(256) pe_k_average_rgba8_px(FramePtrs frames, int n_frames, unsigned* __restrict__ out, size_t n) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    const unsigned nn = (unsigned)n_frames;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned r = 0, g = 0, b = 0;
        for (int f = 0; f < n_frames; f++) acc_sq(reinterpret_cast<const unsigned*>(frames.p[f])[i], r, g, b);
        out[i] = l_to_s(r / nn) | (l_to_s(g / nn) << 8) | (l_to_s(b / nn) << 16) | 0xff000000u;
    }
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

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

// How many 16-byte words of two frames differ (bit comparison; pe_autotune's guard: a candidate variant must give the frame
// the first candidate gave).  One atomicAdd per warp that saw a difference.
__global__ void __launch_bounds__(256) pe_k_count_diff(const uint4* __restrict__ a, const uint4* __restrict__ b, size_t n16,
                                                       unsigned* __restrict__ count) {
    unsigned mine = 0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += size_t(gridDim.x) * blockDim.x) {
        const uint4 x = a[i], y = b[i];
        mine += ((x.x ^ y.x) | (x.y ^ y.y) | (x.z ^ y.z) | (x.w ^ y.w)) != 0u;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(count, mine);
}


}  // namespace

namespace pe_host {

int launch_quantize_rgba8(const void* in, void* out, size_t n, int sms, cudaStream_t s) {
    if (aligned16(out)) pe_k_quantize_rgba8<<<grid_for((n + 3) / 4, sms), 256, 0, s>>>((const float4*)in, (uchar4*)out, n);
    else pe_k_quantize_rgba8_px<<<grid_for(n, sms), 256, 0, s>>>((const float4*)in, (unsigned*)out, n);
    return (int)cudaGetLastError();
}

int launch_deinterleave(const void* gathered, void* frame, int width, int height, int strip_rows, int n_ranks,
                        int strips_per_rank, int sms, cudaStream_t s) {
    // rows on blockIdx.y; one block column when a row fits 256 threads x 4 pixels, else enough columns for one pass
    const unsigned gx = unsigned((width + 1023) / 1024);
    unsigned gy = unsigned(sms) * 8u / gx;
    if (gy > unsigned(height)) gy = unsigned(height);
    if (gy < 1) gy = 1;
    pe_k_deinterleave<<<dim3(gx, gy), 256, 0, s>>>((const float4*)gathered, (float4*)frame, width, height, strip_rows,
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
    bool vec = aligned16(out);
    for (int i = 0; i < n_frames; i++) vec = vec && aligned16(frames[i]);
    const unsigned magic = n_frames > 1 ? unsigned(((1ull << 32) + unsigned(n_frames) - 1) / unsigned(n_frames)) : 0u;
    if (vec) pe_k_average_rgba8<<<grid_for((n + 3) / 4, sms), 256, 0, s>>>(fp, n_frames, magic, (uchar4*)out, n);
    else pe_k_average_rgba8_px<<<grid_for(n, sms), 256, 0, s>>>(fp, n_frames, (unsigned*)out, n);
    return (int)cudaGetLastError();
}

int launch_count_diff(const void* a, const void* b, size_t n16, unsigned* count, int sms, cudaStream_t s) {
    pe_k_count_diff<<<grid_for(n16, sms), 256, 0, s>>>((const uint4*)a, (const uint4*)b, n16, count);
    return (int)cudaGetLastError();
}

}  // namespace pe_host
