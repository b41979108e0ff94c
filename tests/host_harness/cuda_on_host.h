// Just enough of the CUDA execution model to run a generated scene program (pe_scene_source) as ordinary C++:
// one OS thread plays every CUDA thread of the grid in turn (the simple scheduler's threads never communicate).
// TEST INFRASTRUCTURE: included only by tests/host_harness/run_program.cpp.
#pragma once
#include <math.h>

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __forceinline__ inline
#define __noinline__
#define __constant__
#define __global__
#define __device__
#define __host__
#define __launch_bounds__(...)
#define __restrict__
#define __shared__
#define __align__(n) alignas(n)
static inline void __syncthreads() {}

struct uchar4 { unsigned char x, y, z, w; };
struct float4 { float x, y, z, w; };
struct pe_uint3 { unsigned x, y, z; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
static thread_local pe_uint3 threadIdx, blockIdx, blockDim, gridDim;

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
// truncating, saturating like cvt.rzi.s32.f32 (NaN -> 0)
static inline int __float2int_rz(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return -2147483647 - 1;
    return int(v);
}
// round to nearest even, saturating like cvt.rni.s32.f32 (NaN -> 0)
static inline int __float2int_rn(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return -2147483647 - 1;
    return int(::rintf(v));
}

// g++ resolves the non-dependent `a * mat4(b)` inside pe_glsl.cuh's smat4 operator templates when it parses them
// (nvcc / NVRTC defer it): declare the overload they mean ahead of its definition.
namespace pe { struct mat4; inline mat4 operator*(const mat4& a, const mat4& b); }
