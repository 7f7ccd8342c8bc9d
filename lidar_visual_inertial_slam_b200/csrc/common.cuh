// common.cuh -- shared device/host helpers for liblvreg (sm_100a only).
//
// Numerics contract: every kernel in this library is compiled with -fmad=false, so each
// fp32/fp64 operation rounds once, exactly like the reference's x86-64 build (-O3, no -march:
// lidar_odometry/CMakeLists.txt:9-11).  Division and sqrt are the IEEE-correct variants (nvcc
// defaults -prec-div=true -prec-sqrt=true; --use_fast_math is never used).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace lvreg {

struct Affine {            // row-major 3x4, as produced by pcl::getTransformation (MO:399-407)
    float m[12];
};

// out = R p + t with the reference's left-to-right evaluation (MO:341-343, MO:360-362)
__device__ __forceinline__ float3 apply_affine(const Affine& T, float x, float y, float z) {
    float3 o;
    o.x = T.m[0] * x + T.m[1] * y + T.m[2] * z + T.m[3];
    o.y = T.m[4] * x + T.m[5] * y + T.m[6] * z + T.m[7];
    o.z = T.m[8] * x + T.m[9] * y + T.m[10] * z + T.m[11];
    return o;
}

// L2_Simple: ((dx*dx) + dy*dy) + dz*dz, one rounding per op (FLANN dist.h semantics)
__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float r = dx * dx;
    r += dy * dy;
    r += dz * dz;
    return r;
}

// streaming 128-bit load that does not pollute L1 (data touched once per kernel)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// order-preserving float <-> uint mapping for atomicMin/Max on floats
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    union { uint32_t u; float f; } c;
    c.u = v;
    return c.f;
#endif
}

constexpr int kNumSMs = 148;     // B200

}  // namespace lvreg
