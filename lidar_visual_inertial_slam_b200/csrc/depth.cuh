// depth.cuh -- LiDAR depth for tracked visual features:
//   FN: = feature_tracker/src/feature_tracker_node.cpp:273-375  lidar_callback (stack + 0.2 m VoxelGrid)
//   FT: = feature_tracker/src/feature_tracker.h:150-283         DepthRegister::get_depth
//
//   ViewFlagIn / ViewCompactOut   camera-view filter + transform into the odometry frame, order kept
//                                 (FN:313-331), as one stream compaction
//   depth_bin_kernel              depth cloud into the camera frame, 360 x 360 range image: per
//                                 0.5 deg bin the closest point, the earliest among equals (FT:170-196)
//                                 = atomicMin over (distance bits << 32 | index)
//   BinFlagIn / BinCompactOut     the occupied bins in row-major order (FT:198-207) and their
//                                 unit-sphere projection with the range in .w (FT:211-222)
//   depth_feature_kernel          one block per feature: exact 3-NN on the unit sphere ((d2, index)
//                                 order), ray / plane intersection and the clamps of FT:236-268
//
// atan2 is evaluated in double and rounded to float (the oracle does the same; the reference's glibc
// atan2f may differ by one ulp for a few inputs -- stated in DESIGN.md).
#pragma once

#include "common.cuh"

namespace lvreg {

__device__ __forceinline__ bool in_camera_view(float x, float y, float z) {
    // p.x >= 0 && abs(p.y / p.x) <= 10 && abs(p.z / p.x) <= 10   (NaN and inf compare false)
    return x >= 0.f && fabsf(y / x) <= 10.f && fabsf(z / x) <= 10.f;
}

struct ViewFlagIn {
    const float4* pts;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        const float4 p = pts[i];
        return in_camera_view(p.x, p.y, p.z) ? 1u : 0u;
    }
    __device__ __forceinline__ void load_vec(uint32_t i, uint32_t (&v)[8]) const {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (*this)(i + k);
    }
};
struct ViewCompactOut {
    const float4* pts;
    Affine T;                // transNow, FN:331
    float4* out;
    __device__ __forceinline__ void operator()(uint32_t i, uint32_t flag, uint32_t pre) const {
        if (flag) {
            const float4 p = pts[i];
            const float3 q = apply_affine(T, p.x, p.y, p.z);
            out[pre] = make_float4(q.x, q.y, q.z, p.w);
        }
    }
    __device__ __forceinline__ void store_vec(uint32_t i, const uint32_t (&v)[8], uint32_t pre) const {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            (*this)(i + k, v[k], pre);
            pre += v[k];
        }
    }
};

__device__ __forceinline__ float atan2_rounded(float a, float b) { return (float)atan2((double)a, (double)b); }

constexpr unsigned long long kBinEmpty = ~0ull;

__global__ void __launch_bounds__(256) depth_bin_kernel(const float4* __restrict__ cloud, uint32_t m, Affine Tinv,
                                                        int num_bins, unsigned long long* __restrict__ bins) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    const float4 p0 = cloud[i];
    const float3 p = apply_affine(Tinv, p0.x, p0.y, p0.z);
    if (p.x < 0.f || fabsf(p.y / p.x) > 10.f || fabsf(p.z / p.x) > 10.f) return;
    const float bin_res = (float)(180.0 / (double)(float)num_bins);
    const float row_angle = (float)((double)atan2_rounded(p.z, sqrtf(p.x * p.x + p.y * p.y)) * 180.0 / 3.14159265358979323846 + 90.0);
    const int row_id = (int)roundf(row_angle / bin_res);
    const float col_angle = (float)((double)atan2_rounded(p.x, p.y) * 180.0 / 3.14159265358979323846);
    const int col_id = (int)roundf(col_angle / bin_res);
    if (row_id < 0 || row_id >= num_bins || col_id < 0 || col_id >= num_bins) return;
    const float dist = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
    if (!(dist < 3.402823466e+38f)) return;          // never closer than the initial FLT_MAX
    // "if (dist < rangeImage(row, col))" over the points in order == minimum over (dist, index)
    const unsigned long long key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)i;
    atomicMin(&bins[(size_t)row_id * num_bins + col_id], key);
}

struct BinFlagIn {
    const unsigned long long* bins;
    __device__ __forceinline__ uint32_t operator()(uint32_t b) const {
        return bins[b] != kBinEmpty ? 1u : 0u;          // rangeImage != FLT_MAX
    }
    __device__ __forceinline__ void load_vec(uint32_t b, uint32_t (&v)[8]) const {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (*this)(b + k);
    }
};
struct BinCompactOut {
    const unsigned long long* bins;
    const float4* cloud;
    Affine Tinv;
    float4* local;           // depth_cloud_local after the range-image filter
    float4* unit;            // depth_cloud_unit_sphere, .w = range
    __device__ __forceinline__ void operator()(uint32_t b, uint32_t flag, uint32_t pre) const {
        if (!flag) return;
        const uint32_t i = (uint32_t)bins[b];
        const float4 p0 = cloud[i];
        const float3 p = apply_affine(Tinv, p0.x, p0.y, p0.z);
        local[pre] = make_float4(p.x, p.y, p.z, p0.w);
        const float r = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
        unit[pre] = make_float4(p.x / r, p.y / r, p.z / r, r);
    }
    __device__ __forceinline__ void store_vec(uint32_t b, const uint32_t (&v)[8], uint32_t pre) const {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            (*this)(b + k, v[k], pre);
            pre += v[k];
        }
    }
};

__device__ __forceinline__ void top3_insert(unsigned long long (&t)[3], unsigned long long key) {
    if (key < t[2]) {
        t[2] = key;
        if (t[2] < t[1]) { unsigned long long x = t[1]; t[1] = t[2]; t[2] = x; }
        if (t[1] < t[0]) { unsigned long long x = t[0]; t[0] = t[1]; t[1] = x; }
    }
}

constexpr int kDepthThreads = 256;

// one block per feature.  feat: n x 3 normalised image coordinates (z = 1).  depth_out[i] = -1 when
// no depth; feat3d_out (optional): features_3d_sphere as published (FT:270).
__global__ void __launch_bounds__(kDepthThreads) depth_feature_kernel(const float* __restrict__ feat, uint32_t n,
                                                                      const float4* __restrict__ unit,
                                                                      const uint32_t* __restrict__ n_unit_p,
                                                                      float thr, float* __restrict__ depth_out,
                                                                      float4* __restrict__ feat3d_out) {
    const uint32_t f = blockIdx.x;
    if (f >= n) return;
    const uint32_t nu = *n_unit_p;
    // Eigen::Vector3f::normalize(), then the ROS axis convention (FT:157-165)
    const float x = feat[3 * f], y = feat[3 * f + 1], z = feat[3 * f + 2];
    const float nrm = sqrtf(x * x + y * y + z * z);
    float vx = z / nrm, vy = -(x / nrm), vz = -(y / nrm);
    float inten = -1.f;
    unsigned long long t[3] = {~0ull, ~0ull, ~0ull};
    if (nu >= 10) {                                                              // FT:224-225
        for (uint32_t j = threadIdx.x; j < nu; j += kDepthThreads) {
            const float4 u = __ldg(unit + j);
            const float d = sqdist(vx, vy, vz, u.x, u.y, u.z);
            top3_insert(t, ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)j);
        }
    }
    __shared__ unsigned long long cand[kDepthThreads * 3];
    cand[threadIdx.x * 3 + 0] = t[0];
    cand[threadIdx.x * 3 + 1] = t[1];
    cand[threadIdx.x * 3 + 2] = t[2];
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int k = 3; k < kDepthThreads * 3; ++k) top3_insert(t, cand[k]);
    if (nu >= 10 && t[2] != ~0ull && __uint_as_float((uint32_t)(t[2] >> 32)) < thr) {
        const float4 u0 = unit[(uint32_t)t[0]], u1 = unit[(uint32_t)t[1]], u2 = unit[(uint32_t)t[2]];
        const float r1 = u0.w, r2 = u1.w, r3 = u2.w;
        const float Ax = u0.x * r1, Ay = u0.y * r1, Az = u0.z * r1;
        const float Bx = u1.x * r2, By = u1.y * r2, Bz = u1.z * r2;
        const float Cx = u2.x * r3, Cy = u2.y * r3, Cz = u2.z * r3;
        const float abx = Ax - Bx, aby = Ay - By, abz = Az - Bz;
        const float bcx = Bx - Cx, bcy = By - Cy, bcz = Bz - Cz;
        const float Nx = aby * bcz - abz * bcy, Ny = abz * bcx - abx * bcz, Nz = abx * bcy - aby * bcx;
        float s = (Nx * Ax + Ny * Ay + Nz * Az) / (Nx * vx + Ny * vy + Nz * vz);
        const float min_depth = fminf(r1, fminf(r2, r3));
        const float max_depth = fmaxf(r1, fmaxf(r2, r3));
        bool keep = true;
        if (max_depth - min_depth > 2.f || (double)s <= 0.5) keep = false;
        else if (s - max_depth > 0.f) s = max_depth;
        else if (s - min_depth < 0.f) s = min_depth;
        if (keep) {
            vx *= s; vy *= s; vz *= s;
            inten = vx;                      // depth along the camera's optical axis (lidar x = camera z)
        }
    }
    depth_out[f] = (double)inten > 3.0 ? inten : -1.f;
    if (feat3d_out) feat3d_out[f] = make_float4(vx, vy, vz, inten);
}

}  // namespace lvreg
