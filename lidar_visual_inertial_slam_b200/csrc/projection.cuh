// projection.cuh -- the LiDAR front end's per-point work on the device:
//   IP:495-526 findRotation, IP:538-569 deskewPoint, IP:571-623 projectPointCloud, IP:625-647 cloudExtraction
//   (IP: = lidar_odometry/src/imageProjection.cpp)
//
// The reference walks the scan sequentially; the order dependence is resolved like this:
//   * Livox column index = how many earlier points of the same ring passed the range / ring /
//     down-sample filters  ->  one stable radix-sort pass by ring, column = sorted position - ring start
//   * "first point to reach a (row, column) cell wins"  ->  atomicMin of the point index per cell
//   * transStartInverse comes from the first point that reaches deskewPoint  ->  the minimum index
//     over all cell candidates (its cell is necessarily still empty)
//   * cloudExtraction's row-major walk  ->  one scan over the cells that also emits the ring starts
// Trigonometry: sin / cos / atan2 in double, rounded to float (see fit.cuh, depth.cuh).
#pragma once

#include "common.cuh"
#include "fit.cuh"

namespace lvreg {

struct ProjParams {
    int n_scan, horizon, downsample_rate, sensor;
    float min_range, max_range;
    int deskew, imu_pointer_cur;
    double time_scan_cur;
    const double* imu_time;      // device, imu_pointer_cur + 1 entries
    const double* imu_rx;
    const double* imu_ry;
    const double* imu_rz;
};

struct RawLayout {               // AoS laserCloudIn
    const uint8_t* data;
    uint32_t stride, intensity_off, ring_off, time_off;
};

constexpr uint32_t kProjNone = 0xffffffffu;

// filters of IP:584-594; key = ring for the points that pass (kept for the Livox ranking), 255 otherwise.
// For Velodyne / Ouster the column follows from the point itself (IP:597-602).
__global__ void __launch_bounds__(256) proj_classify_kernel(RawLayout in, uint32_t n, ProjParams P,
                                                            float4* __restrict__ pts, float* __restrict__ range_out,
                                                            uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                            int32_t* __restrict__ col_out, uint32_t* __restrict__ ring_count) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = in.data + (size_t)i * in.stride;
    const float x = *reinterpret_cast<const float*>(p), y = *reinterpret_cast<const float*>(p + 4),
                z = *reinterpret_cast<const float*>(p + 8);
    const float inten = *reinterpret_cast<const float*>(p + in.intensity_off);
    const int row = (int)*reinterpret_cast<const uint16_t*>(p + in.ring_off);
    pts[i] = make_float4(x, y, z, inten);
    const float range = sqrtf(x * x + y * y + z * z);
    range_out[i] = range;
    bool ok = !(range < P.min_range || range > P.max_range);
    ok = ok && row >= 0 && row < P.n_scan;
    ok = ok && (row % P.downsample_rate == 0);
    int col = -1;
    if (ok && P.sensor != 2) {
        const float horizonAngle = (float)((double)((float)atan2((double)x, (double)y) * 180.f) / 3.14159265358979323846);
        const float ang_res_x = (float)(360.0 / (double)(float)P.horizon);
        col = (int)(-round(((double)horizonAngle - 90.0) / (double)ang_res_x) + (double)(P.horizon / 2));
        if (col >= P.horizon) col -= P.horizon;
    }
    col_out[i] = col;
    keys[i] = ok ? (uint32_t)row : 255u;
    vals[i] = i;
    if (ok && P.sensor == 2) atomicAdd(&ring_count[row], 1u);
}

// Livox: exclusive prefix of the per-ring counts (n_scan <= 255), one thread
__global__ void proj_ring_starts_kernel(const uint32_t* __restrict__ ring_count, int n_scan, uint32_t* __restrict__ ring_start) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t s = 0;
    for (int r = 0; r < n_scan; ++r) { ring_start[r] = s; s += ring_count[r]; }
}

// Livox: after the stable sort by ring, sorted position - ring start = columnIdnCountVec at that point
__global__ void __launch_bounds__(256) proj_livox_columns_kernel(const uint32_t* __restrict__ sorted_keys,
                                                                 const uint32_t* __restrict__ sorted_vals, uint32_t n,
                                                                 const uint32_t* __restrict__ ring_start,
                                                                 int32_t* __restrict__ col_out) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const uint32_t r = sorted_keys[j];
    if (r == 255u) return;
    col_out[sorted_vals[j]] = (int32_t)(j - ring_start[r]);
}

// cell ownership: the first point (lowest index) that reaches a cell keeps it (IP:610-611)
__global__ void __launch_bounds__(256) proj_claim_kernel(const uint32_t* __restrict__ keys, const int32_t* __restrict__ col,
                                                         uint32_t n, int horizon, uint32_t* __restrict__ owner,
                                                         uint32_t* __restrict__ first_idx) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    uint32_t mine = kProjNone;
    if (i < n) {
        const uint32_t r = keys[i];
        const int c = col[i];
        if (r != 255u && c >= 0 && c < horizon) {
            atomicMin(&owner[r * (uint32_t)horizon + (uint32_t)c], i);
            mine = i;
        }
    }
    // block minimum, one atomic per block
    __shared__ uint32_t red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t m = red[0];
        for (int w = 1; w < 8; ++w) m = min(m, red[w]);
        if (m != kProjNone) atomicMin(first_idx, m);
    }
}

// findRotation, IP:495-526
__device__ inline void find_rotation(const ProjParams& P, double point_time, float rot[3]) {
    int f = 0;
    while (f < P.imu_pointer_cur) {
        if (point_time < P.imu_time[f]) break;
        ++f;
    }
    if (point_time > P.imu_time[f] || f == 0) {
        rot[0] = (float)P.imu_rx[f]; rot[1] = (float)P.imu_ry[f]; rot[2] = (float)P.imu_rz[f];
    } else {
        const int b = f - 1;
        const double rf = (point_time - P.imu_time[b]) / (P.imu_time[f] - P.imu_time[b]);
        const double rb = (P.imu_time[f] - point_time) / (P.imu_time[f] - P.imu_time[b]);
        rot[0] = (float)(P.imu_rx[f] * rf + P.imu_rx[b] * rb);
        rot[1] = (float)(P.imu_ry[f] * rf + P.imu_ry[b] * rb);
        rot[2] = (float)(P.imu_rz[f] * rf + P.imu_rz[b] * rb);
    }
}

__device__ inline void rotation_affine(const ProjParams& P, double point_time, Affine* T) {
    float rot[3];
    find_rotation(P, point_time, rot);
    const float pose[6] = {rot[0], rot[1], rot[2], 0.f, 0.f, 0.f};      // findPosition returns zeros (IP:528-536)
    Trig g;
    pose_to_affine_dev(pose, T, &g);
}

// Eigen::Transform<float,3,Affine>::inverse(): cofactor inverse of the linear part, translation = -inv * t
__device__ inline void affine_inverse_dev(const Affine& T, Affine* R) {
    const float* m = T.m;
    auto M = [&](int i, int j) { return m[4 * i + j]; };
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return M(i1, j1) * M(i2, j2) - M(i1, j2) * M(i2, j1);
    };
    const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    const float det = (c0 * M(0, 0) + c1 * M(1, 0)) + c2 * M(2, 0);
    const float invdet = 1.0f / det;
    float inv[3][3];
    inv[0][0] = c0 * invdet; inv[0][1] = c1 * invdet; inv[0][2] = c2 * invdet;
    inv[1][0] = cof(0, 1) * invdet; inv[1][1] = cof(1, 1) * invdet; inv[1][2] = cof(2, 1) * invdet;
    inv[2][0] = cof(0, 2) * invdet; inv[2][1] = cof(1, 2) * invdet; inv[2][2] = cof(2, 2) * invdet;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R->m[4 * i + j] = inv[i][j];
        R->m[4 * i + 3] = -((inv[i][0] * m[3] + inv[i][1] * m[7]) + inv[i][2] * m[11]);
    }
}

// transStartInverse from the first point that reaches deskewPoint (IP:551-555), one thread
__global__ void proj_start_kernel(ProjParams P, RawLayout in, const uint32_t* __restrict__ first_idx, Affine* __restrict__ start_inv) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t i = *first_idx;
    Affine I;
    for (int k = 0; k < 12; ++k) I.m[k] = (k % 5 == 0) ? 1.f : 0.f;
    if (i == kProjNone || !P.deskew) { *start_inv = I; return; }
    const float rel = *reinterpret_cast<const float*>(in.data + (size_t)i * in.stride + in.time_off);
    Affine T;
    rotation_affine(P, P.time_scan_cur + (double)rel, &T);
    affine_inverse_dev(T, start_inv);
}

// cloudExtraction as a scan over the cells: flag = cell owned
struct CellFlagIn {
    const uint32_t* owner;
    __device__ __forceinline__ uint32_t operator()(uint32_t c) const { return owner[c] != kProjNone ? 1u : 0u; }
    __device__ __forceinline__ void load_vec(uint32_t c, uint32_t (&v)[8]) const {
        const uint4 a = *reinterpret_cast<const uint4*>(owner + c);
        const uint4 b = *reinterpret_cast<const uint4*>(owner + c + 4);
        v[0] = a.x != kProjNone; v[1] = a.y != kProjNone; v[2] = a.z != kProjNone; v[3] = a.w != kProjNone;
        v[4] = b.x != kProjNone; v[5] = b.y != kProjNone; v[6] = b.z != kProjNone; v[7] = b.w != kProjNone;
    }
};
struct CellExtractOut {
    const uint32_t* owner;
    const float4* pts;
    const float* range;
    RawLayout in;
    ProjParams P;
    const Affine* start_inv;
    float4* extracted;
    float* point_range;
    int32_t* point_col_ind;
    uint32_t* row_prefix;          // n_scan entries: extracted points before the row
    __device__ __forceinline__ void operator()(uint32_t c, uint32_t flag, uint32_t pre) const {
        const uint32_t colc = c % (uint32_t)P.horizon;
        if (colc == 0) row_prefix[c / (uint32_t)P.horizon] = pre;
        if (!flag) return;
        const uint32_t i = owner[c];
        float4 p = pts[i];
        if (P.deskew) {                                                  // deskewPoint, IP:538-569
            const float rel = *reinterpret_cast<const float*>(in.data + (size_t)i * in.stride + in.time_off);
            Affine T;
            rotation_affine(P, P.time_scan_cur + (double)rel, &T);
            // transBt = transStartInverse * transFinal (Affine3f product)
            const Affine S = *start_inv;
            Affine Bt;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    Bt.m[4 * r + j] = (S.m[4 * r] * T.m[j] + S.m[4 * r + 1] * T.m[4 + j]) + S.m[4 * r + 2] * T.m[8 + j];
                Bt.m[4 * r + 3] = ((S.m[4 * r] * T.m[3] + S.m[4 * r + 1] * T.m[7]) + S.m[4 * r + 2] * T.m[11]) + S.m[4 * r + 3];
            }
            const float3 q = apply_affine(Bt, p.x, p.y, p.z);
            p.x = q.x; p.y = q.y; p.z = q.z;
        }
        extracted[pre] = p;
        point_range[pre] = range[i];
        point_col_ind[pre] = (int32_t)colc;
    }
    __device__ __forceinline__ void store_vec(uint32_t c, const uint32_t (&v)[8], uint32_t pre) const {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            (*this)(c + k, v[k], pre);
            pre += v[k];
        }
    }
};

// start_ring_index[i] = count_before - 1 + 5, end_ring_index[i] = count_after - 1 - 5 (IP:631, 646)
__global__ void proj_ring_index_kernel(const uint32_t* __restrict__ row_prefix, const uint32_t* __restrict__ total,
                                       int n_scan, int32_t* __restrict__ start_idx, int32_t* __restrict__ end_idx) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_scan) return;
    const uint32_t before = row_prefix[r];
    const uint32_t after = r + 1 < n_scan ? row_prefix[r + 1] : *total;
    start_idx[r] = (int32_t)before - 1 + 5;
    end_idx[r] = (int32_t)after - 1 - 5;
}

}  // namespace lvreg
