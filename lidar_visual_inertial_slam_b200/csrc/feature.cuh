// feature.cuh -- "next" row 8f-1 of SURVEY.md: the reference's FeatureExtraction node body
// (lidar_odometry/src/featureExtraction.cpp) on the device, so that the scan-to-map path can
// start from the deskewed, ring-ordered cloud instead of host-side feature clouds.
//
//   fe_smooth_kernel    calculateSmoothness :87-111 + markOccludedPoints :113-148   (flat, parallel)
//   fe_ring_kernel      extractFeatures :150-245 -- one block per ring.  The six sectors of a ring
//                       are processed in order (picked flags leak up to 5 points into the next
//                       sector, never into another ring); each sector's [sp, ep) is sorted by
//                       (curvature, index) with a bitonic network in shared memory -- the index
//                       tie-break pins std::sort's unspecified order -- and the greedy corner /
//                       surface labelling runs on one thread over shared-memory state
//   fe_ring_bbox_kernel + fe_surf_keys_kernel + (two stable radix sorts) + centroid_kernel
//                       the per-ring pcl::VoxelGrid of the surface candidates (:236-241), all
//                       rings in one pass: keys are PCL's idx computed with each ring's own
//                       bounding box, order = (ring, idx, input order)
#pragma once

#include "common.cuh"
#include "prims.cuh"

namespace lvreg {

constexpr int kFeMaxCorner = 40;                 // largestPickedNum <= 40 (:178)
constexpr int kFeCornerStride = 6 * kFeMaxCorner;

__global__ void __launch_bounds__(256) fe_smooth_kernel(const float* __restrict__ range,
                                                        const int32_t* __restrict__ col, int n,
                                                        float* __restrict__ curv, uint8_t* __restrict__ picked) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float c = 0.0f;
    if (i >= 5 && i < n - 5) {
        const float diff = range[i - 2] + range[i - 1] - range[i] * 4 + range[i + 1] + range[i + 2];
        c = diff * diff;
    }
    curv[i] = c;
    if (i >= 5 && i < n - 6) {          // `picked` was zeroed; all writers store 1
        const float d1 = range[i], d2 = range[i + 1];
        int cd = col[i + 1] - col[i];
        cd = cd < 0 ? -cd : cd;
        if (cd < 10) {
            if ((double)(d1 - d2) > 0.3) {
                picked[i - 1] = 1;
                picked[i] = 1;
            } else if ((double)(d2 - d1) > 0.3) {
                picked[i + 1] = 1;
                picked[i + 2] = 1;
            }
        }
        const float diff1 = fabsf(range[i - 1] - range[i]);
        const float diff2 = fabsf(range[i + 1] - range[i]);
        if ((double)diff1 > 0.1 * (double)range[i] && (double)diff2 > 0.1 * (double)range[i]) picked[i] = 1;
    }
}

typedef unsigned long long fe_u64;

// in-place bitonic sort of `len_pow2` keys in shared memory (ascending), whole block
__device__ __forceinline__ void fe_bitonic(fe_u64* keys, int len_pow2) {
    for (int k = 2; k <= len_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < len_pow2; t += blockDim.x) {
                const int partner = t ^ j;
                if (partner > t) {
                    const fe_u64 a = keys[t], b = keys[partner];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) { keys[t] = b; keys[partner] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// dynamic shared memory layout of fe_ring_kernel for a region of `cap` points and `sort_cap` keys
__host__ __device__ inline size_t fe_ring_smem_bytes(int cap, int sort_cap) {
    return (size_t)sort_cap * 8 + (size_t)cap * (4 + 4 + 1 + 1) + 64;
}

__global__ void __launch_bounds__(256) fe_ring_kernel(const float* __restrict__ curv_g, const uint8_t* __restrict__ picked_g,
                                                      const int32_t* __restrict__ col_g, int n,
                                                      const int32_t* __restrict__ start_ring,
                                                      const int32_t* __restrict__ end_ring, float edge_th, float surf_th,
                                                      int cap, int sort_cap, int8_t* __restrict__ label_g,
                                                      uint32_t* __restrict__ surf_flag_g, uint8_t* __restrict__ ring_of_g,
                                                      int32_t* __restrict__ corner_idx /*[n_scan][240]*/,
                                                      int32_t* __restrict__ corner_cnt, int32_t* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char fe_smem[];
    fe_u64* skeys = reinterpret_cast<fe_u64*>(fe_smem);
    float* curv = reinterpret_cast<float*>(fe_smem + (size_t)sort_cap * 8);
    int32_t* col = reinterpret_cast<int32_t*>(curv + cap);
    uint8_t* picked = reinterpret_cast<uint8_t*>(col + cap);
    int8_t* label = reinterpret_cast<int8_t*>(picked + cap);

    const int r = blockIdx.x;
    const int start = start_ring[r], end = end_ring[r];
    int lo = start - 5, hi = end + 4;              // every index this ring can read or mark
    if (lo < 0) lo = 0;
    if (hi > n - 1) hi = n - 1;
    const int len = hi - lo + 1;
    if (threadIdx.x == 0) corner_cnt[r] = 0;
    if (len <= 0 || end - start < 1) return;
    if (len > cap) {
        if (threadIdx.x == 0) atomicExch(err_flag, 1);
        return;
    }
    for (int t = threadIdx.x; t < len; t += blockDim.x) {
        curv[t] = curv_g[lo + t];
        col[t] = col_g[lo + t];
        picked[t] = picked_g[lo + t];
        label[t] = 0;
    }
    __syncthreads();

    int ncorner = 0;
    for (int j = 0; j < 6; ++j) {
        const int sp = (start * (6 - j) + end * j) / 6;
        const int ep = (start * (5 - j) + end * (j + 1)) / 6 - 1;
        if (sp >= ep) continue;
        const int slen = ep - sp;                  // std::sort(begin+sp, begin+ep): ep itself is NOT sorted
        int p2 = 1;
        while (p2 < slen) p2 <<= 1;
        if (p2 > sort_cap) {
            if (threadIdx.x == 0) atomicExch(err_flag, 1);
            return;
        }
        for (int t = threadIdx.x; t < p2; t += blockDim.x) {
            fe_u64 key = ~0ull;
            if (t < slen) key = ((fe_u64)__float_as_uint(curv[sp + t - lo]) << 32) | (fe_u64)(uint32_t)(sp + t);
            skeys[t] = key;
        }
        __syncthreads();
        fe_bitonic(skeys, p2);
        if (threadIdx.x == 0) {
            // position k in [sp, ep) -> sorted index; position ep -> ep itself
            auto sorted_ind = [&](int k) { return k == ep ? ep : (int)(uint32_t)skeys[k - sp]; };
            auto mark_neighbours = [&](int ind) {
                for (int l = 1; l <= 5; ++l) {
                    if (ind + l > hi) break;
                    int cd = col[ind + l - lo] - col[ind + l - 1 - lo];
                    cd = cd < 0 ? -cd : cd;
                    if (cd > 10) break;
                    picked[ind + l - lo] = 1;
                }
                for (int l = -1; l >= -5; --l) {
                    if (ind + l < lo) break;
                    int cd = col[ind + l - lo] - col[ind + l + 1 - lo];
                    cd = cd < 0 ? -cd : cd;
                    if (cd > 10) break;
                    picked[ind + l - lo] = 1;
                }
            };
            // [sp, ep) is sorted ascending and position ep is unsorted, so both greedy loops can stop
            // as soon as the sorted part can no longer satisfy their threshold (same result as the
            // reference's full sweeps, which only test and skip from there on)
            int largest = 0;
            for (int k = ep; k >= sp; --k) {
                const int ind = sorted_ind(k);
                if (k < ep && !(curv[ind - lo] > edge_th)) break;
                if (picked[ind - lo] == 0 && curv[ind - lo] > edge_th) {
                    ++largest;
                    if (largest <= kFeMaxCorner) {
                        label[ind - lo] = 1;
                        corner_idx[r * kFeCornerStride + ncorner++] = ind;
                    } else {
                        break;
                    }
                    picked[ind - lo] = 1;
                    mark_neighbours(ind);
                }
            }
            for (int k = sp; k <= ep; ++k) {
                const int ind = sorted_ind(k);
                if (k < ep && !(curv[ind - lo] < surf_th)) k = ep - 1;      // jump to the unsorted tail element
                else if (picked[ind - lo] == 0 && curv[ind - lo] < surf_th) {
                    label[ind - lo] = -1;
                    picked[ind - lo] = 1;
                    mark_neighbours(ind);
                }
            }
        }
        __syncthreads();
        // surfaceCloudScan: every k in [sp, ep] whose label is <= 0 (:229-233)
        for (int k = sp + threadIdx.x; k <= ep; k += blockDim.x) {
            surf_flag_g[k] = label[k - lo] <= 0 ? 1u : 0u;
            ring_of_g[k] = (uint8_t)r;
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < len; t += blockDim.x) label_g[lo + t] = label[t];
    if (threadIdx.x == 0) corner_cnt[r] = ncorner;
}

// corner cloud in ring order: per-ring counts -> offsets (one block), then gather
__global__ void __launch_bounds__(256) fe_corner_gather_kernel(const float4* __restrict__ pts,
                                                               const int32_t* __restrict__ corner_idx,
                                                               const int32_t* __restrict__ corner_cnt, int n_scan,
                                                               float4* __restrict__ out, uint32_t* __restrict__ total) {
    __shared__ int off_s;
    const int r = blockIdx.x;                     // one block per ring
    if (threadIdx.x == 0) {
        int s = 0;
        for (int q = 0; q < r; ++q) s += corner_cnt[q];
        off_s = s;
        if (r == n_scan - 1) *total = (uint32_t)(s + corner_cnt[r]);
    }
    __syncthreads();
    const int c = corner_cnt[r];
    for (int t = threadIdx.x; t < c; t += blockDim.x) out[off_s + t] = pts[corner_idx[r * kFeCornerStride + t]];
}

struct FlagIn {
    const uint32_t* f;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return f[i]; }
    __device__ __forceinline__ void load_vec(uint32_t i, uint32_t (&v)[8]) const {
        const uint4 a = *reinterpret_cast<const uint4*>(f + i);
        const uint4 b = *reinterpret_cast<const uint4*>(f + i + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};
// exclusive prefix to `pos[i]`, and the compacted candidate list cand[pos] = i
struct CompactOut {
    uint32_t* pos;
    uint32_t* cand;
    __device__ __forceinline__ void operator()(uint32_t i, uint32_t flag, uint32_t pre) const {
        pos[i] = pre;
        if (flag) cand[pre] = i;
    }
    __device__ __forceinline__ void store_vec(uint32_t i, const uint32_t (&v)[8], uint32_t pre) const {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            pos[i + k] = pre;
            if (v[k]) cand[pre] = i + k;
            pre += v[k];
        }
    }
};

struct RingSpec {                 // per-ring VoxelGrid spec (device computed, PCL arithmetic)
    float inv;
    int min_b[3];
    int mul[3];
    int passthrough;              // PCL's overflow rule: the ring's cloud is returned unchanged
    uint32_t first_pos;           // position of the ring's first candidate in the compacted list
};

// one block per ring: bounding box of the ring's surface candidates -> RingSpec
__global__ void __launch_bounds__(256) fe_ring_bbox_kernel(const float4* __restrict__ pts,
                                                           const uint32_t* __restrict__ surf_flag,
                                                           const uint32_t* __restrict__ pos,
                                                           const int32_t* __restrict__ start_ring,
                                                           const int32_t* __restrict__ end_ring, int n, float leaf,
                                                           RingSpec* __restrict__ spec) {
    __shared__ float red[6][8];
    const int r = blockIdx.x;
    int a = start_ring[r], b = end_ring[r] - 1;
    if (a < 0) a = 0;
    if (b > n - 1) b = n - 1;
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = a + threadIdx.x; i <= b; i += blockDim.x) {
        if (surf_flag[i]) {
            const float4 p = pts[i];
            mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
            mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int k = 0; k < 3; ++k) { red[k][warp] = mn[k]; red[3 + k][warp] = mx[k]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k)
            for (int w = 1; w < 8; ++w) {
                red[k][0] = fminf(red[k][0], red[k][w]);
                red[3 + k][0] = fmaxf(red[3 + k][0], red[3 + k][w]);
            }
        RingSpec s;
        s.inv = 1.0f / leaf;
        s.passthrough = 0;
        s.first_pos = a <= b ? pos[a] : 0;
        long long d[3];
        int div_b[3] = {1, 1, 1};
        const bool empty = red[0][0] > red[3][0];
        for (int k = 0; k < 3; ++k) {
            const float lo = empty ? 0.f : red[k][0], hi = empty ? 0.f : red[3 + k][0];
            d[k] = (long long)((hi - lo) * s.inv) + 1;
            s.min_b[k] = (int)floorf(lo * s.inv);
            div_b[k] = (int)floorf(hi * s.inv) - s.min_b[k] + 1;
        }
        if (d[0] * d[1] * d[2] > 2147483647ll) s.passthrough = 1;
        s.mul[0] = 1;
        s.mul[1] = div_b[0];
        s.mul[2] = div_b[0] * div_b[1];
        spec[r] = s;
    }
}

// per candidate: PCL's voxel idx with its ring's spec (or its rank inside the ring when the ring is
// passed through unchanged), payload = candidate position
__global__ void __launch_bounds__(256) fe_surf_keys_kernel(const float4* __restrict__ pts,
                                                           const uint32_t* __restrict__ cand, uint32_t n_cand,
                                                           const uint8_t* __restrict__ ring_of,
                                                           const RingSpec* __restrict__ spec, uint32_t* __restrict__ keys,
                                                           uint32_t* __restrict__ vals, uint32_t* __restrict__ idx_by_pos) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n_cand) return;
    const uint32_t i = cand[j];
    const RingSpec s = spec[ring_of[i]];
    uint32_t key;
    if (s.passthrough) {
        key = j - s.first_pos;
    } else {
        const float4 p = pts[i];
        const int ix = (int)(floorf(p.x * s.inv) - (float)s.min_b[0]);
        const int iy = (int)(floorf(p.y * s.inv) - (float)s.min_b[1]);
        const int iz = (int)(floorf(p.z * s.inv) - (float)s.min_b[2]);
        key = (uint32_t)(ix * s.mul[0] + iy * s.mul[1] + iz * s.mul[2]);
    }
    keys[j] = key;
    vals[j] = j;
    idx_by_pos[j] = key;
}

// second sort key: the ring of each (idx-sorted) candidate
__global__ void __launch_bounds__(256) fe_ring_keys_kernel(const uint32_t* __restrict__ sorted_vals, uint32_t n_cand,
                                                           const uint32_t* __restrict__ cand,
                                                           const uint8_t* __restrict__ ring_of, uint32_t* __restrict__ keys) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n_cand) return;
    keys[j] = ring_of[cand[sorted_vals[j]]];
}

struct RingVoxelHeadIn {          // a new voxel starts where (ring, idx) changes
    const uint32_t* ring_keys;    // sorted
    const uint32_t* vals;         // candidate positions, sorted by (ring, idx)
    const uint32_t* idx_by_pos;
    __device__ __forceinline__ uint32_t operator()(uint32_t j) const {
        if (j == 0) return 1u;
        return (ring_keys[j] != ring_keys[j - 1] || idx_by_pos[vals[j]] != idx_by_pos[vals[j - 1]]) ? 1u : 0u;
    }
    __device__ __forceinline__ void load_vec(uint32_t j, uint32_t (&v)[8]) const {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (*this)(j + k);
    }
};

// candidate positions -> point indices, so that centroid_kernel can gather from the input cloud
__global__ void __launch_bounds__(256) fe_point_index_kernel(const uint32_t* __restrict__ sorted_vals, uint32_t n_cand,
                                                             const uint32_t* __restrict__ cand, uint32_t* __restrict__ out) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n_cand) return;
    out[j] = cand[sorted_vals[j]];
}

}  // namespace lvreg
