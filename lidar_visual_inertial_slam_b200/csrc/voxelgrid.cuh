// voxelgrid.cuh -- sort-by-voxel-key replacement for pcl::VoxelGrid<PointXYZI>::filter
// (call sites MO:959-965 local map, MO:991-997 current scan; algorithm SURVEY Appendix A.1) and
// the fused keyframe transform + concatenation of extractCloud (MO:931-957, transformPointCloud
// MO:347-366).
//
//   transform_concat_kernel  keyframe clouds (sensor frame) -> one world-frame cloud + bbox
//   minmax_kernel            getMinMax3D
//   voxel_keys_kernel        idx = ijk . divb_mul, bit-exact to PCL's fp32/int32 arithmetic
//   (radix sort, prims.cuh)  stable, so within-voxel order == input order
//   head flags + scan        one output per distinct idx, ascending idx
//   centroid_kernel          CentroidPoint: sequential fp32 sums in sorted order, / count
//
// HBM traffic per input point (16-byte float4 points): transform+concat 16 R + 16 W, bbox 0
// (fused), keys 16 R + 8 W, sort passes 16 R/W each, centroid 16 R (gather) -- all streaming.
#pragma once

#include "common.cuh"
#include "prims.cuh"

namespace lvreg {

// ---- AoS (PCL layout) <-> float4 -----------------------------------------------------------
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* __restrict__ src, uint32_t n,
                                                   uint32_t stride, uint32_t ioff,
                                                   float4* __restrict__ dst) {
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = src + (size_t)i * stride;
    float4 v;
    if ((stride & 15u) == 0) {
        float4 a = *reinterpret_cast<const float4*>(p);          // x y z pad
        v.x = a.x; v.y = a.y; v.z = a.z;
    } else {
        const float* f = reinterpret_cast<const float*>(p);
        v.x = f[0]; v.y = f[1]; v.z = f[2];
    }
    v.w = *reinterpret_cast<const float*>(p + ioff);
    dst[i] = v;
}

__global__ void __launch_bounds__(256) unpack_kernel(const float4* __restrict__ src, uint32_t n,
                                                     uint32_t stride, uint32_t ioff,
                                                     uint8_t* __restrict__ dst) {
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float4 v = src[i];
    uint8_t* p = dst + (size_t)i * stride;
    float* f = reinterpret_cast<float*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z;
    if (stride >= 32 && ioff != 12) f[3] = 1.0f;                  // PCL_ADD_POINT4D: data[3] = 1
    *reinterpret_cast<float*>(p + ioff) = v.w;
}

// ---- bbox ------------------------------------------------------------------------------------
// mm[0..2] = ordered(min xyz), mm[3..5] = ordered(max xyz); init with 0xffffffff / 0
__device__ __forceinline__ void block_minmax_commit(float mnx, float mny, float mnz, float mxx,
                                                    float mxy, float mxz, uint32_t* mm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
    }
    __shared__ float red[6][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = mnx; red[1][warp] = mny; red[2][warp] = mnz;
        red[3][warp] = mxx; red[4][warp] = mxy; red[5][warp] = mxz;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[threadIdx.x][0];
        const int nw = blockDim.x >> 5;
        for (int w = 1; w < nw; ++w)
            v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][w]) : fmaxf(v, red[threadIdx.x][w]);
        if (threadIdx.x < 3) atomicMin(&mm[threadIdx.x], float_to_ordered(v));
        else atomicMax(&mm[threadIdx.x], float_to_ordered(v));
    }
}

__global__ void __launch_bounds__(256) minmax_kernel(const float4* __restrict__ pts, uint32_t n,
                                                     uint32_t* __restrict__ mm) {
    float mnx = 3.4e38f, mny = 3.4e38f, mnz = 3.4e38f, mxx = -3.4e38f, mxy = -3.4e38f, mxz = -3.4e38f;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        float4 p = ld_stream(pts + i);
        mnx = fminf(mnx, p.x); mny = fminf(mny, p.y); mnz = fminf(mnz, p.z);
        mxx = fmaxf(mxx, p.x); mxy = fmaxf(mxy, p.y); mxz = fmaxf(mxz, p.z);
    }
    block_minmax_commit(mnx, mny, mnz, mxx, mxy, mxz, mm);
}

// ---- fused transformPointCloud + concatenation + bbox ----------------------------------------
struct Segment {                 // one selected keyframe cloud
    const float4* src;           // sensor-frame points
    uint32_t begin;              // first index in the concatenated cloud
    uint32_t n;
    Affine T;                    // pclPointToAffine3f(cloudKeyPoses6D[id]), computed on the host
    // voxel-ordered world-frame cache only (voxelgrid_bucket.cuh): per point its voxel coordinates relative to
    // kminb, packed 11 | 11 | 10 bits (x | y | z)
    const uint32_t* wkey;
    int kminb[3];
};

__global__ void __launch_bounds__(256) transform_concat_kernel(const Segment* __restrict__ segs,
                                                               uint32_t nseg, uint32_t total,
                                                               float4* __restrict__ out,
                                                               uint32_t* __restrict__ mm) {
    float mnx = 3.4e38f, mny = 3.4e38f, mnz = 3.4e38f, mxx = -3.4e38f, mxy = -3.4e38f, mxz = -3.4e38f;
    __shared__ uint32_t seg0_s;
    // chunks of 2048 consecutive points: one binary search per chunk (last segment with begin <= first point; empty
    // segments share a begin and are skipped), then the points walk forward from it
    for (uint32_t c0 = blockIdx.x * 2048u; c0 < total; c0 += gridDim.x * 2048u) {
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t lo = 0, hi = nseg;
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if (segs[mid].begin <= c0) lo = mid; else hi = mid;
            }
            seg0_s = lo;
        }
        __syncthreads();
        uint32_t lo = seg0_s;
        for (int r = 0; r < 8; ++r) {
            const uint32_t i = c0 + r * 256 + threadIdx.x;
            if (i >= total) break;
            while (lo + 1 < nseg && segs[lo + 1].begin <= i) ++lo;
            const Segment& s = segs[lo];
            float4 p = ld_stream(s.src + (i - s.begin));
            float3 q = apply_affine(s.T, p.x, p.y, p.z);
            out[i] = make_float4(q.x, q.y, q.z, p.w);
            mnx = fminf(mnx, q.x); mny = fminf(mny, q.y); mnz = fminf(mnz, q.z);
            mxx = fmaxf(mxx, q.x); mxy = fmaxf(mxy, q.y); mxz = fmaxf(mxz, q.z);
        }
    }
    block_minmax_commit(mnx, mny, mnz, mxx, mxy, mxz, mm);
}

// transformPointCloud for one cloud (stage-level API)
__global__ void __launch_bounds__(256) transform_kernel(const float4* __restrict__ in, uint32_t n,
                                                        Affine T, float4* __restrict__ out) {
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float4 p = in[i];
    float3 q = apply_affine(T, p.x, p.y, p.z);
    out[i] = make_float4(q.x, q.y, q.z, p.w);
}

// ---- voxel keys --------------------------------------------------------------------------------
struct VoxelSpec {
    float inv;           // 1.0f / leaf
    int min_b[3];        // (int)floor(min_p * inv)
    int mul[3];          // divb_mul = (1, div_b.x, div_b.x * div_b.y)
    int key_bits;        // bits needed for the largest idx
};

// grid-stride; also accumulates the radix-sort digit histograms of the keys it produces
__global__ void __launch_bounds__(256) voxel_keys_kernel(const float4* __restrict__ pts, uint32_t n,
                                                         VoxelSpec vs, uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ vals, uint32_t* __restrict__ ghist,
                                                         int passes) {
    __shared__ uint32_t hist_s[8][kSortMaxPasses][256];
    HistAccumulator acc;
    acc.init(hist_s);
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        float4 p = ld_stream(pts + i);
        int ix = (int)(floorf(p.x * vs.inv) - (float)vs.min_b[0]);
        int iy = (int)(floorf(p.y * vs.inv) - (float)vs.min_b[1]);
        int iz = (int)(floorf(p.z * vs.inv) - (float)vs.min_b[2]);
        const uint32_t key = (uint32_t)(ix * vs.mul[0] + iy * vs.mul[1] + iz * vs.mul[2]);
        keys[i] = key;
        vals[i] = i;
        acc.add(key, passes);
    }
    acc.flush(ghist, passes);
}

// ---- the same two kernels over CACHED world-frame keyframe clouds (laserCloudMapContainer, MO:942-954) ----
// The reference transforms a keyframe cloud once and keeps the result (MO:945-952); the concatenation is what it
// redoes per scan.  Here the cached clouds are not even concatenated: the key kernel walks the segment table, and
// the sort's payload is (segment << kSegShift | offset) instead of a global index, so the centroid kernel gathers
// straight from the cached clouds.  Input order (segment order, then point order) is what the stable sort keeps,
// exactly as with a concatenated cloud.
constexpr int kSegShift = 22;                       // <= 1024 segments of <= 4 Mi points each
constexpr uint32_t kSegOffMask = (1u << kSegShift) - 1u;

__global__ void __launch_bounds__(256) voxel_keys_seg_kernel(const Segment* __restrict__ segs, uint32_t nseg, uint32_t n,
                                                             VoxelSpec vs, uint32_t* __restrict__ keys,
                                                             uint32_t* __restrict__ vals, uint32_t* __restrict__ ghist,
                                                             int passes) {
    __shared__ uint32_t hist_s[8][kSortMaxPasses][256];
    HistAccumulator acc;
    acc.init(hist_s);
    __shared__ uint32_t seg0_s;
    // a block takes chunks of 2048 consecutive points: one binary search per chunk finds its first segment, the
    // points then walk forward from it (a chunk spans one or two keyframe clouds) -- not a search per point
    for (uint32_t c0 = blockIdx.x * 2048u; c0 < n; c0 += gridDim.x * 2048u) {
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t lo = 0, hi = nseg;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (segs[mid].begin <= c0) lo = mid; else hi = mid;
            }
            seg0_s = lo;
        }
        __syncthreads();
        uint32_t lo = seg0_s;
#pragma unroll 4
        for (int r = 0; r < 8; ++r) {
            const uint32_t i = c0 + r * 256 + threadIdx.x;
            if (i >= n) break;
            while (lo + 1 < nseg && segs[lo + 1].begin <= i) ++lo;
            const Segment& sg = segs[lo];
            const uint32_t off = i - sg.begin;
            const float4 p = ld_stream(sg.src + off);
            const int ix = (int)(floorf(p.x * vs.inv) - (float)vs.min_b[0]);
            const int iy = (int)(floorf(p.y * vs.inv) - (float)vs.min_b[1]);
            const int iz = (int)(floorf(p.z * vs.inv) - (float)vs.min_b[2]);
            const uint32_t key = (uint32_t)(ix * vs.mul[0] + iy * vs.mul[1] + iz * vs.mul[2]);
            keys[i] = key;
            vals[i] = (lo << kSegShift) | off;
            acc.add(key, passes);
        }
    }
    acc.flush(ghist, passes);
}

__global__ void __launch_bounds__(128) centroid_seg_kernel(const Segment* __restrict__ segs,
                                                           const uint32_t* __restrict__ sorted_keys,
                                                           const uint32_t* __restrict__ sorted_vals,
                                                           const uint32_t* __restrict__ start,
                                                           const uint32_t* __restrict__ nvox_p, uint32_t n,
                                                           float4* __restrict__ out) {
    const uint32_t nvox = *nvox_p;
    uint32_t v = blockIdx.x * 128 + threadIdx.x;
    if (v >= nvox) return;
    const uint32_t b = start[v];
    const uint32_t e = (v + 1 < nvox) ? start[v + 1] : n;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (uint32_t j = b; j < e; ++j) {
        const uint32_t pv = sorted_vals[j];
        const float4 p = __ldg(segs[pv >> kSegShift].src + (pv & kSegOffMask));
        sx += p.x; sy += p.y; sz += p.z; si += p.w;
    }
    const float cnt = (float)(e - b);
    out[v] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
}

// transformPointCloud of one keyframe cloud into the cache + its bounding box (slot = 6 ordered uints, initialised
// to 0xffffffff x3 / 0 x3 by the caller)
__global__ void __launch_bounds__(256) transform_bbox_kernel(const float4* __restrict__ in, uint32_t n, Affine T,
                                                             float4* __restrict__ out, uint32_t* __restrict__ mm) {
    float mnx = 3.4e38f, mny = 3.4e38f, mnz = 3.4e38f, mxx = -3.4e38f, mxy = -3.4e38f, mxz = -3.4e38f;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const float4 p = in[i];
        const float3 q = apply_affine(T, p.x, p.y, p.z);
        out[i] = make_float4(q.x, q.y, q.z, p.w);
        mnx = fminf(mnx, q.x); mny = fminf(mny, q.y); mnz = fminf(mnz, q.z);
        mxx = fmaxf(mxx, q.x); mxy = fmaxf(mxy, q.y); mxz = fmaxf(mxz, q.z);
    }
    block_minmax_commit(mnx, mny, mnz, mxx, mxy, mxz, mm);
}

__global__ void bbox_slots_init_kernel(uint32_t* mm, uint32_t nslots) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nslots * 6) mm[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
}

// ---- run heads -> voxel starts -----------------------------------------------------------------
struct HeadFlagIn {
    const uint32_t* keys;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
    }
    // 8 consecutive flags starting at a multiple of 8 (two 128-bit loads + one scalar)
    __device__ __forceinline__ void load_vec(uint32_t i, uint32_t (&v)[8]) const {
        const uint4 a = *reinterpret_cast<const uint4*>(keys + i);
        const uint4 b = *reinterpret_cast<const uint4*>(keys + i + 4);
        const uint32_t prev = i ? keys[i - 1] : ~a.x;
        v[0] = a.x != prev; v[1] = a.y != a.x; v[2] = a.z != a.y; v[3] = a.w != a.z;
        v[4] = b.x != a.w;  v[5] = b.y != b.x; v[6] = b.z != b.y; v[7] = b.w != b.z;
    }
};
struct VoxelStartOut {
    uint32_t* start;
    __device__ __forceinline__ void operator()(uint32_t i, uint32_t flag, uint32_t pre) const {
        if (flag) start[pre] = i;
    }
    __device__ __forceinline__ void store_vec(uint32_t i, const uint32_t (&v)[8], uint32_t pre) const {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (v[k]) start[pre] = i + k;
            pre += v[k];
        }
    }
};

// ---- centroids -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) centroid_kernel(const float4* __restrict__ pts,
                                                       const uint32_t* __restrict__ sorted_keys,
                                                       const uint32_t* __restrict__ sorted_vals,
                                                       const uint32_t* __restrict__ start,
                                                       const uint32_t* __restrict__ nvox_p, uint32_t n,
                                                       float4* __restrict__ out,
                                                       uint32_t* __restrict__ out_keys) {
    const uint32_t nvox = *nvox_p;
    uint32_t v = blockIdx.x * 128 + threadIdx.x;
    if (v >= nvox) return;
    const uint32_t b = start[v];
    const uint32_t e = (v + 1 < nvox) ? start[v + 1] : n;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (uint32_t j = b; j < e; ++j) {
        float4 p = __ldg(pts + sorted_vals[j]);
        sx += p.x; sy += p.y; sz += p.z; si += p.w;
    }
    const float cnt = (float)(e - b);
    out[v] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
    if (out_keys) out_keys[v] = sorted_keys[b];
}


// ---- small clouds: the whole filter in ONE block ------------------------------------------------------
// Up to kVgSmallMax points (key poses in extractNearby MO:910-911, the MID360 corner cloud, ...): bounding
// box, PCL's bounds / overflow rule, keys, a bitonic sort of (key << 32 | input index) -- stable by
// construction --, run heads and the sequential centroids, without leaving the SM and without a host round
// trip in between.  Same fp32 expressions as the multi-kernel path, so the result is bit-identical to it.
constexpr int kVgSmallMax = 2048;
constexpr int kVgSmallThreads = 1024;
struct VgSmallInfo {
    uint32_t nvox;
    int32_t passthrough;
    float mn[3], mx[3];
};

__global__ void __launch_bounds__(kVgSmallThreads) voxelgrid_small_kernel(const float4* __restrict__ pts, uint32_t n, float leaf,
                                                                          float4* __restrict__ out, uint32_t* __restrict__ out_keys,
                                                                          uint32_t* __restrict__ point_keys,
                                                                          VgSmallInfo* __restrict__ info) {
    __shared__ unsigned long long sk[kVgSmallMax];
    __shared__ uint32_t start[kVgSmallMax];
    __shared__ float red[6][32];
    __shared__ uint32_t wsum[32];
    __shared__ VoxelSpec vs;
    __shared__ int pass_s;
    __shared__ uint32_t nv_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float mnx = 3.4e38f, mny = 3.4e38f, mnz = 3.4e38f, mxx = -3.4e38f, mxy = -3.4e38f, mxz = -3.4e38f;
    for (uint32_t i = tid; i < n; i += kVgSmallThreads) {
        const float4 p = pts[i];
        mnx = fminf(mnx, p.x); mny = fminf(mny, p.y); mnz = fminf(mnz, p.z);
        mxx = fmaxf(mxx, p.x); mxy = fmaxf(mxy, p.y); mxz = fmaxf(mxz, p.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o)); mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
    }
    if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mny; red[2][warp] = mnz; red[3][warp] = mxx; red[4][warp] = mxy; red[5][warp] = mxz; }
    __syncthreads();
    if (tid == 0) {
        float mn[3], mx[3];
        for (int a = 0; a < 3; ++a) {
            float lo = red[a][0], hi = red[3 + a][0];
            for (int w = 1; w < kVgSmallThreads / 32; ++w) { lo = fminf(lo, red[a][w]); hi = fmaxf(hi, red[3 + a][w]); }
            mn[a] = lo; mx[a] = hi;
            info->mn[a] = lo; info->mx[a] = hi;
        }
        // PCL voxel_grid.hpp: leaf-size overflow rule and bounds, fp32 exactly as PCL computes them
        const float inv = 1.0f / leaf;
        long long d[3];
        for (int a = 0; a < 3; ++a) d[a] = (long long)((mx[a] - mn[a]) * inv) + 1;
        pass_s = d[0] * d[1] * d[2] > 2147483647ll ? 1 : 0;
        vs.inv = inv;
        int div_b[3];
        for (int a = 0; a < 3; ++a) {
            vs.min_b[a] = (int)floorf(mn[a] * inv);
            const int max_b = (int)floorf(mx[a] * inv);
            div_b[a] = max_b - vs.min_b[a] + 1;
        }
        vs.mul[0] = 1; vs.mul[1] = div_b[0]; vs.mul[2] = div_b[0] * div_b[1];
        vs.key_bits = 32;
    }
    __syncthreads();
    if (pass_s) {
        for (uint32_t i = tid; i < n; i += kVgSmallThreads) {
            out[i] = pts[i];
            if (point_keys) point_keys[i] = 0;
        }
        if (tid == 0) { info->nvox = n; info->passthrough = 1; }
        return;
    }
    // keys; the sort size is the next power of two, padded with keys that sort last
    uint32_t n2 = 2;
    while (n2 < n) n2 <<= 1;
    for (uint32_t i = tid; i < n2; i += kVgSmallThreads) {
        unsigned long long kk = ~0ull;
        if (i < n) {
            const float4 p = pts[i];
            const int ix = (int)(floorf(p.x * vs.inv) - (float)vs.min_b[0]);
            const int iy = (int)(floorf(p.y * vs.inv) - (float)vs.min_b[1]);
            const int iz = (int)(floorf(p.z * vs.inv) - (float)vs.min_b[2]);
            const uint32_t key = (uint32_t)(ix * vs.mul[0] + iy * vs.mul[1] + iz * vs.mul[2]);
            if (point_keys) point_keys[i] = key;
            kk = ((unsigned long long)key << 32) | (unsigned long long)i;
        }
        sk[i] = kk;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= n2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = tid; t < n2 / 2; t += kVgSmallThreads) {
                const uint32_t i = 2 * t - (t & (j - 1));          // lower index of the pair
                const uint32_t l = i + j;
                const unsigned long long a = sk[i], b = sk[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { sk[i] = b; sk[l] = a; }
            }
            __syncthreads();
        }
    // run heads -> voxel starts (two consecutive elements per thread, block exclusive scan)
    uint32_t f0 = 0, f1 = 0;
    {
        const uint32_t j0 = 2 * tid, j1 = 2 * tid + 1;
        if (j0 < n) f0 = (j0 == 0 || (uint32_t)(sk[j0] >> 32) != (uint32_t)(sk[j0 - 1] >> 32)) ? 1u : 0u;
        if (j1 < n) f1 = ((uint32_t)(sk[j1] >> 32) != (uint32_t)(sk[j1 - 1] >> 32)) ? 1u : 0u;
    }
    const uint32_t mine = f0 + f1;
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wsum[lane];
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += v;
        }
        wsum[lane] = winc - w;
        if (lane == 31) nv_s = winc;
    }
    __syncthreads();
    uint32_t pre = inc - mine + wsum[warp];
    if (f0) start[pre] = 2 * tid;
    pre += f0;
    if (f1) start[pre] = 2 * tid + 1;
    __syncthreads();
    const uint32_t nv = nv_s;
    for (uint32_t v = tid; v < nv; v += kVgSmallThreads) {
        const uint32_t b = start[v], e = v + 1 < nv ? start[v + 1] : n;
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        for (uint32_t j = b; j < e; ++j) {
            const float4 p = __ldg(pts + (uint32_t)sk[j]);
            sx += p.x; sy += p.y; sz += p.z; si += p.w;
        }
        const float cnt = (float)(e - b);
        out[v] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
        if (out_keys) out_keys[v] = (uint32_t)(sk[b] >> 32);
    }
    if (tid == 0) { info->nvox = nv; info->passthrough = 0; }
}

}  // namespace lvreg
