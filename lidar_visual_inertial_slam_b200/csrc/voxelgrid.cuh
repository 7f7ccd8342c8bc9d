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
};

__global__ void __launch_bounds__(256) transform_concat_kernel(const Segment* __restrict__ segs,
                                                               uint32_t nseg, uint32_t total,
                                                               float4* __restrict__ out,
                                                               uint32_t* __restrict__ mm) {
    float mnx = 3.4e38f, mny = 3.4e38f, mnz = 3.4e38f, mxx = -3.4e38f, mxy = -3.4e38f, mxz = -3.4e38f;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        // last segment with begin <= i (empty segments share a begin and are skipped by the search)
        uint32_t lo = 0, hi = nseg;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (segs[mid].begin <= i) lo = mid; else hi = mid;
        }
        const Segment& s = segs[lo];
        float4 p = ld_stream(s.src + (i - s.begin));
        float3 q = apply_affine(s.T, p.x, p.y, p.z);
        out[i] = make_float4(q.x, q.y, q.z, p.w);
        mnx = fminf(mnx, q.x); mny = fminf(mny, q.y); mnz = fminf(mnz, q.z);
        mxx = fmaxf(mxx, q.x); mxy = fmaxf(mxy, q.y); mxz = fmaxf(mxz, q.z);
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

}  // namespace lvreg
