// voxelgrid_mid.cuh -- pcl::VoxelGrid<PointXYZI>::filter (MO:959-965, MO:991-997) for clouds of up to
// kVgMidMax points as ONE cooperative kernel launch.
//
// The multi-kernel pipeline of voxelgrid.cuh costs ~15 launches and two host round trips per cloud (bounding box ->
// voxel spec / pass count, voxel count -> output size).  For a MID360 scan (~20 k feature points) or the local map of
// an indoor sequence (~10^5 points) every one of those kernels runs for a few microseconds: the filter is bound by
// launch latency and by the host synchronisations, not by the device.  Here the whole filter -- (transform +
// concatenate,) bounding box, PCL's bounds / overflow rule, voxel keys, stable LSD radix sort, run heads, sequential
// centroids -- runs inside one grid of <= 148 blocks, with grid.sync() between the phases and every decision taken
// on the device.  Nothing returns to the host before the caller's single synchronisation at the end.
//
// One block owns one tile of 2048 points (256 threads x 8, warp w owns the 256 consecutive points w*256..): the
// tile's keys and ranks stay in registers across the barriers.  A radix pass needs no look-back chain: all tiles
// publish their digit counts, the grid synchronises, every tile sums the counts of its predecessors itself.
// Same fp32 expressions, same stable order, same sequential centroid sums as the multi-kernel path and the
// single-block path: the three are bit-identical (tests/test_gpu_parity.py).
#pragma once

#include <cooperative_groups.h>

#include "common.cuh"
#include "prims.cuh"
#include "voxelgrid.cuh"
#include "knn.cuh"

namespace lvreg {

constexpr int kVgMidThreads = 256;
constexpr int kVgMidItems = 8;
constexpr int kVgMidTile = kVgMidThreads * kVgMidItems;       // 2048
constexpr int kVgMidMaxTiles = 128;
constexpr int kVgMidMax = kVgMidTile * kVgMidMaxTiles;        // 262144 points

struct VgMidArgs {
    const float4* pts;          // flat input cloud, or
    const Segment* segs;        // keyframe segments to transform + concatenate (then `pts` is ignored)
    uint32_t nseg;
    uint32_t n;
    float leaf;
    float4* world;              // n: the transformed cloud (segments only)
    uint32_t *k0, *v0, *k1, *v1;   // n each: sort ping-pong
    uint32_t* tile_hist;        // [tiles][256] digit counts of the current pass, then [tiles] head counts
    float* tile_bb;             // [tiles][6] per-tile bounding boxes
    uint32_t* vstart;           // n: first sorted position of every voxel
    float4* out;                // n: centroids (only nvox are written)
    uint32_t* out_keys;         // optional: idx of every output voxel
    uint32_t* point_keys;       // optional: idx of every input point
    float4* morton_out;         // optional: the output cloud once more, ordered by the Morton code of its 2 m cell
                                // (the query order of the registration kernels, knn.cuh); n entries
    VgSmallInfo* info;
};

// how many bytes of scratch the hist / bbox arrays need
constexpr size_t vg_mid_hist_words() { return (size_t)kVgMidMaxTiles * 256; }
constexpr size_t vg_mid_bb_floats() { return (size_t)kVgMidMaxTiles * 6; }

__global__ void __launch_bounds__(kVgMidThreads) voxelgrid_mid_kernel(VgMidArgs a) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t wcnt[kVgMidThreads / 32][256];
    __shared__ float red[6][8];
    __shared__ VoxelSpec vs;
    __shared__ int pass_s;
    __shared__ float bb_s[6];
    __shared__ uint32_t sc_a, sc_b;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = a.n;
    const uint32_t tiles = (n + kVgMidTile - 1) / kVgMidTile;        // == gridDim.x
    const uint32_t tile = blockIdx.x;
    const uint32_t base = tile * kVgMidTile + warp * (32 * kVgMidItems);
    const float4* __restrict__ cloud = a.segs ? a.world : a.pts;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // ---- phase 0: (transform + concatenate,) bounding box of the tile ----
    float mnx = 3.4e38f, mny = 3.4e38f, mnz = 3.4e38f, mxx = -3.4e38f, mxy = -3.4e38f, mxz = -3.4e38f;
#pragma unroll
    for (int r = 0; r < kVgMidItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        if (i < n) {
            float4 p;
            if (a.segs) {
                uint32_t lo = 0, hi = a.nseg;
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (a.segs[mid].begin <= i) lo = mid; else hi = mid;
                }
                const Segment& s = a.segs[lo];
                const float4 q = ld_stream(s.src + (i - s.begin));
                const float3 w = apply_affine(s.T, q.x, q.y, q.z);
                p = make_float4(w.x, w.y, w.z, q.w);
                a.world[i] = p;
            } else {
                p = a.pts[i];
            }
            mnx = fminf(mnx, p.x); mny = fminf(mny, p.y); mnz = fminf(mnz, p.z);
            mxx = fmaxf(mxx, p.x); mxy = fmaxf(mxy, p.y); mxz = fmaxf(mxz, p.z);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o)); mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
    }
    if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mny; red[2][warp] = mnz; red[3][warp] = mxx; red[4][warp] = mxy; red[5][warp] = mxz; }
    __syncthreads();
    if (tid < 6) {
        float v = red[tid][0];
        for (int w = 1; w < kVgMidThreads / 32; ++w) v = tid < 3 ? fminf(v, red[tid][w]) : fmaxf(v, red[tid][w]);
        a.tile_bb[tile * 6 + tid] = v;
    }
    __threadfence();
    grid.sync();

    // ---- phase 1: every block derives the same voxel spec from the cloud's bounding box; keys ----
    if (tid < 6) {
        float v = a.tile_bb[tid];
        for (uint32_t t = 1; t < tiles; ++t) v = tid < 3 ? fminf(v, a.tile_bb[t * 6 + tid]) : fmaxf(v, a.tile_bb[t * 6 + tid]);
        bb_s[tid] = v;
    }
    __syncthreads();
    if (tid == 0) {
        // PCL voxel_grid.hpp: leaf-size overflow rule and bounds, fp32 exactly as PCL computes them
        const float inv = 1.0f / a.leaf;
        long long d[3];
        for (int k = 0; k < 3; ++k) d[k] = (long long)((bb_s[3 + k] - bb_s[k]) * inv) + 1;
        pass_s = d[0] * d[1] * d[2] > 2147483647ll ? 1 : 0;
        vs.inv = inv;
        int div_b[3];
        for (int k = 0; k < 3; ++k) {
            vs.min_b[k] = (int)floorf(bb_s[k] * inv);
            const int max_b = (int)floorf(bb_s[3 + k] * inv);
            div_b[k] = max_b - vs.min_b[k] + 1;
        }
        vs.mul[0] = 1; vs.mul[1] = div_b[0]; vs.mul[2] = div_b[0] * div_b[1];
        unsigned long long span = (unsigned long long)div_b[0] * (unsigned long long)div_b[1] * (unsigned long long)div_b[2];
        span = span ? span - 1 : 0;
        int bits = 1;
        while (bits < 32 && (span >> bits) != 0) ++bits;
        vs.key_bits = bits;
    }
    __syncthreads();
    if (pass_s) {                                 // PCL: "leaf size too small" -> the input, unchanged
#pragma unroll
        for (int r = 0; r < kVgMidItems; ++r) {
            const uint32_t i = base + r * 32 + lane;
            if (i < n) {
                a.out[i] = cloud[i];
                if (a.point_keys) a.point_keys[i] = 0;
            }
        }
        if (tile == 0 && tid == 0) {
            a.info->nvox = n; a.info->passthrough = 1;
            for (int k = 0; k < 3; ++k) { a.info->mn[k] = bb_s[k]; a.info->mx[k] = bb_s[3 + k]; }
        }
        return;
    }
    uint32_t k[kVgMidItems], v[kVgMidItems];
#pragma unroll
    for (int r = 0; r < kVgMidItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        k[r] = 0xffffffffu;                       // padding sorts last inside the (final) tile
        v[r] = i;
        if (i < n) {
            const float4 p = cloud[i];
            const int ix = (int)(floorf(p.x * vs.inv) - (float)vs.min_b[0]);
            const int iy = (int)(floorf(p.y * vs.inv) - (float)vs.min_b[1]);
            const int iz = (int)(floorf(p.z * vs.inv) - (float)vs.min_b[2]);
            k[r] = (uint32_t)(ix * vs.mul[0] + iy * vs.mul[1] + iz * vs.mul[2]);
            if (a.point_keys) a.point_keys[i] = k[r];
        }
    }

    // ---- stable LSD radix sort, 8 bits per pass, of the (k, v) pairs the tiles hold in registers ----
    uint32_t *kin = a.k0, *vin = a.v0, *kout = a.k1, *vout = a.v1;
    auto radix_sort = [&](uint32_t n_items, int passes) {
        for (int p = 0; p < passes; ++p) {
            const int shift = 8 * p;
            if (p > 0) {
#pragma unroll
                for (int r = 0; r < kVgMidItems; ++r) {
                    const uint32_t i = base + r * 32 + lane;
                    k[r] = i < n_items ? kin[i] : 0xffffffffu;
                    v[r] = i < n_items ? vin[i] : 0u;
                }
            }
            for (int i = tid; i < (kVgMidThreads / 32) * 256; i += kVgMidThreads) (&wcnt[0][0])[i] = 0;
            __syncthreads();
            // rank inside the warp's 256-point chunk: lanes with the same digit from 8 ballots, warp-private counters
            uint16_t rank[kVgMidItems];
#pragma unroll
            for (int r = 0; r < kVgMidItems; ++r) {
                const uint32_t dg = (k[r] >> shift) & 255u;
                uint32_t peers = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const bool bit = (dg & (1u << b)) != 0;
                    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                    peers &= bal ^ (bit ? 0u : 0xffffffffu);
                }
                const uint32_t lower = peers & lt_mask;
                const uint32_t old = wcnt[warp][dg];
                if (lower == 0) wcnt[warp][dg] = old + (uint32_t)__popc(peers);
                __syncwarp();
                rank[r] = (uint16_t)(old + __popc(lower));
            }
            __syncthreads();
            // the tile's digit counts (padding sits under digit 255 of every pass, behind every real element in index
            // order: it is counted, shifts nothing and is never written)
            uint32_t cnt = 0;
#pragma unroll
            for (int w = 0; w < kVgMidThreads / 32; ++w) cnt += wcnt[w][tid];
            a.tile_hist[tile * 256 + tid] = cnt;
            __threadfence();
            grid.sync();
            // global offset of this tile's run of digit `tid`: all smaller digits + the same digit in earlier tiles
            uint32_t excl = 0, tot = 0;
            for (uint32_t t = 0; t < tiles; ++t) {
                const uint32_t c = a.tile_hist[t * 256 + tid];
                tot += c;
                if (t < tile) excl += c;
            }
            uint32_t total;
            uint32_t gofs = block_exclusive_scan(tot, &total) + excl;
#pragma unroll
            for (int w = 0; w < kVgMidThreads / 32; ++w) {
                const uint32_t c = wcnt[w][tid];
                wcnt[w][tid] = gofs;
                gofs += c;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < kVgMidItems; ++r) {
                const uint32_t i = base + r * 32 + lane;
                if (i < n_items) {
                    const uint32_t dst = wcnt[warp][(k[r] >> shift) & 255u] + rank[r];
                    kout[dst] = k[r];
                    vout[dst] = v[r];
                }
            }
            __threadfence();
            grid.sync();
            uint32_t* t0 = kin; kin = kout; kout = t0;
            t0 = vin; vin = vout; vout = t0;
        }
    };
    radix_sort(n, (vs.key_bits + 7) / 8);
    // sorted pairs are in (kin, vin) -- for passes == 0 (cannot happen: key_bits >= 1) they would be in registers only

    // ---- run heads -> voxel starts ----
    uint32_t flag[kVgMidItems];
    uint32_t heads = 0;
#pragma unroll
    for (int r = 0; r < kVgMidItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        k[r] = i < n ? kin[i] : 0xffffffffu;
        flag[r] = 0;
        if (i < n) flag[r] = (i == 0 || kin[i - 1] != k[r]) ? 1u : 0u;
        heads += flag[r];
    }
    uint32_t total;
    block_exclusive_scan(heads, &total);
    if (tid == 0) a.tile_hist[tile] = total;
    __threadfence();
    grid.sync();
    if (tid == 0) {
        uint32_t before = 0, all = 0;
        for (uint32_t t = 0; t < tiles; ++t) {
            const uint32_t c = a.tile_hist[t];
            all += c;
            if (t < tile) before += c;
        }
        sc_a = before;
        sc_b = all;
    }
    // exclusive scan of the flags in index order = (warp chunk, round, lane)
    if (lane == 0) wcnt[0][warp] = 0;
    __syncthreads();
    uint32_t pos[kVgMidItems];
    uint32_t run = 0;
#pragma unroll
    for (int r = 0; r < kVgMidItems; ++r) {
        const uint32_t bal = __ballot_sync(0xffffffffu, flag[r] != 0);
        pos[r] = run + (uint32_t)__popc(bal & lt_mask);
        run += (uint32_t)__popc(bal);
    }
    if (lane == 0) wcnt[0][warp] = run;
    __syncthreads();
    uint32_t wpre = sc_a;
    for (int w = 0; w < warp; ++w) wpre += wcnt[0][w];
    const uint32_t nvox = sc_b;
#pragma unroll
    for (int r = 0; r < kVgMidItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        if (flag[r]) a.vstart[wpre + pos[r]] = i;
    }
    __threadfence();
    grid.sync();

    // ---- centroids: sequential fp32 sums in sorted (= input) order, true division ----
    for (uint32_t vx = blockIdx.x * kVgMidThreads + tid; vx < nvox; vx += gridDim.x * kVgMidThreads) {
        const uint32_t b = a.vstart[vx], e = vx + 1 < nvox ? a.vstart[vx + 1] : n;
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        for (uint32_t j = b; j < e; ++j) {
            const float4 p = cloud[vin[j]];
            sx += p.x; sy += p.y; sz += p.z; si += p.w;
        }
        const float c = (float)(e - b);
        a.out[vx] = make_float4(sx / c, sy / c, sz / c, si / c);
        if (a.out_keys) a.out_keys[vx] = kin[b];
    }
    if (tile == 0 && tid == 0) {
        a.info->nvox = nvox; a.info->passthrough = 0;
        for (int q = 0; q < 3; ++q) { a.info->mn[q] = bb_s[q]; a.info->mx[q] = bb_s[3 + q]; }
    }
    if (a.morton_out == nullptr) return;

    // ---- the output once more in Morton order of its 2 m cells (same keys as morton_keys_kernel, stable) ----
    __threadfence();
    grid.sync();
#pragma unroll
    for (int r = 0; r < kVgMidItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        k[r] = 0xffffffffu;
        v[r] = i;
        if (i < nvox) {
            const float4 p = a.out[i];
            k[r] = morton_cell_key(p.x, p.y, p.z);
        }
    }
    kin = a.k0; vin = a.v0; kout = a.k1; vout = a.v1;
    radix_sort(nvox, 3);
#pragma unroll
    for (int r = 0; r < kVgMidItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        if (i < nvox) a.morton_out[i] = a.out[vin[i]];
    }
}

}  // namespace lvreg
