// register_warm.cuh -- the fused scan-to-map registration loop (scan2MapOptimization MO:1315-1343), thread per
// query, with two changes against register_tpq_kernel that leave every result bit-identical:
//
//   * WARM-STARTED SEARCH RADIUS.  The 5 neighbours a query found in the previous LM iteration are still
//     5 map points after the pose update, so the largest of their 5 distances to the moved query is an exact
//     upper bound of its new 5th-neighbour distance.  The search (MO:1019, MO:1111) starts with that radius
//     instead of the 1 m gate: rows and cells outside it are never loaded and only candidates inside it reach
//     the top-5 insertion -- the divergent part of the scan (ncu: a quarter of all issued instructions at 7
//     active lanes).  The candidate set is a superset of the true 5-NN (ties at the radius included, the
//     (d2, index) key order decides as before), so neighbours, coefficients and poses do not change.
//   * STATIC TILES.  Tile t runs on block t % gridDim, warp t / gridDim, every iteration; warps add their
//     tiles in tile order, blocks their warps in warp order, the grid its blocks in block order.  The fp64 sums
//     of MO:1257-1259 therefore have one fixed order: results are bit-reproducible run to run.  A tile is 32
//     consecutive queries of the Morton-ordered scan (neighbours in space: similar paths through the search,
//     shared cache lines).  Dealing the queries out instead (lane l of tile t takes query l * tiles + t, every
//     tile a uniform sample of the scan; RegArgs::dealt) balances the tiles perfectly but every tile then
//     costs as much as the slowest one did: measured 0.46 ms against 0.42 ms per C3 registration.
//   The row loop is rolled (one copy of the scan + insertion code instead of nine) to keep the hot loop inside
//   the instruction cache, and a candidate batch is tested against the radius in fp32 before any 64-bit key is
//   formed.  Tried and dropped (measured on C3, results unchanged): draining the insertions of a whole row in
//   warp-aligned rounds from per-lane shared-memory stacks (0.444 ms against 0.425 ms: the stack traffic and
//   the extra warp reduction per row cost what the aligned insertions save).
#pragma once

#include <cooperative_groups.h>

#include "common.cuh"
#include "fit.cuh"
#include "knn.cuh"
#include "register.cuh"

namespace lvreg {

// largest float below a positive finite x: "d <= below(x)" is "d < x"
__device__ __forceinline__ float float_below(float x) { return __uint_as_float(__float_as_uint(x) - 1u); }

// GATED search with an initial radius.  tau_le: only candidates with d2 <= tau_le can be among the 5 nearest
// (the caller guarantees that, or passes float_below(gate)).  Same visiting order, pruning rule and key order
// as thread_knn5_gated.
__device__ __forceinline__ void thread_knn5_radius(const GridView& g, float qx, float qy, float qz, float tau_le,
                                                   float gate_sq, u64 (&t)[5]) {
    if (g.hashed) {                 // hashed directory (huge sparse extents): the plain gated search, still exact
        u64 hb[5];
        group_knn5_hashed<1>(g, qx, qy, qz, 0, 1u << (threadIdx.x & 31), false, gate_sq, hb);
#pragma unroll
        for (int i = 0; i < 5; ++i) t[i] = hb[i];
        return;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) t[i] = kKeyNone;
    float ux, uy, uz;
    const int cx = cell_coord(qx, g.ox, g.inv, g.dx, &ux);
    const int cy = cell_coord(qy, g.oy, g.inv, g.dy, &uy);
    const int cz = cell_coord(qz, g.oz, g.inv, g.dz, &uz);
    const bool far = ux < -1.f || uy < -1.f || uz < -1.f || ux > (float)g.dx + 1.f ||
                     uy > (float)g.dy + 1.f || uz > (float)g.dz + 1.f;
    if (far) return;
    const float slx = 0.002f + 4e-7f * fabsf(ux), sly = 0.002f + 4e-7f * fabsf(uy), slz = 0.002f + 4e-7f * fabsf(uz);
    const float gxm = fmaxf(ux - (float)cx - slx, 0.f) * g.cell, gxp = fmaxf((float)(cx + 1) - ux - slx, 0.f) * g.cell;
    const float gym = fmaxf(uy - (float)cy - sly, 0.f) * g.cell, gyp = fmaxf((float)(cy + 1) - uy - sly, 0.f) * g.cell;
    const float gzm = fmaxf(uz - (float)cz - slz, 0.f) * g.cell, gzp = fmaxf((float)(cz + 1) - uz - slz, 0.f) * g.cell;
    const float gxm2 = gxm * gxm, gxp2 = gxp * gxp;
    const int xl = cx > 0 ? cx - 1 : cx, xh = cx + 1 < g.dx ? cx + 2 : cx + 1;
    // (dy, dz) visiting order of thread_knn5_gated: centre, faces, diagonals; two bits per entry, value + 1
    const uint32_t ody = 1u | (0u << 2) | (2u << 4) | (1u << 6) | (1u << 8) | (0u << 10) | (2u << 12) | (0u << 14) | (2u << 16);
    const uint32_t odz = 1u | (1u << 2) | (1u << 4) | (0u << 6) | (2u << 8) | (0u << 10) | (0u << 12) | (2u << 14) | (2u << 16);
    const float inf = __int_as_float(0x7f800000);
#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
        const int dyy = (int)((ody >> (2 * r)) & 3u) - 1, dzz = (int)((odz >> (2 * r)) & 3u) - 1;
        const int yy = cy + dyy, zz = cz + dzz;
        if (yy < 0 || yy >= g.dy || zz < 0 || zz >= g.dz) continue;
        const float gy = dyy == 0 ? 0.f : (dyy < 0 ? gym : gyp);
        const float gz = dzz == 0 ? 0.f : (dzz < 0 ? gzm : gzp);
        const float rb = gy * gy + gz * gz;
        if (rb > tau_le) continue;
        const uint32_t* row = g.cell_start + ((uint32_t)zz * g.dy + yy) * g.dx;
        const uint32_t s = __ldg(row + ((rb + gxm2 > tau_le) ? cx : xl));
        const uint32_t e = __ldg(row + ((rb + gxp2 > tau_le) ? cx + 1 : xh));
        for (uint32_t c = s; c < e; c += 4) {
            const uint32_t last = e - 1;
            const float4 p0 = __ldg(g.pts + c);
            const float4 p1 = __ldg(g.pts + min(c + 1, last));
            const float4 p2 = __ldg(g.pts + min(c + 2, last));
            const float4 p3 = __ldg(g.pts + min(c + 3, last));
            const float d0 = sqdist(qx, qy, qz, p0.x, p0.y, p0.z);
            const float d1 = c + 1 < e ? sqdist(qx, qy, qz, p1.x, p1.y, p1.z) : inf;
            const float d2 = c + 2 < e ? sqdist(qx, qy, qz, p2.x, p2.y, p2.z) : inf;
            const float d3 = c + 3 < e ? sqdist(qx, qy, qz, p3.x, p3.y, p3.z) : inf;
            if (fminf(fminf(d0, d1), fminf(d2, d3)) <= tau_le) {
                if (d0 <= tau_le) top5_insert(t, make_key(d0, p0.w));
                if (d1 <= tau_le) top5_insert(t, make_key(d1, p1.w));
                if (d2 <= tau_le) top5_insert(t, make_key(d2, p2.w));
                if (d3 <= tau_le) top5_insert(t, make_key(d3, p3.w));
                if (t[4] != kKeyNone) tau_le = fminf(tau_le, key_d2(t[4]));
            }
        }
    }
}

constexpr size_t register_warm_smem_bytes() { return (size_t)kRegWarps * 32 * 9 * sizeof(float); }

// TILE queries per warp tile (lanes >= TILE idle during search and fit), MINB resident blocks per SM (register cap)
template <int TILE, int MINB>
__global__ void __launch_bounds__(kRegThreads, MINB) register_warm_kernel_t(RegArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ Affine sT;
    __shared__ Trig sTrig;
    __shared__ float sPose[6];
    __shared__ float sRow[kRegWarps][TILE][9];
    __shared__ double sRed[kRegWarps][kRegTerms];
    __shared__ double sSum[kRegTerms];
    __shared__ int sStop;
    __shared__ LmState sLm;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const RegParams P = a.prm;
    const uint32_t tiles_c = (a.n[0] + TILE - 1) / TILE, tiles_s = (a.n[1] + TILE - 1) / TILE;
    const uint32_t tiles = tiles_c + tiles_s;
    // consecutive tiles go to different blocks: neighbouring (similarly expensive) tiles spread over all SMs
    const uint32_t first_tile = (uint32_t)warp * gridDim.x + blockIdx.x;
    const uint32_t tile_stride = gridDim.x * kRegWarps;
    const float gate_le = float_below(P.knn_gate_sq);

    if (threadIdx.x < 6) sPose[threadIdx.x] = a.pose_in[threadIdx.x];
    if (threadIdx.x == 0) { sLm = *a.lm; sStop = 0; }
    __syncthreads();

    int ti = 0, tj = 0;
    if (lane < 28) term_pair(lane, &ti, &tj);

    int iter = 0;
    int converged = 0;
    for (; iter < P.max_iters; ++iter) {
        if (blockIdx.x == 0 && threadIdx.x == 0) a.out->stamp[iter][0] = gtimer();
        if (warp == 0) pose_to_affine_warp(sPose, &sT, &sTrig);
        __syncthreads();
        const Affine T = sT;
        const Trig trig = sTrig;
        double acc = 0.0;

        for (uint32_t tile = first_tile; tile < tiles; tile += tile_stride) {
            const unsigned long long tile_t0 = (a.tile_ns && iter == 1) ? gtimer() : 0ull;
            const int cls = tile < tiles_c ? 0 : 1;
            const uint32_t tl = cls == 0 ? tile : tile - tiles_c, tcls = cls == 0 ? tiles_c : tiles_s;
            const uint32_t qi = a.dealt ? (uint32_t)lane * tcls + tl : tl * TILE + lane;
            const uint32_t ncls = tcls * TILE;                        // slots of the neighbour cache
            const bool valid = lane < TILE && qi < a.n[cls];
            const float4* __restrict__ map = a.map[cls];
            int32_t* __restrict__ nnp = a.nn_prev[cls] + (size_t)tl * TILE + lane - qi;   // [5][tiles * 32], tile-major
            float4 ori = valid ? __ldg(a.scan[cls] + qi) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float3 sel = apply_affine(T, ori.x, ori.y, ori.z);
            float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            bool ok = false;
            // search radius: the previous iteration's 5 neighbours bound the new 5th distance
            float tau_le = gate_le;
            if (valid && iter > 0) {
                int pn[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) pn[i] = nnp[(size_t)i * ncls + qi];
                if (pn[4] >= 0) {
                    float dm = 0.f;
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const float4 p = __ldg(map + pn[i]);
                        dm = fmaxf(dm, sqdist(sel.x, sel.y, sel.z, p.x, p.y, p.z));
                    }
                    tau_le = fminf(tau_le, dm);
                }
            }
            if (valid) {
                u64 best[5];
                thread_knn5_radius(a.grid[cls], sel.x, sel.y, sel.z, tau_le, P.knn_gate_sq, best);
                int nn[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    nn[i] = key_idx(best[i]);
                    nnp[(size_t)i * ncls + qi] = nn[i];
                }
                float4 coeff;
                ok = fit_query(cls, map, nn, key_d2(best[4]), ori, sel, P, &coeff);
                if (ok) jacobian_row(trig, ori.x, ori.y, ori.z, coeff, row);
            }
            if (lane < TILE) {
#pragma unroll
                for (int i = 0; i < 7; ++i) sRow[warp][lane][i] = row[i];
                sRow[warp][lane][7] = ok ? 1.0f : 0.0f;
            }
            __syncwarp();
            if (lane < kRegTerms) {
                if (lane < 28) {
#pragma unroll 8
                    for (int r = 0; r < TILE; ++r) acc += (double)sRow[warp][r][ti] * (double)sRow[warp][r][tj];
                } else {
#pragma unroll 8
                    for (int r = 0; r < TILE; ++r) acc += (double)sRow[warp][r][7];
                }
            }
            __syncwarp();
            if (a.tile_ns && iter == 1 && lane == 0) a.tile_ns[tile] = (uint32_t)(gtimer() - tile_t0);
        }

        // block partial in warp order, grid total in block order (fixed order end to end)
        if (lane < kRegTerms) sRed[warp][lane] = acc;
        __syncthreads();
        double* part = a.partials + ((size_t)(iter & 1) * gridDim.x + blockIdx.x) * kRegTerms;
        if (threadIdx.x < kRegTerms) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kRegWarps; ++w) s += sRed[w][threadIdx.x];
            part[threadIdx.x] = s;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) a.out->stamp[iter][1] = gtimer();
        grid.sync();
        if (blockIdx.x == 0 && threadIdx.x == 0) a.out->stamp[iter][2] = gtimer();
        {
            const double* all = a.partials + (size_t)(iter & 1) * gridDim.x * kRegTerms;
            // 8 chains per term, each with 4 independent accumulators (loads in flight instead of one dependent
            // add per L2 round trip); every partial keeps a fixed place in a fixed order
            const int t = threadIdx.x % 32, chain = threadIdx.x / 32;
            double s = 0.0;
            if (t < kRegTerms) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                uint32_t b = chain;
                for (; b + 3 * kRegWarps < gridDim.x; b += 4 * kRegWarps) {
                    s0 += all[(size_t)b * kRegTerms + t];
                    s1 += all[(size_t)(b + kRegWarps) * kRegTerms + t];
                    s2 += all[(size_t)(b + 2 * kRegWarps) * kRegTerms + t];
                    s3 += all[(size_t)(b + 3 * kRegWarps) * kRegTerms + t];
                }
                for (; b < gridDim.x; b += kRegWarps) s0 += all[(size_t)b * kRegTerms + t];
                s = (s0 + s1) + (s2 + s3);
            }
            __syncthreads();
            if (t < kRegTerms) sRed[chain][t] = s;
            __syncthreads();
            if (threadIdx.x < kRegTerms) {
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < kRegWarps; ++w) tot += sRed[w][threadIdx.x];
                sSum[threadIdx.x] = tot;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            if (blockIdx.x == 0) a.out->stamp[iter][3] = gtimer();
            const int n_sel = (int)(sSum[28] + 0.5);
            int conv = 0;
            if (n_sel >= P.min_matches) {
                float AtA[36], Atb[6];
                int t = 0;
#pragma unroll
                for (int i = 0; i < 7; ++i)
#pragma unroll
                    for (int j = i; j < 7; ++j, ++t) {
                        if (j < 6) { AtA[i * 6 + j] = (float)sSum[t]; AtA[j * 6 + i] = (float)sSum[t]; }
                        else if (i < 6) Atb[i] = (float)sSum[t];
                    }
                conv = lm_solve(AtA, Atb, iter, sPose, &sLm, P, nullptr) ? 1 : 0;
            }
            // n_sel < min_matches: LMOptimization returns false without touching the pose (MO:1209-1212); the
            // remaining iterations would repeat the same work -> stop, reporting max_iters like the reference
            sStop = conv ? 1 : (n_sel < P.min_matches ? 2 : 0);
            if (blockIdx.x == 0) {
                a.out->n_sel[iter] = n_sel;
                a.out->cost[iter] = (float)sSum[27];
                for (int i = 0; i < 6; ++i) a.out->pose_iter[iter][i] = sPose[i];
                a.out->stamp[iter][4] = gtimer();
            }
        }
        __syncthreads();
        const int stop = sStop;
        if (stop == 1) { converged = 1; ++iter; break; }
        if (stop == 2) {
            if (blockIdx.x == 0 && threadIdx.x == 0)
                for (int k = iter + 1; k < P.max_iters; ++k) {
                    a.out->n_sel[k] = a.out->n_sel[iter];
                    a.out->cost[k] = a.out->cost[iter];
                    for (int i = 0; i < 6; ++i) a.out->pose_iter[k][i] = sPose[i];
                }
            iter = P.max_iters;
            break;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.out->iterations = iter;
        a.out->converged = converged;
        a.out->degenerate = sLm.is_degenerate;
        a.out->pad = sLm.last_path;
        for (int i = 0; i < 6; ++i) a.out->pose[i] = sPose[i];
        *a.lm = sLm;
    }
}

// Tried and dropped: 16-query tiles at 4 resident blocks per SM (64 registers, twice the warps per scheduler):
// 0.58 ms against 0.44 ms per C3 registration -- the fits spill (1 KB of spill traffic per thread).
inline const void* register_warm_kernel_ptr(int) { return (const void*)register_warm_kernel_t<32, 2>; }

}  // namespace lvreg
