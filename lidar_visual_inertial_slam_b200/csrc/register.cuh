// register.cuh -- the fused scan-to-map registration: scan2MapOptimization's loop
// (MO:1325-1337) as ONE cooperative, persistent kernel launch.
//
// Per iteration, with no host round trip:
//   pointAssociateToMap (MO:339-345)      sensor-frame feature -> map frame, in registers
//   nearestKSearch(5)   (MO:1019, 1111)   grid search, LPQ lanes per query (knn.cuh)
//   corner / surf fit   (MO:1025-1092, 1121-1163)  one lane per query, registers (fit.cuh)
//   combineOptimizationCoeffs (MO:1169-1188)  disappears: flags feed the reduction directly
//   matAt*matA, matAt*matB (MO:1257-1259) 28 unique products of [J | r] + the match count,
//                                         fp64 accumulators (cv::gemm uses double for CV_32F),
//                                         fixed-order warp -> block -> grid reduction
//   cv::solve, degeneracy, pose update, convergence (MO:1260-1311)   every block, redundantly,
//                                         after one grid.sync(): no second barrier, no broadcast
//
// Work decomposition: a warp owns a tile of 32 consecutive queries.  Phase A: the warp's
// 32/LPQ lane groups each search LPQ of the tile's queries (all lanes busy on candidate scans).
// Phase B: lane l fits query l (all lanes busy on the 3x3 eigen / 5x3 QR).  Phase C: lanes 0..28
// each own one of the 29 reduction terms and add the tile's 32 rows in row order from shared
// memory, so the accumulation order is fixed (bit-reproducible run to run).
#pragma once

#include <cooperative_groups.h>

#include "common.cuh"
#include "fit.cuh"
#include "knn.cuh"

namespace lvreg {

namespace cg = cooperative_groups;

constexpr int kRegThreads = 256;
constexpr int kRegWarps = kRegThreads / 32;
constexpr int kRegTerms = 29;            // 28 products of [J0..J5, r] (i <= j) + match count

struct RegOut {                          // device mirror of lvreg_result's dynamic part
    int iterations, converged, degenerate, pad;
    int n_sel[32];
    float pose_iter[32][6];
    float cost[32];
    float pose[6];                       // final transformTobeMapped
    // block 0's phase timestamps per iteration (globaltimer ns): start, tiles done, barrier
    // passed, reduction done, solve done -- diagnostics for lvreg_get_iteration_profile
    unsigned long long stamp[32][5];
};

struct RegArgs {
    GridView grid[2];                    // corner / surf search grids
    const float4* map[2];                // laserCloud{Corner,Surf}FromMapDS (VoxelGrid order)
    const float4* scan[2];               // laserCloud{Corner,Surf}LastDS (sensor frame)
    uint32_t n[2];
    RegParams prm;
    const float* pose_in;                // transformTobeMapped on entry
    double* partials;                    // [2][gridDim.x][kRegTerms]
    uint32_t* tile_counter;              // [LVREG max iters] zeroed by the host before the launch
    uint32_t* tile_ns;                   // optional diagnostics: duration of every tile in iteration 1 (ns)
    int dealt;                           // register_warm_kernel: deal the queries out over the tiles (experiment)
    int32_t* nn_prev[2];                 // register_warm_kernel: the 5 neighbours of the previous iteration, [5][n]
    uint32_t* stage_stats;               // optional diagnostics of register_staged_kernel: tiles staged whole / as
                                         // halves / as quarters, lanes on the global-memory search, barrier time-outs
    RegOut* out;
    LmState* lm;
};

// term t (0..27) -> the (i, j) pair, i <= j over 7 columns
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void term_pair(int t, int* i, int* j) {
    int a = 0, rem = t;
    while (rem >= 7 - a) { rem -= 7 - a; ++a; }
    *i = a;
    *j = a + rem;
}

// neighbour gather + fit for one query.  Returns the acceptance flag and fills coeff.
__device__ __forceinline__ bool fit_query(int cls, const float4* __restrict__ map, const int (&nn)[5],
                                          float d5, float4 ori, float3 sel, const RegParams& P,
                                          float4* coeff) {
    *coeff = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!((double)d5 < (double)P.knn_gate_sq)) return false;
    float nbx[5], nby[5], nbz[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        float4 p = __ldg(map + nn[j]);
        nbx[j] = p.x; nby[j] = p.y; nbz[j] = p.z;
    }
    float4 c;
    bool ok = (cls == 0) ? corner_residual(nbx, nby, nbz, sel.x, sel.y, sel.z, P, &c)
                         : surf_residual(nbx, nby, nbz, ori.x, ori.y, ori.z, sel.x, sel.y, sel.z, P, &c);
    if (ok) *coeff = c;
    return ok;
}

template <int LPQ, int TILE>
__global__ void __launch_bounds__(kRegThreads, 2) register_kernel(RegArgs a) {
    constexpr int GROUPS = 32 / LPQ;                 // lane groups per warp
    constexpr int QPG = TILE / GROUPS;               // queries searched by one group per tile
    static_assert(TILE % GROUPS == 0 && QPG >= 1, "tile / group mismatch");
    cg::grid_group grid = cg::this_grid();
    __shared__ Affine sT;
    __shared__ Trig sTrig;
    __shared__ float sPose[6];
    __shared__ int sNN[kRegWarps][TILE][5];
    __shared__ float sD5[kRegWarps][TILE];
    __shared__ float sRow[kRegWarps][TILE][9];
    __shared__ double sRed[kRegWarps][kRegTerms];
    __shared__ double sSum[kRegTerms];
    __shared__ int sStop;
    __shared__ LmState sLm;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % LPQ, grp = lane / LPQ;
    const unsigned gmask = (LPQ == 32) ? 0xffffffffu : (((1u << LPQ) - 1u) << (lane - gl));
    const RegParams P = a.prm;
    const uint32_t tiles_c = (a.n[0] + TILE - 1) / TILE, tiles_s = (a.n[1] + TILE - 1) / TILE;
    const uint32_t tiles = tiles_c + tiles_s;

    if (threadIdx.x < 6) sPose[threadIdx.x] = a.pose_in[threadIdx.x];
    if (threadIdx.x == 0) { sLm = *a.lm; sStop = 0; }
    __syncthreads();

    int ti = 0, tj = 0;
    if (lane < 28) term_pair(lane, &ti, &tj);

    int iter = 0;
    int converged = 0;
    for (; iter < P.max_iters; ++iter) {
        if (threadIdx.x == 0) {
            if (blockIdx.x == 0) a.out->stamp[iter][0] = gtimer();
            pose_to_affine_dev(sPose, &sT, &sTrig);
        }
        __syncthreads();
        const Affine T = sT;
        const Trig trig = sTrig;
        double acc = 0.0;                                   // this lane's term, over the warp's tiles

        // dynamic tile scheduler: corner tiles (denser cells, costlier) are handed out first
        for (;;) {
            uint32_t tile = 0;
            if (lane == 0) tile = atomicAdd(a.tile_counter + iter, 1u);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile >= tiles) break;
            const int cls = tile < tiles_c ? 0 : 1;
            const uint32_t base = (cls == 0 ? tile : tile - tiles_c) * TILE;
            const uint32_t cnt = a.n[cls];
            const uint32_t qi = base + lane;
            const bool valid = lane < TILE && qi < cnt;
            float4 ori = valid ? __ldg(a.scan[cls] + qi) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float3 sel = apply_affine(T, ori.x, ori.y, ori.z);

            // phase A: each lane group searches QPG of the tile's queries
#pragma unroll 1
            for (int k = 0; k < QPG; ++k) {
                const int ql = grp * QPG + k;                // query (= lane) inside the tile
                const float qx = __shfl_sync(0xffffffffu, sel.x, ql);
                const float qy = __shfl_sync(0xffffffffu, sel.y, ql);
                const float qz = __shfl_sync(0xffffffffu, sel.z, ql);
                u64 best[5];
                if (base + ql < cnt) {
                    group_knn5_gated<LPQ>(a.grid[cls], qx, qy, qz, gl, gmask, P.knn_gate_sq, best);
                } else {
#pragma unroll
                    for (int i = 0; i < 5; ++i) best[i] = kKeyNone;
                }
                if (gl == 0) {
#pragma unroll
                    for (int i = 0; i < 5; ++i) sNN[warp][ql][i] = key_idx(best[i]);
                    sD5[warp][ql] = key_d2(best[4]);
                }
            }
            __syncwarp();

            // phase B: lane l fits query l
            if (lane < TILE) {
                int nn[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) nn[i] = sNN[warp][lane][i];
                float4 coeff;
                const bool ok = valid && fit_query(cls, a.map[cls], nn, sD5[warp][lane], ori, sel, P, &coeff);
                float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (ok) jacobian_row(trig, ori.x, ori.y, ori.z, coeff, row);
#pragma unroll
                for (int i = 0; i < 7; ++i) sRow[warp][lane][i] = row[i];
                sRow[warp][lane][7] = ok ? 1.0f : 0.0f;
            }
            __syncwarp();

            // phase C: lane t adds term t of the tile's rows, in row order
            if (lane < kRegTerms) {
                if (lane < 28) {
#pragma unroll 8
                    for (int r = 0; r < TILE; ++r)
                        acc += (double)sRow[warp][r][ti] * (double)sRow[warp][r][tj];
                } else {
#pragma unroll 8
                    for (int r = 0; r < TILE; ++r) acc += (double)sRow[warp][r][7];
                }
            }
            __syncwarp();
        }

        // block reduction (fixed order), then publish this block's partial sums
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

        // every block: grid reduction in block order, then the 6x6 solve (bit-identical everywhere)
        {
            const double* all = a.partials + (size_t)(iter & 1) * gridDim.x * kRegTerms;
            // 8 partial chains per term to shorten the dependent-add chain; combined in fixed order
            const int t = threadIdx.x % 32, chain = threadIdx.x / 32;
            double s = 0.0;
            if (t < kRegTerms)
                for (uint32_t b = chain; b < gridDim.x; b += kRegWarps) s += all[(size_t)b * kRegTerms + t];
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
            // n_sel < min_matches: LMOptimization returns false without touching the pose
            // (MO:1209-1212); the remaining iterations would repeat the same work -> stop,
            // reporting max_iters like the reference's loop counter.
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

// ---- thread-per-query variant ---------------------------------------------------------------------
// Same loop, but one lane owns one query end to end (search + fit in registers, no shared-memory
// hand-off, no shuffles in the search).  Preferred when there are enough queries to give every
// scheduler several warps; the grouped variant above wins for small scans.
__global__ void __launch_bounds__(kRegThreads, 2) register_tpq_kernel(RegArgs a) {
    constexpr int TILE = 32;
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

    if (threadIdx.x < 6) sPose[threadIdx.x] = a.pose_in[threadIdx.x];
    if (threadIdx.x == 0) { sLm = *a.lm; sStop = 0; }
    __syncthreads();

    int ti = 0, tj = 0;
    if (lane < 28) term_pair(lane, &ti, &tj);

    int iter = 0;
    int converged = 0;
    for (; iter < P.max_iters; ++iter) {
        if (threadIdx.x == 0) {
            if (blockIdx.x == 0) a.out->stamp[iter][0] = gtimer();
            pose_to_affine_dev(sPose, &sT, &sTrig);
        }
        __syncthreads();
        const Affine T = sT;
        const Trig trig = sTrig;
        double acc = 0.0;

        for (;;) {
            uint32_t tile = 0;
            if (lane == 0) tile = atomicAdd(a.tile_counter + iter, 1u);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile >= tiles) break;
            const unsigned long long tile_t0 = (a.tile_ns && iter == 1) ? gtimer() : 0ull;
            const int cls = tile < tiles_c ? 0 : 1;
            const uint32_t base = (cls == 0 ? tile : tile - tiles_c) * TILE;
            const uint32_t qi = base + lane;
            const bool valid = qi < a.n[cls];
            float4 ori = valid ? __ldg(a.scan[cls] + qi) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float3 sel = apply_affine(T, ori.x, ori.y, ori.z);
            float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            bool ok = false;
            if (valid) {
                u64 best[5];
                thread_knn5_gated(a.grid[cls], sel.x, sel.y, sel.z, P.knn_gate_sq, best);
                int nn[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) nn[i] = key_idx(best[i]);
                float4 coeff;
                ok = fit_query(cls, a.map[cls], nn, key_d2(best[4]), ori, sel, P, &coeff);
                if (ok) jacobian_row(trig, ori.x, ori.y, ori.z, coeff, row);
            }
#pragma unroll
            for (int i = 0; i < 7; ++i) sRow[warp][lane][i] = row[i];
            sRow[warp][lane][7] = ok ? 1.0f : 0.0f;
            __syncwarp();
            if (lane < kRegTerms) {
                if (lane < 28) {
#pragma unroll 8
                    for (int r = 0; r < TILE; ++r)
                        acc += (double)sRow[warp][r][ti] * (double)sRow[warp][r][tj];
                } else {
#pragma unroll 8
                    for (int r = 0; r < TILE; ++r) acc += (double)sRow[warp][r][7];
                }
            }
            __syncwarp();
            if (a.tile_ns && iter == 1 && lane == 0) a.tile_ns[tile] = (uint32_t)(gtimer() - tile_t0);
        }

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
            const int t = threadIdx.x % 32, chain = threadIdx.x / 32;
            double s = 0.0;
            if (t < kRegTerms)
                for (uint32_t b = chain; b < gridDim.x; b += kRegWarps) s += all[(size_t)b * kRegTerms + t];
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

// ---- stage-level: materialised cornerOptimization / surfOptimization ---------------------------
template <int LPQ>
__global__ void __launch_bounds__(kRegThreads) residual_kernel(GridView g, const float4* __restrict__ map,
                                                               const float4* __restrict__ scan, uint32_t n,
                                                               int cls, Affine T, RegParams P,
                                                               float4* __restrict__ coeff_out,
                                                               uint8_t* __restrict__ flag_out,
                                                               int32_t* __restrict__ nn_out) {
    __shared__ int sNN[kRegWarps][32][5];
    __shared__ float sD5[kRegWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % LPQ, grp = lane / LPQ;
    const unsigned gmask = (LPQ == 32) ? 0xffffffffu : (((1u << LPQ) - 1u) << (lane - gl));
    const uint32_t tiles = (n + 31) / 32;
    for (uint32_t tile = blockIdx.x * kRegWarps + warp; tile < tiles; tile += gridDim.x * kRegWarps) {
        const uint32_t base = tile * 32, qi = base + lane;
        const bool valid = qi < n;
        float4 ori = valid ? __ldg(scan + qi) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float3 sel = apply_affine(T, ori.x, ori.y, ori.z);
#pragma unroll 1
        for (int k = 0; k < LPQ; ++k) {
            const int ql = grp * LPQ + k;
            const float qx = __shfl_sync(0xffffffffu, sel.x, ql);
            const float qy = __shfl_sync(0xffffffffu, sel.y, ql);
            const float qz = __shfl_sync(0xffffffffu, sel.z, ql);
            u64 best[5];
            if (base + ql < n) {
                group_knn5_gated<LPQ>(g, qx, qy, qz, gl, gmask, P.knn_gate_sq, best);
            } else {
#pragma unroll
                for (int i = 0; i < 5; ++i) best[i] = kKeyNone;
            }
            if (gl == 0) {
#pragma unroll
                for (int i = 0; i < 5; ++i) sNN[warp][ql][i] = key_idx(best[i]);
                sD5[warp][ql] = key_d2(best[4]);
            }
        }
        __syncwarp();
        int nn[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) nn[i] = sNN[warp][lane][i];
        float4 coeff;
        const bool ok = valid && fit_query(cls, map, nn, sD5[warp][lane], ori, sel, P, &coeff);
        if (valid) {
            coeff_out[qi] = ok ? coeff : make_float4(0.f, 0.f, 0.f, 0.f);
            flag_out[qi] = ok ? 1 : 0;
            if (nn_out)
#pragma unroll
                for (int i = 0; i < 5; ++i) nn_out[(size_t)qi * 5 + i] = nn[i];
        }
        __syncwarp();
    }
}

// ---- stage-level: LMOptimization on explicit rows (lvreg_lm_step) ------------------------------
// single block; rows reduced in fixed order with fp64 accumulators
__global__ void __launch_bounds__(256) lm_step_kernel(const float4* __restrict__ ori,
                                                      const float4* __restrict__ coeff, uint32_t n,
                                                      int iter, Trig trig, RegParams P, float* pose,
                                                      LmState* lm, float* AtA_out, float* Atb_out,
                                                      float* x_out, int* conv_out) {
    __shared__ double sRed[8][kRegTerms];
    __shared__ float sRow[8][32][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ti = 0, tj = 0;
    if (lane < 28) term_pair(lane, &ti, &tj);
    double acc = 0.0;
    const uint32_t tiles = (n + 31) / 32;
    for (uint32_t tile = warp; tile < tiles; tile += 8) {
        const uint32_t qi = tile * 32 + lane;
        float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (qi < n) {
            float4 o = ori[qi];
            jacobian_row(trig, o.x, o.y, o.z, coeff[qi], row);
        }
#pragma unroll
        for (int i = 0; i < 7; ++i) sRow[warp][lane][i] = row[i];
        sRow[warp][lane][7] = qi < n ? 1.0f : 0.0f;
        __syncwarp();
        if (lane < 28) {
            for (int r = 0; r < 32; ++r) acc += (double)sRow[warp][r][ti] * (double)sRow[warp][r][tj];
        } else if (lane == 28) {
            for (int r = 0; r < 32; ++r) acc += (double)sRow[warp][r][7];
        }
        __syncwarp();
    }
    if (lane < kRegTerms) sRed[warp][lane] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum[kRegTerms];
        for (int t = 0; t < kRegTerms; ++t) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += sRed[w][t];
            sum[t] = s;
        }
        int conv = 0;
        if ((int)n >= P.min_matches) {
            float AtA[36], Atb[6];
            int t = 0;
            for (int i = 0; i < 7; ++i)
                for (int j = i; j < 7; ++j, ++t) {
                    if (j < 6) { AtA[i * 6 + j] = (float)sum[t]; AtA[j * 6 + i] = (float)sum[t]; }
                    else if (i < 6) Atb[i] = (float)sum[t];
                }
            LmState st = *lm;
            float p[6];
            for (int i = 0; i < 6; ++i) p[i] = pose[i];
            conv = lm_solve(AtA, Atb, iter, p, &st, P, x_out) ? 1 : 0;
            for (int i = 0; i < 6; ++i) pose[i] = p[i];
            *lm = st;
            for (int i = 0; i < 36; ++i) AtA_out[i] = AtA[i];
            for (int i = 0; i < 6; ++i) Atb_out[i] = Atb[i];
        }
        *conv_out = conv;
    }
}

}  // namespace lvreg
