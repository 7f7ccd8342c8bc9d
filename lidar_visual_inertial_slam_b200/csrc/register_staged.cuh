// register_staged.cuh -- the fused scan-to-map registration loop (scan2MapOptimization MO:1315-1343)
// with the 5-NN search (MO:1019, MO:1111) running out of SHARED MEMORY:
//
//   * a warp owns a tile of 32 consecutive queries.  The lanes' map cells span a small box of the
//     search grid; the box grown by one cell holds every candidate any lane can need (gate radius <=
//     cell edge).  Its rows are contiguous float4 runs of the cell-sorted map (x is the fastest
//     cell index), so the whole box is brought into the warp's shared-memory tile by ONE
//     cp.async.bulk per row, issued by one lane each and tracked by an mbarrier: no issue slots are
//     spent on the copy and every global load of the search is in flight at once (one latency
//     instead of one per row / candidate batch).
//   * the lanes then run the pruned 27-cell search of knn.cuh on the staged tile: rows nearest
//     first, tau = min(gate, current 5th distance), a row / x-cell is skipped when its slab gap
//     exceeds tau.  Keys are (bits(d2) << 32 | map index), so ties break on the smaller index
//     exactly as in the oracle.  Results are identical to thread_knn5_gated by construction (same
//     candidates in the same order, same arithmetic).
//   * a tile whose box is too large (rows, x extent or points) is retried as two halves, then four
//     quarters; what still does not fit takes the global-memory search.  Every path is exact.
//   * tiles are assigned to warps STATICALLY (tile t -> block t % gridDim, warp t / gridDim), each
//     warp adds its tiles in tile order, blocks add their warps in warp order and the grid adds the
//     blocks in block order: the fp64 accumulation order of the 28 + 1 normal-equation terms
//     (MO:1257-1259) is a function of the launch geometry only, so results are bit-reproducible.
//   * fit (MO:1025-1092, MO:1121-1163), Jacobian row (MO:1222-1255), solve / degeneracy / pose
//     update / convergence (MO:1260-1311) are shared with the other kernels (fit.cuh).
#pragma once

#include <cooperative_groups.h>

#include "common.cuh"
#include "fit.cuh"
#include "knn.cuh"
#include "register.cuh"

namespace lvreg {

constexpr int kStCap = 576;          // staged candidates per warp tile (float4 each)
constexpr int kStMaxNx = 16;         // cells per staged row
constexpr int kStMaxRows = 64;       // staged rows (two per lane)

struct StWarp {                                        // per-warp shared memory
    float4 cand[kStCap];                               // the staged box, row after row
    uint16_t tab[kStMaxRows][kStMaxNx + 2];            // cell boundaries relative to the row start
    uint16_t rowoff[kStMaxRows + 4];                   // first staged slot of each row
    float row[32][9];                                  // [J | r | flag] rows of the tile (phase C)
    unsigned long long mbar;                           // completion barrier of the bulk copies
};

constexpr size_t register_staged_smem_bytes() { return sizeof(StWarp) * kRegWarps; }

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void st_mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void st_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(bytes) : "memory");
}
// contiguous global -> shared copy by the TMA engine (SASS: UBLKCP); bytes % 16 == 0, both sides 16-B aligned
__device__ __forceinline__ void st_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(st_smem_u32(dst)), "l"(src), "r"(bytes), "r"(st_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool st_mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(st_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void st_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// the pruned 27-cell search of thread_knn5_gated on a staged box.  (x0, y0, z0) = first staged cell,
// ny = staged rows per z slab.
__device__ __forceinline__ void staged_knn5_gated(const GridView& g, const StWarp& S, int x0, int y0, int z0, int ny,
                                                  int cx, int cy, int cz, float ux, float uy, float uz, float qx, float qy,
                                                  float qz, float gate_sq, u64 (&t)[5]) {
    const float slx = 0.002f + 4e-7f * fabsf(ux), sly = 0.002f + 4e-7f * fabsf(uy), slz = 0.002f + 4e-7f * fabsf(uz);
    const float gxm = fmaxf(ux - (float)cx - slx, 0.f) * g.cell, gxp = fmaxf((float)(cx + 1) - ux - slx, 0.f) * g.cell;
    const float gym = fmaxf(uy - (float)cy - sly, 0.f) * g.cell, gyp = fmaxf((float)(cy + 1) - uy - sly, 0.f) * g.cell;
    const float gzm = fmaxf(uz - (float)cz - slz, 0.f) * g.cell, gzp = fmaxf((float)(cz + 1) - uz - slz, 0.f) * g.cell;
    const float gxm2 = gxm * gxm, gxp2 = gxp * gxp;
    const int ix = cx - x0;
    const int il = cx > 0 ? ix - 1 : ix, ih = cx + 1 < g.dx ? ix + 2 : ix + 1;
    // (dy, dz) visiting order of thread_knn5_gated: centre, faces, diagonals; two bits per entry, value + 1
    //   dy: 0 -1 1 0 0 -1 1 -1 1      dz: 0 0 0 -1 1 -1 -1 1 1
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
        const float tau = t[4] == kKeyNone ? gate_sq : key_d2(t[4]);
        if (rb > tau) continue;
        const int rr = (yy - y0) + (zz - z0) * ny;
        const uint16_t* tb = S.tab[rr];
        const uint32_t base = S.rowoff[rr];
        const uint32_t s = base + ((rb + gxm2 > tau) ? tb[ix] : tb[il]);
        const uint32_t e = base + ((rb + gxp2 > tau) ? tb[ix + 1] : tb[ih]);
        for (uint32_t c = s; c < e; c += 4) {
            const uint32_t last = e - 1;
            const float4 p0 = S.cand[c];
            const float4 p1 = S.cand[min(c + 1, last)];
            const float4 p2 = S.cand[min(c + 2, last)];
            const float4 p3 = S.cand[min(c + 3, last)];
            const float d0 = sqdist(qx, qy, qz, p0.x, p0.y, p0.z);
            const float d1 = c + 1 < e ? sqdist(qx, qy, qz, p1.x, p1.y, p1.z) : inf;
            const float d2 = c + 2 < e ? sqdist(qx, qy, qz, p2.x, p2.y, p2.z) : inf;
            const float d3 = c + 3 < e ? sqdist(qx, qy, qz, p3.x, p3.y, p3.z) : inf;
            if (d0 < gate_sq) top5_insert(t, make_key(d0, p0.w));
            if (d1 < gate_sq) top5_insert(t, make_key(d1, p1.w));
            if (d2 < gate_sq) top5_insert(t, make_key(d2, p2.w));
            if (d3 < gate_sq) top5_insert(t, make_key(d3, p3.w));
        }
    }
}

// the global-memory search, out of line: it only serves tiles that could not be staged
__device__ __noinline__ void fallback_knn5_gated(const GridView& g, float qx, float qy, float qz, float gate_sq, u64* out) {
    u64 t[5];
    thread_knn5_gated(g, qx, qy, qz, gate_sq, t);
#pragma unroll
    for (int i = 0; i < 5; ++i) out[i] = t[i];
}

// Stages the box of the lanes in `seg` and searches it.  Warp-uniform return value: false = the box
// does not fit (nothing was changed).  All 32 lanes must call.
__device__ __forceinline__ bool stage_and_search(const GridView& g, StWarp& S, uint32_t& parity, unsigned seg, int lane,
                                                 int cx, int cy, int cz, float ux, float uy, float uz, float qx, float qy,
                                                 float qz, float gate_sq, u64 (&best)[5], int* err) {
    const unsigned full = 0xffffffffu;
    if (g.hashed) return false;                                // hashed directory: rows are not contiguous runs
    const bool in = (seg >> lane) & 1u;
    const int big = 0x7fffffff;
    const int xmin = __reduce_min_sync(full, in ? cx : big), xmax = __reduce_max_sync(full, in ? cx : -1);
    const int ymin = __reduce_min_sync(full, in ? cy : big), ymax = __reduce_max_sync(full, in ? cy : -1);
    const int zmin = __reduce_min_sync(full, in ? cz : big), zmax = __reduce_max_sync(full, in ? cz : -1);
    const int x0 = max(xmin - 1, 0), x1 = min(xmax + 1, g.dx - 1);
    const int y0 = max(ymin - 1, 0), y1 = min(ymax + 1, g.dy - 1);
    const int z0 = max(zmin - 1, 0), z1 = min(zmax + 1, g.dz - 1);
    const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
    const int rows = ny * nz;
    if (nx > kStMaxNx || rows > kStMaxRows) return false;

    // row table: lane r (and r + 32) fetches the nx + 1 cell boundaries of its row
    uint32_t gs[2] = {0, 0}, len[2] = {0, 0}, off[2] = {0, 0};
    uint32_t total = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int r = k * 32 + lane;
        if (k * 32 < rows) {                                   // warp-uniform
            if (r < rows) {
                const int yy = y0 + r % ny, zz = z0 + r / ny;
                const uint32_t* cs = g.cell_start + ((uint32_t)zz * g.dy + yy) * g.dx + x0;
                const uint32_t c0 = __ldg(cs);
                uint32_t ci = c0;
                S.tab[r][0] = 0;
                for (int i = 1; i <= nx; ++i) {
                    ci = __ldg(cs + i);
                    S.tab[r][i] = (uint16_t)min(ci - c0, 0xffffu);
                }
                gs[k] = c0;
                len[k] = ci - c0;
            }
            const uint32_t incl = warp_inclusive_scan(len[k], lane);
            off[k] = total + incl - len[k];
            total += __shfl_sync(full, incl, 31);
        }
    }
    if (total > (uint32_t)kStCap) return false;
    if (total == 0) return true;                               // empty box: nothing within the gate
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int r = k * 32 + lane;
        if (r < rows) S.rowoff[r] = (uint16_t)off[k];
    }
    // generic-proxy accesses of the tile by the previous tile / segment come before the async-proxy writes
    st_fence_proxy_async();
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (len[k]) st_bulk_g2s(&S.cand[off[k]], g.pts + gs[k], len[k] * 16u, &S.mbar);
    __syncwarp();
    if (lane == 0) st_mbar_expect_tx(&S.mbar, total * 16u);
    {
        int spins = 0;
        while (!st_mbar_try_wait(&S.mbar, parity)) {
            if (++spins > (1 << 22)) { *err = 1; break; }      // never observed; a hang would cost the whole box
        }
    }
    parity ^= 1u;
    if (in) staged_knn5_gated(g, S, x0, y0, z0, ny, cx, cy, cz, ux, uy, uz, qx, qy, qz, gate_sq, best);
    __syncwarp();
    return true;
}

// 5-NN of the tile's 32 queries (one per lane; `q` in the map frame).  Whole tile, else halves, else
// quarters from shared memory, else the global-memory search.  stats[0..3] (optional, lane 0 counts).
__device__ __forceinline__ void tile_search(const GridView& g, StWarp& S, uint32_t& parity, int lane, bool valid,
                                            float qx, float qy, float qz, float gate_sq, u64 (&best)[5], int* err,
                                            uint32_t* stats) {
    float ux, uy, uz;
    const int cx = cell_coord(qx, g.ox, g.inv, g.dx, &ux);
    const int cy = cell_coord(qy, g.oy, g.inv, g.dy, &uy);
    const int cz = cell_coord(qz, g.oz, g.inv, g.dz, &uz);
    const bool far = ux < -1.f || uy < -1.f || uz < -1.f || ux > (float)g.dx + 1.f ||
                     uy > (float)g.dy + 1.f || uz > (float)g.dz + 1.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) best[i] = kKeyNone;
    unsigned todo = __ballot_sync(0xffffffffu, valid && !far);
#pragma unroll 1
    for (int width = 32; todo != 0 && width >= 8; width >>= 1) {
        unsigned failed = 0;
#pragma unroll 1
        for (int lo = 0; lo < 32; lo += width) {
            const unsigned seg = (width == 32 ? 0xffffffffu : (((1u << width) - 1u) << lo)) & todo;
            if (seg == 0) continue;
            const bool ok = stage_and_search(g, S, parity, seg, lane, cx, cy, cz, ux, uy, uz, qx, qy, qz, gate_sq, best, err);
            if (!ok) failed |= seg;
            else if (stats) stats[width == 32 ? 0 : (width == 16 ? 1 : 2)] += 1;
        }
        todo = failed;
    }
    if (todo) {
        if ((todo >> lane) & 1u) {
            u64 fb[5];                                         // its address escapes; `best` stays in registers
            fallback_knn5_gated(g, qx, qy, qz, gate_sq, fb);
#pragma unroll
            for (int i = 0; i < 5; ++i) best[i] = fb[i];
        }
        if (stats) stats[3] += (uint32_t)__popc(todo);
        __syncwarp();
    }
}

// ---- stage-level kernel: materialised 5-NN through the shared-memory search (lvreg_knn5, variant STAGED) ----
__global__ void __launch_bounds__(kRegThreads, 2) knn5_staged_kernel(GridView g, const float4* __restrict__ queries,
                                                                     uint32_t nq, float gate_sq,
                                                                     int32_t* __restrict__ idx_out,
                                                                     float* __restrict__ d2_out, uint32_t* stage_stats) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    StWarp& S = reinterpret_cast<StWarp*>(st_smem)[warp];
    if (lane == 0) st_mbar_init(&S.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t parity = 0;
    int err = 0;
    uint32_t stats[4] = {0, 0, 0, 0};
    const uint32_t tiles = (nq + 31) / 32;
    for (uint32_t tile = (uint32_t)warp * gridDim.x + blockIdx.x; tile < tiles; tile += gridDim.x * kRegWarps) {
        const uint32_t q = tile * 32 + lane;
        const bool valid = q < nq;
        const float4 p = valid ? __ldg(queries + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        u64 best[5];
        tile_search(g, S, parity, lane, valid, p.x, p.y, p.z, gate_sq, best, &err, stage_stats ? stats : nullptr);
        if (valid) {
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                idx_out[(size_t)q * 5 + i] = key_idx(best[i]);
                d2_out[(size_t)q * 5 + i] = key_d2(best[i]);
            }
        }
    }
    if (stage_stats && lane == 0) {
        for (int i = 0; i < 4; ++i)
            if (stats[i]) atomicAdd(stage_stats + i, stats[i]);
        if (err) atomicAdd(stage_stats + 4, 1u);
    }
}

__global__ void __launch_bounds__(kRegThreads, 2) register_staged_kernel(RegArgs a) {
    constexpr int TILE = 32;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char st_smem[];
    __shared__ Affine sT;
    __shared__ Trig sTrig;
    __shared__ float sPose[6];
    __shared__ double sRed[kRegWarps][kRegTerms];
    __shared__ double sSum[kRegTerms];
    __shared__ int sStop;
    __shared__ LmState sLm;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    StWarp& S = reinterpret_cast<StWarp*>(st_smem)[warp];
    const RegParams P = a.prm;
    const uint32_t tiles_c = (a.n[0] + TILE - 1) / TILE, tiles_s = (a.n[1] + TILE - 1) / TILE;
    const uint32_t tiles = tiles_c + tiles_s;
    // consecutive tiles go to different blocks: the (denser) corner tiles spread over all SMs
    const uint32_t first_tile = (uint32_t)warp * gridDim.x + blockIdx.x;
    const uint32_t tile_stride = gridDim.x * kRegWarps;

    if (threadIdx.x < 6) sPose[threadIdx.x] = a.pose_in[threadIdx.x];
    if (threadIdx.x == 0) { sLm = *a.lm; sStop = 0; }
    if (lane == 0) st_mbar_init(&S.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t parity = 0;
    int stage_err = 0;
    uint32_t stats[4] = {0, 0, 0, 0};                        // diagnostics of iteration 0 (lane 0 counts)

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

        for (uint32_t tile = first_tile; tile < tiles; tile += tile_stride) {
            const unsigned long long tile_t0 = (a.tile_ns && iter == 1) ? gtimer() : 0ull;
            const int cls = tile < tiles_c ? 0 : 1;
            const uint32_t base = (cls == 0 ? tile : tile - tiles_c) * TILE;
            const uint32_t qi = base + lane;
            const bool valid = qi < a.n[cls];
            const GridView& g = a.grid[cls];
            float4 ori = valid ? __ldg(a.scan[cls] + qi) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float3 sel = apply_affine(T, ori.x, ori.y, ori.z);
            u64 best[5];
            tile_search(g, S, parity, lane, valid, sel.x, sel.y, sel.z, P.knn_gate_sq, best, &stage_err,
                        (a.stage_stats && iter == 0) ? stats : nullptr);

            // ---- fit + Jacobian row ----
            float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            bool ok = false;
            if (valid) {
                int nn[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) nn[i] = key_idx(best[i]);
                float4 coeff;
                ok = fit_query(cls, a.map[cls], nn, key_d2(best[4]), ori, sel, P, &coeff);
                if (ok) jacobian_row(trig, ori.x, ori.y, ori.z, coeff, row);
            }
#pragma unroll
            for (int i = 0; i < 7; ++i) S.row[lane][i] = row[i];
            S.row[lane][7] = ok ? 1.0f : 0.0f;
            __syncwarp();
            if (lane < kRegTerms) {
                if (lane < 28) {
#pragma unroll 8
                    for (int r = 0; r < TILE; ++r) acc += (double)S.row[r][ti] * (double)S.row[r][tj];
                } else {
#pragma unroll 8
                    for (int r = 0; r < TILE; ++r) acc += (double)S.row[r][7];
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
    if (a.stage_stats && lane == 0)
        for (int i = 0; i < 4; ++i)
            if (stats[i]) atomicAdd(a.stage_stats + i, stats[i]);
    if (a.stage_stats && stage_err) atomicAdd(a.stage_stats + 4, 1u);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.out->iterations = iter;
        a.out->converged = converged;
        a.out->degenerate = sLm.is_degenerate;
        a.out->pad = sLm.last_path;
        for (int i = 0; i < 6; ++i) a.out->pose[i] = sPose[i];
        *a.lm = sLm;
    }
}

}  // namespace lvreg
