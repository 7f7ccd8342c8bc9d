// knn.cuh -- exact 5-nearest-neighbour search over a uniform cell grid in HBM; replaces
// pcl::KdTreeFLANN::setInputCloud (MO:1322-1323) and nearestKSearch(pointSel, 5, ...)
// (MO:1019, MO:1111).
//
// Structure (built once per local map):
//   cell_pts[m]        float4 {x, y, z, bits(original index)} sorted by cell key (x fastest), so
//                      the 3 x-adjacent cells of a row are ONE contiguous, coalesced float4 run
//   cell_start[nc+1]   exclusive prefix of the per-cell counts: a DENSE directory while it is small next to
//                      the map (cells <= max(2^24, 16 M)), else an open-addressing HASH of the occupied cells
//                      (memory proportional to the map; a sparse map spread over kilometres keeps the 1 m cell)
//   cell edge          gate radius * (1 + 1/128): any point closer than the gate (sqrt(1.0 m^2),
//                      MO:1025/1121) lies in the 3x3x3 block around the query's cell, with margin
//                      for the fp32 rounding of the cell coordinate (dims are capped at 2048/axis)
//
// Search: a group of LPQ lanes owns one query.  Lanes fetch the 9 row ranges in parallel, then
// stride over each row's points (coalesced), each lane keeping a sorted top-5 of 64-bit keys
// (bits(d2) << 32 | index): unsigned order on that key IS the (d2, index) lexicographic order of
// the oracle, so ties break deterministically on the smaller index.  Lane lists are merged with
// 5 rounds of xor-shuffle minimum.  d2 = ((dx*dx)+dy*dy)+dz*dz without FMA (L2_Simple).
//
// GATED mode stops after the 3x3x3 block: exact whenever d2[4] < gate (the only case the
// reference uses).  EXACT mode keeps adding Chebyshev shells until the 5th distance is provably
// inside the searched cube (or the whole grid was visited).
#pragma once

#include <cooperative_groups.h>

#include "common.cuh"
#include "prims.cuh"

namespace lvreg {

typedef unsigned long long u64;
constexpr u64 kKeyNone = 0xffffffffffffffffull;

struct GridView {
    const float4* pts;           // cell-sorted points, w = original index bits
    const uint32_t* cell_start;  // dense: ncells + 1 prefix; hashed: hmask + 2 prefix over the table slots
    const unsigned long long* hkeys;   // hashed: linear cell id of every table slot (kHashEmpty = free)
    uint32_t hmask;              // hashed: table size - 1 (power of two)
    int hashed;                  // 0 = dense cell directory, 1 = open-addressing hash of the occupied cells
    float ox, oy, oz;            // grid origin (bbox min of the map)
    float inv;                   // 1 / cell
    float cell;
    int dx, dy, dz;              // cells per axis
    uint32_t m;                  // number of map points
};

__device__ __forceinline__ u64 make_key(float d2, float idx_bits) {
    return ((u64)__float_as_uint(d2) << 32) | (u64)__float_as_uint(idx_bits);
}
__device__ __forceinline__ float key_d2(u64 k) {
    return k == kKeyNone ? __int_as_float(0x7f800000) : __uint_as_float((uint32_t)(k >> 32));
}
__device__ __forceinline__ int key_idx(u64 k) { return k == kKeyNone ? -1 : (int)(uint32_t)k; }

__device__ __forceinline__ void top5_insert(u64 (&t)[5], u64 key) {
    if (key < t[4]) {
        t[4] = key;
#pragma unroll
        for (int i = 4; i > 0; --i) {
            if (t[i] < t[i - 1]) { u64 tmp = t[i]; t[i] = t[i - 1]; t[i - 1] = tmp; }
        }
    }
}

template <int LPQ>
__device__ __forceinline__ void scan_range(const GridView& g, uint32_t s, uint32_t e, int gl,
                                           float qx, float qy, float qz, float skip_d2,
                                           u64 (&t)[5]) {
    for (uint32_t c = s + gl; c < e; c += LPQ) {
        float4 p = __ldg(g.pts + c);
        float d = sqdist(qx, qy, qz, p.x, p.y, p.z);
        if (d < skip_d2) top5_insert(t, make_key(d, p.w));
    }
}

// merge the lane-local lists of a group; every lane of the group returns the same 5 keys
template <int LPQ>
__device__ __forceinline__ void group_merge(unsigned gmask, u64 (&t)[5], u64 (&out)[5]) {
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        u64 m = t[0];
#pragma unroll
        for (int o = LPQ / 2; o > 0; o >>= 1) {
            u64 other = __shfl_xor_sync(gmask, m, o);
            m = other < m ? other : m;
        }
        out[r] = m;
        if (t[0] == m) {                       // the owner pops (keys are unique except kKeyNone)
            t[0] = t[1]; t[1] = t[2]; t[2] = t[3]; t[3] = t[4]; t[4] = kKeyNone;
        }
    }
}

// cell coordinate along one axis, clamped into the grid; u (unclamped, in cells) is returned too
__device__ __forceinline__ int cell_coord(float q, float o, float inv, int dim, float* u_out) {
    float u = (q - o) * inv;
    *u_out = u;
    float f = floorf(u);
    int c = f < 0.f ? 0 : (f > (float)(dim - 1) ? dim - 1 : (int)f);
    return c;
}

// scan [s, e) keeping only candidates with d2 <= tau (ties at tau are resolved by the key order)
template <int LPQ>
__device__ __forceinline__ void scan_range_le(const GridView& g, uint32_t s, uint32_t e, int gl,
                                              float qx, float qy, float qz, float tau, u64 (&t)[5]) {
    for (uint32_t c = s + gl; c < e; c += LPQ) {
        float4 p = __ldg(g.pts + c);
        float d = sqdist(qx, qy, qz, p.x, p.y, p.z);
        if (d <= tau) top5_insert(t, make_key(d, p.w));
    }
}

// ---- hashed directory ------------------------------------------------------------------------------
constexpr unsigned long long kHashEmpty = 0xffffffffffffffffull;
__host__ __device__ __forceinline__ uint32_t cell_hash(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 29;
    return (uint32_t)k;
}
// points of cell (x, y, z): [*s, *e) in g.pts; false when the cell is empty
__device__ __forceinline__ bool hash_cell_range(const GridView& g, int x, int y, int z, uint32_t* s, uint32_t* e) {
    const unsigned long long key = ((unsigned long long)z * (unsigned long long)g.dy + (unsigned long long)y) *
                                   (unsigned long long)g.dx + (unsigned long long)x;
    uint32_t slot = cell_hash(key) & g.hmask;
    for (;;) {
        const unsigned long long k = __ldg(g.hkeys + slot);
        if (k == key) {
            *s = __ldg(g.cell_start + slot);
            *e = __ldg(g.cell_start + slot + 1);
            return true;
        }
        if (k == kHashEmpty) return false;
        slot = (slot + 1) & g.hmask;
    }
}

// The search over a hashed directory (maps whose dense directory would dwarf them).  Same semantics as
// group_knn5 / group_knn5_gated: Chebyshev shells of cells around the query's cell, one hash probe per cell,
// until the 5th neighbour is provably inside the searched cube (exact) or after the 3x3x3 block (gated).
template <int LPQ>
__device__ __noinline__ void group_knn5_hashed(const GridView& g, float qx, float qy, float qz, int gl, unsigned gmask,
                                               bool exact, float gate_sq, u64* best_out) {
    u64 t[5] = {kKeyNone, kKeyNone, kKeyNone, kKeyNone, kKeyNone};
    u64 best[5] = {kKeyNone, kKeyNone, kKeyNone, kKeyNone, kKeyNone};
    float ux, uy, uz;
    const int cx = cell_coord(qx, g.ox, g.inv, g.dx, &ux);
    const int cy = cell_coord(qy, g.oy, g.inv, g.dy, &uy);
    const int cz = cell_coord(qz, g.oz, g.inv, g.dz, &uz);
    const float skip = exact ? __int_as_float(0x7f800000) : gate_sq;
    const bool far = !exact && (ux < -1.f || uy < -1.f || uz < -1.f || ux > (float)g.dx + 1.f ||
                                uy > (float)g.dy + 1.f || uz > (float)g.dz + 1.f);
    const float sl_x = 0.002f + 4e-7f * fabsf(ux), sl_y = 0.002f + 4e-7f * fabsf(uy), sl_z = 0.002f + 4e-7f * fabsf(uz);
    for (int rr = 1; !far; ++rr) {
        // shell rr: the cells at Chebyshev distance rr (rr == 1: the whole 3x3x3 block)
        for (int dzz = -rr; dzz <= rr; ++dzz) {
            const int zz = cz + dzz;
            if (zz < 0 || zz >= g.dz) continue;
            for (int dyy = -rr; dyy <= rr; ++dyy) {
                const int yy = cy + dyy;
                if (yy < 0 || yy >= g.dy) continue;
                const bool full = rr == 1 || dzz == -rr || dzz == rr || dyy == -rr || dyy == rr;
                const int step = full ? 1 : 2 * rr;
                for (int xx = cx - rr; xx <= cx + rr; xx += step) {
                    if (xx < 0 || xx >= g.dx) continue;
                    uint32_t s, e;
                    if (hash_cell_range(g, xx, yy, zz, &s, &e)) scan_range<LPQ>(g, s, e, gl, qx, qy, qz, skip, t);
                }
            }
        }
        group_merge<LPQ>(gmask, t, best);
        if (!exact) break;
        const int r = rr;          // cells within Chebyshev distance r are done: same stopping rule as group_knn5
        float bound = 3.0e38f;
        bool open = false;
        if (cx - r > 0)        { bound = fminf(bound, ux - (float)(cx - r) - sl_x); open = true; }
        if (cx + r < g.dx - 1) { bound = fminf(bound, (float)(cx + r + 1) - ux - sl_x); open = true; }
        if (cy - r > 0)        { bound = fminf(bound, uy - (float)(cy - r) - sl_y); open = true; }
        if (cy + r < g.dy - 1) { bound = fminf(bound, (float)(cy + r + 1) - uy - sl_y); open = true; }
        if (cz - r > 0)        { bound = fminf(bound, uz - (float)(cz - r) - sl_z); open = true; }
        if (cz + r < g.dz - 1) { bound = fminf(bound, (float)(cz + r + 1) - uz - sl_z); open = true; }
        if (!open) break;
        if (best[4] != kKeyNone && bound > 0.f) {
            const float bm = bound * g.cell;
            if (key_d2(best[4]) <= bm * bm * 0.9999f) break;
        }
        if (r >= 16) {             // far from any structure: finish with an exhaustive scan (still exact)
#pragma unroll
            for (int i = 0; i < 5; ++i) t[i] = kKeyNone;
            scan_range<LPQ>(g, 0, g.m, gl, qx, qy, qz, skip, t);
            group_merge<LPQ>(gmask, t, best);
            break;
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) t[i] = gl == 0 ? best[i] : kKeyNone;
    }
    if (!exact) {
#pragma unroll
        for (int i = 0; i < 5; ++i)
            if (best[i] != kKeyNone && !(key_d2(best[i]) < gate_sq)) best[i] = kKeyNone;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) best_out[i] = best[i];
}

// GATED search with geometric pruning.  The centre row (3 x-cells) is scanned first and merged;
// its 5th distance (or the gate) becomes tau.  A remaining row / x-cell is visited only when its
// lower-bound distance to the query (gap to the cell slab, minus a rounding slack) does not
// exceed tau, so a typical query touches 6-8 of the 27 cells.  Results are identical to the
// unpruned search: a skipped cell cannot hold a point with d2 <= tau.
template <int LPQ>
__device__ __forceinline__ void group_knn5_gated(const GridView& g, float qx, float qy, float qz, int gl,
                                                 unsigned gmask, float gate_sq, u64 (&best)[5]) {
    float ux, uy, uz;
    const int cx = cell_coord(qx, g.ox, g.inv, g.dx, &ux);
    const int cy = cell_coord(qy, g.oy, g.inv, g.dy, &uy);
    const int cz = cell_coord(qz, g.oz, g.inv, g.dz, &uz);
    const int lane_base = (threadIdx.x & 31) - gl;
    if (g.hashed) {                 // its address escapes into the out-of-line search; `best` stays in registers
        u64 hb[5];
        group_knn5_hashed<LPQ>(g, qx, qy, qz, gl, gmask, false, gate_sq, hb);
#pragma unroll
        for (int i = 0; i < 5; ++i) best[i] = hb[i];
        return;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) best[i] = kKeyNone;
    const bool far = ux < -1.f || uy < -1.f || uz < -1.f || ux > (float)g.dx + 1.f ||
                     uy > (float)g.dy + 1.f || uz > (float)g.dz + 1.f;
    if (far) return;                           // nothing within the gate

    // every lane fetches the 4 cell boundaries {cx-1, cx, cx+1, cx+2} of "its" rows
    constexpr int RPL = (9 + LPQ - 1) / LPQ;
    uint32_t b0[RPL], b1[RPL], b2[RPL], b3[RPL];
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
        const int r = gl + k * LPQ;
        b0[k] = b1[k] = b2[k] = b3[k] = 0;
        if (r < 9) {
            const int yy = cy + (r % 3) - 1, zz = cz + (r / 3) - 1;
            if (yy >= 0 && yy < g.dy && zz >= 0 && zz < g.dz) {
                const uint32_t* row = g.cell_start + ((uint32_t)zz * g.dy + yy) * g.dx;
                const uint32_t c1 = __ldg(row + cx), c2 = __ldg(row + cx + 1);
                b1[k] = c1; b2[k] = c2;
                b0[k] = cx > 0 ? __ldg(row + cx - 1) : c1;
                b3[k] = cx + 1 < g.dx ? __ldg(row + cx + 2) : c2;
            }
        }
    }
    // gaps (in cells, shrunk by the rounding slack) from the query to the neighbouring slabs
    const float slx = 0.002f + 4e-7f * fabsf(ux), sly = 0.002f + 4e-7f * fabsf(uy), slz = 0.002f + 4e-7f * fabsf(uz);
    const float gxm = fmaxf(ux - (float)cx - slx, 0.f) * g.cell, gxp = fmaxf((float)(cx + 1) - ux - slx, 0.f) * g.cell;
    const float gym = fmaxf(uy - (float)cy - sly, 0.f) * g.cell, gyp = fmaxf((float)(cy + 1) - uy - sly, 0.f) * g.cell;
    const float gzm = fmaxf(uz - (float)cz - slz, 0.f) * g.cell, gzp = fmaxf((float)(cz + 1) - uz - slz, 0.f) * g.cell;

    u64 t[5] = {kKeyNone, kKeyNone, kKeyNone, kKeyNone, kKeyNone};
    float tau = gate_sq;                        // candidates need d2 < gate: handled by the final filter
    // ---- centre row (row index 4), all three x cells ----
    {
        const uint32_t s = __shfl_sync(gmask, b0[4 / LPQ], lane_base + (4 % LPQ));
        const uint32_t e = __shfl_sync(gmask, b3[4 / LPQ], lane_base + (4 % LPQ));
        scan_range_le<LPQ>(g, s, e, gl, qx, qy, qz, tau, t);
        group_merge<LPQ>(gmask, t, best);
        if (best[4] != kKeyNone) tau = fminf(tau, key_d2(best[4]));
#pragma unroll
        for (int i = 0; i < 5; ++i) t[i] = gl == 0 ? best[i] : kKeyNone;
    }
    // ---- the other 8 rows, pruned by tau ----
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        if (r == 4) continue;
        const uint32_t c0 = __shfl_sync(gmask, b0[r / LPQ], lane_base + (r % LPQ));
        const uint32_t c1 = __shfl_sync(gmask, b1[r / LPQ], lane_base + (r % LPQ));
        const uint32_t c2 = __shfl_sync(gmask, b2[r / LPQ], lane_base + (r % LPQ));
        const uint32_t c3 = __shfl_sync(gmask, b3[r / LPQ], lane_base + (r % LPQ));
        const int dyy = (r % 3) - 1, dzz = (r / 3) - 1;
        const float gy = dyy == 0 ? 0.f : (dyy < 0 ? gym : gyp);
        const float gz = dzz == 0 ? 0.f : (dzz < 0 ? gzm : gzp);
        const float rb = gy * gy + gz * gz;
        if (rb > tau) continue;
        const uint32_t s = (rb + gxm * gxm > tau) ? c1 : c0;
        const uint32_t e = (rb + gxp * gxp > tau) ? c2 : c3;
        scan_range_le<LPQ>(g, s, e, gl, qx, qy, qz, tau, t);
    }
    group_merge<LPQ>(gmask, t, best);
    // gate: only distances strictly below the gate count (MO:1025 / MO:1121 test d2[4] < 1.0)
#pragma unroll
    for (int i = 0; i < 5; ++i)
        if (best[i] != kKeyNone && !(key_d2(best[i]) < gate_sq)) best[i] = kKeyNone;
}

// ---- thread-per-query GATED search -------------------------------------------------------------
// One lane owns one query and keeps the exact top-5 itself, so there is no merge and the pruning
// threshold tau = min(gate, current 5th distance) is always up to date.  Rows are visited nearest
// first (centre, the four face neighbours, the four diagonals); a row / x-cell whose slab gap
// exceeds tau is skipped.  Candidates are fetched four at a time to keep loads in flight.
__device__ __forceinline__ void thread_scan(const GridView& g, uint32_t s, uint32_t e, float qx, float qy,
                                            float qz, float gate_sq, u64 (&t)[5]) {
    for (uint32_t c = s; c < e; c += 4) {
        const uint32_t last = e - 1;
        const float4 p0 = __ldg(g.pts + c);
        const float4 p1 = __ldg(g.pts + min(c + 1, last));
        const float4 p2 = __ldg(g.pts + min(c + 2, last));
        const float4 p3 = __ldg(g.pts + min(c + 3, last));
        const float inf = __int_as_float(0x7f800000);
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

__device__ __forceinline__ void thread_knn5_gated(const GridView& g, float qx, float qy, float qz,
                                                  float gate_sq, u64 (&t)[5]) {
    if (g.hashed) {
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
    // (dy, dz) visiting order: centre, faces, diagonals
    const int ody[9] = {0, -1, 1, 0, 0, -1, 1, -1, 1};
    const int odz[9] = {0, 0, 0, -1, 1, -1, -1, 1, 1};
    // all 9 rows' cell boundaries are fetched up front (independent loads, one latency)
    uint32_t b0[9], b1[9], b2[9], b3[9];
    const int xl = cx > 0 ? cx - 1 : cx, xh = cx + 1 < g.dx ? cx + 2 : cx + 1;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const int yy = cy + ody[r], zz = cz + odz[r];
        b0[r] = b1[r] = b2[r] = b3[r] = 0;
        if (yy >= 0 && yy < g.dy && zz >= 0 && zz < g.dz) {
            const uint32_t* row = g.cell_start + ((uint32_t)zz * g.dy + yy) * g.dx;
            b0[r] = __ldg(row + xl);
            b1[r] = __ldg(row + cx);
            b2[r] = __ldg(row + cx + 1);
            b3[r] = __ldg(row + xh);
        }
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const int dyy = ody[r], dzz = odz[r];
        const float gy = dyy == 0 ? 0.f : (dyy < 0 ? gym : gyp);
        const float gz = dzz == 0 ? 0.f : (dzz < 0 ? gzm : gzp);
        const float rb = gy * gy + gz * gz;
        // tau: only points with d2 <= tau can still enter the list (strictly below the gate)
        const float tau = t[4] == kKeyNone ? gate_sq : key_d2(t[4]);
        if (rb > tau) continue;
        const uint32_t s = (rb + gxm2 > tau) ? b1[r] : b0[r];
        const uint32_t e = (rb + gxp2 > tau) ? b2[r] : b3[r];
        thread_scan(g, s, e, qx, qy, qz, gate_sq, t);
    }
}

// One query, LPQ cooperating lanes (gl = lane index inside the group, gmask = the group's lanes).
// exact == false: 3x3x3 block only, candidates with d2 >= gate_sq are dropped.
template <int LPQ>
__device__ __forceinline__ void group_knn5(const GridView& g, float qx, float qy, float qz, int gl,
                                           unsigned gmask, bool exact, float gate_sq,
                                           u64 (&best)[5]) {
    u64 t[5] = {kKeyNone, kKeyNone, kKeyNone, kKeyNone, kKeyNone};
    float ux, uy, uz;
    const int cx = cell_coord(qx, g.ox, g.inv, g.dx, &ux);
    const int cy = cell_coord(qy, g.oy, g.inv, g.dy, &uy);
    const int cz = cell_coord(qz, g.oz, g.inv, g.dz, &uz);
    const float skip = exact ? __int_as_float(0x7f800000) : gate_sq;
    const int lane_base = (threadIdx.x & 31) - gl;
    if (g.hashed) {
        u64 hb[5];
        group_knn5_hashed<LPQ>(g, qx, qy, qz, gl, gmask, exact, gate_sq, hb);
#pragma unroll
        for (int i = 0; i < 5; ++i) best[i] = hb[i];
        return;
    }

    if (!exact) {
        // a query more than one cell outside the grid has nothing within the gate
        bool far = ux < -1.f || uy < -1.f || uz < -1.f || ux > (float)g.dx + 1.f ||
                   uy > (float)g.dy + 1.f || uz > (float)g.dz + 1.f;
        if (far) {
#pragma unroll
            for (int i = 0; i < 5; ++i) best[i] = kKeyNone;
            return;
        }
    }

    // ---- 3x3x3 block: 9 rows, ranges fetched in parallel by the group's lanes ----
    constexpr int RPL = (9 + LPQ - 1) / LPQ;
    uint32_t rs[RPL], re[RPL];
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dx - 1);
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
        const int r = gl + k * LPQ;
        rs[k] = 0; re[k] = 0;
        if (r < 9) {
            const int yy = cy + (r % 3) - 1, zz = cz + (r / 3) - 1;
            if (yy >= 0 && yy < g.dy && zz >= 0 && zz < g.dz) {
                const uint32_t row = ((uint32_t)zz * g.dy + yy) * g.dx;
                rs[k] = __ldg(g.cell_start + row + x0);
                re[k] = __ldg(g.cell_start + row + x1 + 1);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const uint32_t s = __shfl_sync(gmask, rs[r / LPQ], lane_base + (r % LPQ));
        const uint32_t e = __shfl_sync(gmask, re[r / LPQ], lane_base + (r % LPQ));
        scan_range<LPQ>(g, s, e, gl, qx, qy, qz, skip, t);
    }
    group_merge<LPQ>(gmask, t, best);
    if (!exact) return;

    // ---- exact mode: grow Chebyshev shells until the 5th neighbour is provably inside ----
    const float sl_x = 0.002f + 4e-7f * fabsf(ux), sl_y = 0.002f + 4e-7f * fabsf(uy),
                sl_z = 0.002f + 4e-7f * fabsf(uz);
    for (int r = 1;; ++r) {
        // distance (in cells) from the query to the nearest face that still hides unvisited cells
        float bound = 3.0e38f;
        bool open = false;
        if (cx - r > 0)        { bound = fminf(bound, ux - (float)(cx - r) - sl_x); open = true; }
        if (cx + r < g.dx - 1) { bound = fminf(bound, (float)(cx + r + 1) - ux - sl_x); open = true; }
        if (cy - r > 0)        { bound = fminf(bound, uy - (float)(cy - r) - sl_y); open = true; }
        if (cy + r < g.dy - 1) { bound = fminf(bound, (float)(cy + r + 1) - uy - sl_y); open = true; }
        if (cz - r > 0)        { bound = fminf(bound, uz - (float)(cz - r) - sl_z); open = true; }
        if (cz + r < g.dz - 1) { bound = fminf(bound, (float)(cz + r + 1) - uz - sl_z); open = true; }
        if (!open) break;                                   // the whole grid has been visited
        if (best[4] != kKeyNone && bound > 0.f) {
            float bm = bound * g.cell;
            if (key_d2(best[4]) <= bm * bm * 0.9999f) break;
        }
        if (r >= 16) {
            // far from any structure: finish with an exhaustive scan (still exact)
#pragma unroll
            for (int i = 0; i < 5; ++i) t[i] = kKeyNone;
            scan_range<LPQ>(g, 0, g.m, gl, qx, qy, qz, skip, t);
            group_merge<LPQ>(gmask, t, best);
            break;
        }
        // restart the lane lists from the merged result (lane 0 keeps it) and add shell r+1
        const int rr = r + 1;
#pragma unroll
        for (int i = 0; i < 5; ++i) t[i] = gl == 0 ? best[i] : kKeyNone;
        for (int dzz = -rr; dzz <= rr; ++dzz) {
            const int zz = cz + dzz;
            if (zz < 0 || zz >= g.dz) continue;
            for (int dyy = -rr; dyy <= rr; ++dyy) {
                const int yy = cy + dyy;
                if (yy < 0 || yy >= g.dy) continue;
                const uint32_t row = ((uint32_t)zz * g.dy + yy) * g.dx;
                const bool full = (dzz == -rr || dzz == rr || dyy == -rr || dyy == rr);
                if (full) {
                    const int a = max(cx - rr, 0), b = min(cx + rr, g.dx - 1);
                    if (a <= b)
                        scan_range<LPQ>(g, __ldg(g.cell_start + row + a), __ldg(g.cell_start + row + b + 1),
                                        gl, qx, qy, qz, skip, t);
                } else {
                    const int a = cx - rr, b = cx + rr;
                    if (a >= 0)
                        scan_range<LPQ>(g, __ldg(g.cell_start + row + a), __ldg(g.cell_start + row + a + 1),
                                        gl, qx, qy, qz, skip, t);
                    if (b < g.dx)
                        scan_range<LPQ>(g, __ldg(g.cell_start + row + b), __ldg(g.cell_start + row + b + 1),
                                        gl, qx, qy, qz, skip, t);
                }
            }
        }
        group_merge<LPQ>(gmask, t, best);
    }
}

// ---- stage-level kernel: materialised 5-NN (lvreg_knn5) ---------------------------------------
template <int LPQ>
__global__ void __launch_bounds__(256) knn5_grid_kernel(GridView g, const float4* __restrict__ queries,
                                                        uint32_t nq, int exact, float gate_sq,
                                                        int32_t* __restrict__ idx_out,
                                                        float* __restrict__ d2_out) {
    constexpr int QPB = 256 / LPQ;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LPQ;
    const unsigned gmask = (LPQ == 32) ? 0xffffffffu : (((1u << LPQ) - 1u) << (lane - gl));
    for (uint32_t q = blockIdx.x * QPB + threadIdx.x / LPQ; q < nq; q += gridDim.x * QPB) {
        float4 p = __ldg(queries + q);
        u64 best[5];
        if (exact) group_knn5<LPQ>(g, p.x, p.y, p.z, gl, gmask, true, gate_sq, best);
        else group_knn5_gated<LPQ>(g, p.x, p.y, p.z, gl, gmask, gate_sq, best);
        if (gl == 0) {
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                idx_out[(size_t)q * 5 + i] = key_idx(best[i]);
                d2_out[(size_t)q * 5 + i] = key_d2(best[i]);
            }
        }
    }
}

// ---- brute force (FP32-pipe bound variant of the micro-benchmark) -----------------------------
// thread per query, map staged through shared memory in tiles; blockIdx.y splits the map.
constexpr int kBruteTile = 1024;
constexpr int kBruteQpt = 2;                 // queries per thread: one shared-memory read serves two distance evaluations
constexpr int kBruteQpb = 256 * kBruteQpt;   // queries per block
// Per pair: 3 subtractions, 3 multiplications, 2 additions (no FMA: FLANN's L2_Simple rounds every operation), one
// fp32 compare against the query's current 5th distance; the 64-bit (d2, index) key is only formed for the few
// candidates that pass it.
__global__ void __launch_bounds__(256) knn5_brute_kernel(const float4* __restrict__ map, uint32_t m,
                                                         const float4* __restrict__ queries,
                                                         uint32_t nq, uint32_t chunk,
                                                         u64* __restrict__ partial) {
    __shared__ float4 tile[kBruteTile];
    const uint32_t q0 = blockIdx.x * kBruteQpb + threadIdx.x, q1 = q0 + 256;
    const float4 a = q0 < nq ? queries[q0] : make_float4(0, 0, 0, 0);
    const float4 b = q1 < nq ? queries[q1] : make_float4(0, 0, 0, 0);
    u64 ta[5] = {kKeyNone, kKeyNone, kKeyNone, kKeyNone, kKeyNone};
    u64 tb[5] = {kKeyNone, kKeyNone, kKeyNone, kKeyNone, kKeyNone};
    float wa = __int_as_float(0x7f800000), wb = wa;          // current 5th distance (inf while the list is short)
    const uint32_t begin = blockIdx.y * chunk;
    const uint32_t end = min(m, begin + chunk);
    for (uint32_t base = begin; base < end; base += kBruteTile) {
        const uint32_t cnt = min((uint32_t)kBruteTile, end - base);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < kBruteTile; i += 256)
            tile[i] = i < cnt ? ld_stream(map + base + i) : make_float4(3.0e19f, 3.0e19f, 3.0e19f, 0.f);
        __syncthreads();
        // eight candidates per step: 16 distances and their two minima, ONE branch; the (rare) step in which some
        // candidate beats a current 5th distance is redone candidate by candidate by a single rolled copy of the
        // insertion code (the distances are recomputed from shared memory: same inputs, same bits)
        for (uint32_t i0 = 0; i0 < cnt; i0 += 8) {
            float ma = __int_as_float(0x7f800000), mb = ma;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float4 p = tile[i0 + k];                 // rows past cnt hold points at infinity
                ma = fminf(ma, sqdist(a.x, a.y, a.z, p.x, p.y, p.z));
                mb = fminf(mb, sqdist(b.x, b.y, b.z, p.x, p.y, p.z));
            }
            if (ma <= wa || mb <= wb) {
#pragma unroll 1
                for (uint32_t i = i0; i < min(i0 + 8, cnt); ++i) {
                    const float4 p = tile[i];
                    const float da = sqdist(a.x, a.y, a.z, p.x, p.y, p.z);
                    const float db = sqdist(b.x, b.y, b.z, p.x, p.y, p.z);
                    if (da <= wa) {
                        top5_insert(ta, ((u64)__float_as_uint(da) << 32) | (u64)(base + i));
                        if (ta[4] != kKeyNone) wa = key_d2(ta[4]);
                    }
                    if (db <= wb) {
                        top5_insert(tb, ((u64)__float_as_uint(db) << 32) | (u64)(base + i));
                        if (tb[4] != kKeyNone) wb = key_d2(tb[4]);
                    }
                }
            }
        }
    }
    if (q0 < nq) {
#pragma unroll
        for (int i = 0; i < 5; ++i) partial[((size_t)blockIdx.y * nq + q0) * 5 + i] = ta[i];
    }
    if (q1 < nq) {
#pragma unroll
        for (int i = 0; i < 5; ++i) partial[((size_t)blockIdx.y * nq + q1) * 5 + i] = tb[i];
    }
}

__global__ void __launch_bounds__(256) knn5_brute_merge_kernel(const u64* __restrict__ partial,
                                                               uint32_t nq, uint32_t splits,
                                                               int32_t* __restrict__ idx_out,
                                                               float* __restrict__ d2_out) {
    const uint32_t q = blockIdx.x * 256 + threadIdx.x;
    if (q >= nq) return;
    u64 t[5] = {kKeyNone, kKeyNone, kKeyNone, kKeyNone, kKeyNone};
    for (uint32_t s = 0; s < splits; ++s)
#pragma unroll
        for (int i = 0; i < 5; ++i) top5_insert(t, partial[((size_t)s * nq + q) * 5 + i]);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        idx_out[(size_t)q * 5 + i] = key_idx(t[i]);
        d2_out[(size_t)q * 5 + i] = key_d2(t[i]);
    }
}

// ---- spatial ordering of the queries -----------------------------------------------------------------
// The registration kernel stages, per 32-query tile, the box of map cells its queries touch; the tile must be
// spatially compact for the box to fit in shared memory.  VoxelGrid order (x-fastest voxel rows of the whole
// scan) is not compact, so the down-sampled scan is additionally ordered by the Morton code of its 2 m cell
// (8 bits per axis, wrapping every 512 m: aliases only cost efficiency, never correctness).  The order is
// taken in the sensor frame; a rigid transform keeps a compact tile compact.
constexpr float kQueryCellInv = 0.5f;
__device__ __forceinline__ uint32_t spread3_8(uint32_t v) {          // 8 bits -> every third bit
    v &= 0xffu;
    v = (v | (v << 8)) & 0x00f00fu;
    v = (v | (v << 4)) & 0x0c30c3u;
    v = (v | (v << 2)) & 0x249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton_cell_key(float x, float y, float z) {
    const int cx = (int)floorf(x * kQueryCellInv), cy = (int)floorf(y * kQueryCellInv), cz = (int)floorf(z * kQueryCellInv);
    return spread3_8((uint32_t)cx) | (spread3_8((uint32_t)cy) << 1) | (spread3_8((uint32_t)cz) << 2);
}
__global__ void __launch_bounds__(256) morton_keys_kernel(const float4* __restrict__ pts, uint32_t n,
                                                          uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    keys[i] = morton_cell_key(p.x, p.y, p.z);
    vals[i] = i;
}
__global__ void __launch_bounds__(256) gather_points_kernel(const float4* __restrict__ pts,
                                                            const uint32_t* __restrict__ order, uint32_t n,
                                                            float4* __restrict__ out) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    out[j] = __ldg(pts + order[j]);
}

// ---- search-grid build ---------------------------------------------------------------------------
struct GridSpec {
    float ox, oy, oz, inv, cell;
    int dx, dy, dz;
};

// Counting-sort build of the search grid: the atomic that counts a cell's points also hands every point its rank
// inside the cell, so the points can be scattered to cell_start[cell] + rank right after the scan of the counts:
// two passes over the map and one scan instead of a radix sort by cell key.  The order of the points INSIDE a
// cell is the arrival order of the atomics; the search keeps the 5 smallest (d2, index) keys, which do not depend
// on the order in which a cell's points are visited, so results are unaffected.
__global__ void __launch_bounds__(256) cell_count_kernel(const float4* __restrict__ pts, uint32_t m, GridSpec gs,
                                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ ranks,
                                                         uint32_t* __restrict__ counts) {
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    float4 p = pts[i];
    float u;
    int cx = cell_coord(p.x, gs.ox, gs.inv, gs.dx, &u);
    int cy = cell_coord(p.y, gs.oy, gs.inv, gs.dy, &u);
    int cz = cell_coord(p.z, gs.oz, gs.inv, gs.dz, &u);
    uint32_t key = ((uint32_t)cz * gs.dy + cy) * gs.dx + cx;
    keys[i] = key;
    ranks[i] = atomicAdd(&counts[key], 1u);
}

// the same count + rank step over a hashed directory: the cell's slot is found (or claimed) by linear probing
__global__ void __launch_bounds__(256) hash_count_kernel(const float4* __restrict__ pts, uint32_t m, GridSpec gs,
                                                         unsigned long long* __restrict__ hkeys, uint32_t hmask,
                                                         uint32_t* __restrict__ slots, uint32_t* __restrict__ ranks,
                                                         uint32_t* __restrict__ counts) {
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    float4 p = pts[i];
    float u;
    const int cx = cell_coord(p.x, gs.ox, gs.inv, gs.dx, &u);
    const int cy = cell_coord(p.y, gs.oy, gs.inv, gs.dy, &u);
    const int cz = cell_coord(p.z, gs.oz, gs.inv, gs.dz, &u);
    const unsigned long long key = ((unsigned long long)cz * (unsigned long long)gs.dy + (unsigned long long)cy) *
                                   (unsigned long long)gs.dx + (unsigned long long)cx;
    uint32_t slot = cell_hash(key) & hmask;
    for (;;) {
        const unsigned long long prev = atomicCAS(hkeys + slot, kHashEmpty, key);
        if (prev == kHashEmpty || prev == key) break;
        slot = (slot + 1) & hmask;
    }
    slots[i] = slot;
    ranks[i] = atomicAdd(&counts[slot], 1u);
}

__global__ void __launch_bounds__(256) cell_scatter_kernel(const float4* __restrict__ pts,
                                                           const uint32_t* __restrict__ keys,
                                                           const uint32_t* __restrict__ ranks,
                                                           const uint32_t* __restrict__ cell_start, uint32_t m,
                                                           float4* __restrict__ out) {
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    float4 p = pts[i];
    out[__ldg(cell_start + keys[i]) + ranks[i]] = make_float4(p.x, p.y, p.z, __uint_as_float(i));
}

// The whole dense-directory build for small maps as ONE cooperative launch: count + rank, exclusive scan of the
// cell counts (block chunk sums -> grid barrier -> every block adds the chunks before its own), scatter.  Six
// launches of a few microseconds each become one (C1 / C2 / C5: the grid build is pure launch latency).  For large
// maps the separate kernels stay faster (measured at C3: 0.078 ms against 0.129 ms with this kernel: two such grids
// cannot be co-resident, and a directory of 10^7 cells wants more threads than a co-resident grid has).
constexpr int kGridMidThreads = 256;
constexpr uint32_t kGridMidMaxPoints = 1u << 18, kGridMidMaxCells = 1u << 20;
__global__ void __launch_bounds__(kGridMidThreads) grid_build_mid_kernel(const float4* __restrict__ pts, uint32_t m, GridSpec gs,
                                                                         uint32_t ncells, uint32_t* __restrict__ keys,
                                                                         uint32_t* __restrict__ ranks,
                                                                         uint32_t* __restrict__ counts /* zeroed, ncells + 1 */,
                                                                         uint32_t* __restrict__ chunk_sums /* gridDim.x */,
                                                                         uint32_t* __restrict__ cell_start /* ncells + 1 */,
                                                                         float4* __restrict__ out) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ uint32_t red[kGridMidThreads / 32];
    __shared__ uint32_t before_s;
    const uint32_t nthreads = gridDim.x * kGridMidThreads, gtid = blockIdx.x * kGridMidThreads + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = gtid; i < m; i += nthreads) {
        const float4 p = pts[i];
        float u;
        const int cx = cell_coord(p.x, gs.ox, gs.inv, gs.dx, &u);
        const int cy = cell_coord(p.y, gs.oy, gs.inv, gs.dy, &u);
        const int cz = cell_coord(p.z, gs.oz, gs.inv, gs.dz, &u);
        const uint32_t key = ((uint32_t)cz * gs.dy + cy) * gs.dx + cx;
        keys[i] = key;
        ranks[i] = atomicAdd(&counts[key], 1u);
    }
    __threadfence();
    grid.sync();
    // block b scans the contiguous chunk [b * per, (b + 1) * per) of the ncells + 1 counts
    const uint32_t total = ncells + 1;
    const uint32_t per = (total + gridDim.x - 1) / gridDim.x;
    const uint32_t c0 = min(blockIdx.x * per, total), c1 = min(c0 + per, total);
    uint32_t s = 0;
    for (uint32_t i = c0 + threadIdx.x; i < c1; i += kGridMidThreads) s += counts[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kGridMidThreads / 32; ++w) t += red[w];
        chunk_sums[blockIdx.x] = t;
    }
    __threadfence();
    grid.sync();
    {   // sum of the chunks before this block's, by the whole block
        uint32_t t = 0;
        for (uint32_t b = threadIdx.x; b < blockIdx.x; b += kGridMidThreads) t += chunk_sums[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        __syncthreads();
        if (lane == 0) red[warp] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t a = 0;
            for (int w = 0; w < kGridMidThreads / 32; ++w) a += red[w];
            before_s = a;
        }
        __syncthreads();
    }
    // exclusive scan of the chunk, 256 cells per step, carried in index order
    uint32_t carry = before_s;
    for (uint32_t i0 = c0; i0 < c1; i0 += kGridMidThreads) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t v = i < c1 ? counts[i] : 0u;
        uint32_t inc = warp_inclusive_scan(v, lane);
        __syncthreads();
        if (lane == 31) red[warp] = inc;
        __syncthreads();
        uint32_t wpre = 0, all = 0;
        for (int w = 0; w < kGridMidThreads / 32; ++w) {
            if (w < warp) wpre += red[w];
            all += red[w];
        }
        if (i < c1) cell_start[i] = carry + wpre + inc - v;
        carry += all;
    }
    __threadfence();
    grid.sync();
    for (uint32_t i = gtid; i < m; i += nthreads) {
        const float4 p = pts[i];
        out[cell_start[keys[i]] + ranks[i]] = make_float4(p.x, p.y, p.z, __uint_as_float(i));
    }
}

struct CountIn {
    const uint32_t* c;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return c[i]; }
    __device__ __forceinline__ void load_vec(uint32_t i, uint32_t (&v)[8]) const {
        const uint4 a = *reinterpret_cast<const uint4*>(c + i);
        const uint4 b = *reinterpret_cast<const uint4*>(c + i + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};
struct StartOut {
    uint32_t* s;
    __device__ __forceinline__ void operator()(uint32_t i, uint32_t, uint32_t pre) const { s[i] = pre; }
    __device__ __forceinline__ void store_vec(uint32_t i, const uint32_t (&v)[8], uint32_t pre) const {
        uint4 a, b;
        a.x = pre; a.y = a.x + v[0]; a.z = a.y + v[1]; a.w = a.z + v[2];
        b.x = a.w + v[3]; b.y = b.x + v[4]; b.z = b.y + v[5]; b.w = b.z + v[6];
        *reinterpret_cast<uint4*>(s + i) = a;
        *reinterpret_cast<uint4*>(s + i + 4) = b;
    }
};

}  // namespace lvreg
