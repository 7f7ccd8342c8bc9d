// icp.cuh -- loop-closure registration on the device: pcl::IterativeClosestPoint as
// configured at MO:578-590 (point-to-point, 1-NN correspondences within max_corr_dist, Umeyama / SVD
// transformation estimate, PCL's default convergence criteria) and Registration::getFitnessScore
// (MO:592), on the submaps of loopFindNearKeyframes (MO:719-741; built with the transform + VoxelGrid
// kernels of voxelgrid.cuh).
//
//   icp_correspond_kernel  1-NN of every (current) source point in the target's cell grid, exact with
//                          the (d2, index) tie-break; kept pairs feed 17 double moments
//                          {n, sum d2, sum s, sum t, sum t s^T}, reduced per block in a fixed order
//   icp_update_kernel      one block: fixed-order reduction of the block partials, Umeyama from the
//                          moments (one-sided Jacobi SVD, double), final <- T * final, iteration count,
//                          DefaultConvergenceCriteria -- the whole per-iteration decision stays on the
//                          device, the host only polls the state every few iterations
//   icp_transform_kernel   source <- T * source
//   icp_fitness_kernel     mean squared 1-NN distance of final * (original source), no range cap
//
// Per iteration and source point: 16 B read + 16 B written + the candidate cells of the grid search
// (L2-resident target, <= a few MB); latency-bound like the LM loop, not bandwidth-bound.
#pragma once

#include "common.cuh"
#include "knn.cuh"

namespace lvreg {

enum { ICP_NOT_CONVERGED = 0, ICP_ITERATIONS = 1, ICP_TRANSFORM = 2, ICP_ABS_MSE = 3, ICP_REL_MSE = 4,
       ICP_NO_CORRESPONDENCES = 5, ICP_NO_INPUT = 6 };

struct IcpState {
    int done, state, iterations, n_corr;
    double mse_prev, mse;
    double fitness_sum;
    float T_inc[16];       // this iteration's transformation_ (row-major 4x4)
    float T_final[16];     // final_transformation_
};

struct IcpParams {
    double max_d2;          // corr_dist_threshold_^2 (double, as PCL compares)
    double rot_thr, trans_thr, rel_mse, abs_mse;
    int max_iterations;
};

constexpr int kIcpMoments = 17;
constexpr int kIcpThreads = 256;

// Running best of a nearest-neighbour search: (d2 bits << 32 | index) key plus the matched point
struct Nn1Best {
    u64 key;
    float d2, x, y, z;
    __device__ __forceinline__ Nn1Best() : key(kKeyNone), d2(INFINITY), x(0.f), y(0.f), z(0.f) {}
};

// Chebyshev shells s = 0 .. s_limit around q's (clamped) cell of grid g.  Every point of shell s is
// at least (s-1) cells away, so the search is complete once the best distance is below that bound
// (returns true), or once the bound passes max_d2 (such a neighbour would be rejected anyway; also
// true).  Returns false when s_limit was reached first.  Rows and cells whose slab cannot beat the
// best distance are skipped without touching memory.
__device__ __forceinline__ bool nn1_shells(const GridView& g, float qx, float qy, float qz, float max_d2,
                                           int s_limit, Nn1Best& best) {
    float ux, uy, uz;
    const int cx = cell_coord(qx, g.ox, g.inv, g.dx, &ux);
    const int cy = cell_coord(qy, g.oy, g.inv, g.dy, &uy);
    const int cz = cell_coord(qz, g.oz, g.inv, g.dz, &uz);
    const float slack = 2e-3f * g.cell;
    const int smax = max(max(max(cx, g.dx - 1 - cx), max(cy, g.dy - 1 - cy)), max(cz, g.dz - 1 - cz));
    for (int s = 0;; ++s) {
        if (s > 1) {
            const float lb = (float)(s - 1) * g.cell * 0.9999f;
            const float lb2 = lb * lb;
            if (best.d2 < lb2 || lb2 > max_d2) return true;
        }
        if (s > smax) return true;                      // the whole grid has been visited
        if (s > s_limit) return false;
        const int z0 = max(cz - s, 0), z1 = min(cz + s, g.dz - 1);
        const int y0 = max(cy - s, 0), y1 = min(cy + s, g.dy - 1);
        for (int z = z0; z <= z1; ++z) {
            float dzl = fmaxf(fmaxf(((float)z - uz) * g.cell, (uz - (float)(z + 1)) * g.cell) - slack, 0.f);
            const float dz2 = dzl * dzl;
            if (dz2 > best.d2) continue;
            const bool zface = (z == cz - s) || (z == cz + s);
            for (int y = y0; y <= y1; ++y) {
                float dyl = fmaxf(fmaxf(((float)y - uy) * g.cell, (uy - (float)(y + 1)) * g.cell) - slack, 0.f);
                const float dyz2 = dz2 + dyl * dyl;
                if (dyz2 > best.d2) continue;
                const bool face = zface || (y == cy - s) || (y == cy + s);
                const uint32_t row = ((uint32_t)z * g.dy + y) * g.dx;
                // a face row contributes its whole x-range, an interior row only its two end cells
                const int nseg = face ? 1 : 2;
                for (int k = 0; k < nseg; ++k) {
                    int xa, xb;
                    if (face) { xa = max(cx - s, 0); xb = min(cx + s, g.dx - 1); }
                    else { xa = xb = k == 0 ? cx - s : cx + s; if (xa < 0 || xa >= g.dx) continue; }
                    if (!face) {
                        float dxl = fmaxf(fmaxf(((float)xa - ux) * g.cell, (ux - (float)(xa + 1)) * g.cell) - slack, 0.f);
                        if (dyz2 + dxl * dxl > best.d2) continue;
                    }
                    const uint32_t b = __ldg(g.cell_start + row + xa), e = __ldg(g.cell_start + row + xb + 1);
                    for (uint32_t c = b; c < e; ++c) {
                        const float4 p = __ldg(g.pts + c);
                        const float d = sqdist(qx, qy, qz, p.x, p.y, p.z);
                        const u64 key = make_key(d, p.w);
                        if (key < best.key) { best.key = key; best.d2 = d; best.x = p.x; best.y = p.y; best.z = p.z; }
                    }
                }
            }
        }
    }
}

// Exact nearest neighbour with the (d2, index) tie-break over two grids of the same target: the fine
// grid (1 m cells) resolves every query whose neighbour lies within one cell -- nearly all of them in
// an overlapping pair of submaps -- from 27 cells; the rest (non-overlapping parts, up to
// max_corr_dist = 30 m away) continue on the coarse grid, where 30 m are 4-5 shells instead of 30.
__device__ __forceinline__ Nn1Best nn1_search(const GridView& fine, const GridView& coarse, float qx, float qy,
                                              float qz, float max_d2) {
    Nn1Best best;
    if (nn1_shells(fine, qx, qy, qz, max_d2, 1, best)) return best;      // complete after shells 0 and 1
    nn1_shells(coarse, qx, qy, qz, max_d2, 0x7fffffff, best);              // the fine result keeps pruning
    return best;
}

// stage-level 1-NN (parity tests): idx -1 / d2 +inf when the target is empty
__global__ void __launch_bounds__(kIcpThreads) nn1_kernel(const float4* __restrict__ q, uint32_t n, GridView g,
                                                          GridView gc, float max_d2, int32_t* __restrict__ idx,
                                                          float* __restrict__ d2) {
    const uint32_t i = blockIdx.x * kIcpThreads + threadIdx.x;
    if (i >= n) return;
    const float4 p = q[i];
    const Nn1Best b = nn1_search(g, gc, p.x, p.y, p.z, max_d2);
    idx[i] = key_idx(b.key);
    d2[i] = key_d2(b.key);
}

// block-wide sum of NV doubles per thread, fixed order; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* [8][NV] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) smem[warp * NV + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = smem[k];
            for (int w = 1; w < kIcpThreads / 32; ++w) s += smem[w * NV + k];
            v[k] = s;
        }
    }
}

__global__ void __launch_bounds__(kIcpThreads) icp_correspond_kernel(const float4* __restrict__ cur, uint32_t n,
                                                                     GridView g, GridView gc, IcpParams P,
                                                                     const IcpState* __restrict__ st,
                                                                     double* __restrict__ partials) {
    if (st->done) return;
    __shared__ double red[8 * kIcpMoments];
    double m[kIcpMoments];
#pragma unroll
    for (int k = 0; k < kIcpMoments; ++k) m[k] = 0.0;
    const uint32_t i = blockIdx.x * kIcpThreads + threadIdx.x;
    if (i < n) {
        const float4 s = cur[i];
        const float gate = P.max_d2 < 3.0e38 ? (float)P.max_d2 * 1.0001f : __int_as_float(0x7f800000);
        const Nn1Best t = nn1_search(g, gc, s.x, s.y, s.z, gate);
        if (t.key != kKeyNone && !((double)t.d2 > P.max_d2)) {           // PCL: skip when d2 > max_dist^2
            m[0] = 1.0;
            m[1] = (double)t.d2;
            m[2] = s.x; m[3] = s.y; m[4] = s.z;
            m[5] = t.x; m[6] = t.y; m[7] = t.z;
            m[8] = (double)t.x * s.x;  m[9] = (double)t.x * s.y;  m[10] = (double)t.x * s.z;
            m[11] = (double)t.y * s.x; m[12] = (double)t.y * s.y; m[13] = (double)t.y * s.z;
            m[14] = (double)t.z * s.x; m[15] = (double)t.z * s.y; m[16] = (double)t.z * s.z;
        }
    }
    block_sum<kIcpMoments>(m, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kIcpMoments; ++k) partials[(size_t)blockIdx.x * kIcpMoments + k] = m[k];
    }
}

// Umeyama without scaling from the raw moments (double); T row-major 4x4 (float)
__device__ inline void umeyama_from_moments(const double* mom, float* T) {
    const double n = mom[0];
    double ms[3], mt[3], A[3][3], V[3][3];
    for (int i = 0; i < 3; ++i) { ms[i] = mom[2 + i] / n; mt[i] = mom[5 + i] / n; }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            A[i][j] = mom[8 + 3 * i + j] / n - mt[i] * ms[j];
            V[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) {
                    alpha += A[i][p] * A[i][p];
                    beta += A[i][q] * A[i][q];
                    gamma += A[i][p] * A[i][q];
                }
                if (gamma == 0.0 || fabs(gamma) <= 1e-15 * sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < 3; ++i) {
                    const double ap = A[i][p], aq = A[i][q];
                    A[i][p] = c * ap - s * aq;
                    A[i][q] = s * ap + c * aq;
                    const double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - s * vq;
                    V[i][q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double sv[3];
    for (int j = 0; j < 3; ++j) sv[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
    int ord[3] = {0, 1, 2};
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2 - a; ++b)
            if (sv[ord[b]] < sv[ord[b + 1]]) { int tmp = ord[b]; ord[b] = ord[b + 1]; ord[b + 1] = tmp; }
    double U[3][3], W[3][3];
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) W[i][j] = V[i][ord[j]];
    for (int j = 0; j < 2; ++j) {
        const double s = sv[ord[j]];
        for (int i = 0; i < 3; ++i) U[i][j] = s > 0.0 ? A[i][ord[j]] / s : (i == j ? 1.0 : 0.0);
    }
    {
        const double s = sv[ord[2]];
        if (s > 1e-12 * sv[ord[0]] && s > 0.0) {
            for (int i = 0; i < 3; ++i) U[i][2] = A[i][ord[2]] / s;
        } else {
            U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
            U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
            U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
        }
    }
    const double detU = U[0][0] * (U[1][1] * U[2][2] - U[1][2] * U[2][1]) - U[0][1] * (U[1][0] * U[2][2] - U[1][2] * U[2][0]) +
                        U[0][2] * (U[1][0] * U[2][1] - U[1][1] * U[2][0]);
    const double detW = W[0][0] * (W[1][1] * W[2][2] - W[1][2] * W[2][1]) - W[0][1] * (W[1][0] * W[2][2] - W[1][2] * W[2][0]) +
                        W[0][2] * (W[1][0] * W[2][1] - W[1][1] * W[2][0]);
    const double d = detU * detW < 0.0 ? -1.0 : 1.0;
    for (int i = 0; i < 3; ++i) {
        double R[3];
        for (int j = 0; j < 3; ++j) R[j] = U[i][0] * W[j][0] + U[i][1] * W[j][1] + d * U[i][2] * W[j][2];
        const double ti = mt[i] - (R[0] * ms[0] + R[1] * ms[1] + R[2] * ms[2]);
        T[i * 4 + 0] = (float)R[0]; T[i * 4 + 1] = (float)R[1]; T[i * 4 + 2] = (float)R[2];
        T[i * 4 + 3] = (float)ti;
    }
    T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
}

__global__ void __launch_bounds__(kIcpThreads) icp_update_kernel(const double* __restrict__ partials, uint32_t nblocks,
                                                                 IcpParams P, IcpState* __restrict__ st) {
    if (st->done) return;
    __shared__ double red[8 * kIcpMoments];
    double m[kIcpMoments];
#pragma unroll
    for (int k = 0; k < kIcpMoments; ++k) m[k] = 0.0;
    for (uint32_t b = threadIdx.x; b < nblocks; b += kIcpThreads) {
#pragma unroll
        for (int k = 0; k < kIcpMoments; ++k) m[k] += partials[(size_t)b * kIcpMoments + k];
    }
    block_sum<kIcpMoments>(m, red);
    if (threadIdx.x != 0) return;
    st->n_corr = (int)m[0];
    if (m[0] < 3.0) {                                  // min_number_correspondences_
        st->state = ICP_NO_CORRESPONDENCES;
        st->done = 1;
        return;
    }
    float T[16];
    umeyama_from_moments(m, T);
    float F[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = T[i * 4 + 0] * st->T_final[0 * 4 + j];
            s = s + T[i * 4 + 1] * st->T_final[1 * 4 + j];
            s = s + T[i * 4 + 2] * st->T_final[2 * 4 + j];
            s = s + T[i * 4 + 3] * st->T_final[3 * 4 + j];
            F[i * 4 + j] = s;
        }
    for (int k = 0; k < 16; ++k) { st->T_final[k] = F[k]; st->T_inc[k] = T[k]; }
    const int it = ++st->iterations;
    int state = ICP_NOT_CONVERGED;
    const double mse = m[1] / m[0];
    if (it >= P.max_iterations) state = ICP_ITERATIONS;
    else {
        const double cos_angle = 0.5 * (double)(T[0] + T[5] + T[10] - 1.0f);
        const double trans_sqr = (double)(T[3] * T[3] + T[7] * T[7] + T[11] * T[11]);
        if (cos_angle >= P.rot_thr && trans_sqr <= P.trans_thr) state = ICP_TRANSFORM;
        else {
            st->mse = mse;
            if (fabs(mse - st->mse_prev) < P.abs_mse) state = ICP_ABS_MSE;
            else if (fabs(mse - st->mse_prev) / st->mse_prev < P.rel_mse) state = ICP_REL_MSE;
            else st->mse_prev = mse;
        }
    }
    st->state = state;
    if (state != ICP_NOT_CONVERGED) st->done = 1;
}

// source <- T_inc * source (skipped once the loop is over: the current cloud is not used afterwards)
__global__ void __launch_bounds__(kIcpThreads) icp_transform_kernel(float4* __restrict__ cur, uint32_t n,
                                                                    const IcpState* __restrict__ st) {
    if (st->done) return;
    const uint32_t i = blockIdx.x * kIcpThreads + threadIdx.x;
    if (i >= n) return;
    const float* T = st->T_inc;
    const float4 p = cur[i];
    float4 o;
    o.x = T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3];
    o.y = T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7];
    o.z = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
    o.w = p.w;
    cur[i] = o;
}

// getFitnessScore: sum of squared 1-NN distances of T_final * src (no range cap), per block
__global__ void __launch_bounds__(kIcpThreads) icp_fitness_kernel(const float4* __restrict__ src, uint32_t n,
                                                                  GridView g, GridView gc,
                                                                  const IcpState* __restrict__ st,
                                                                  double* __restrict__ partials) {
    __shared__ double red[8];
    double v[1] = {0.0};
    const uint32_t i = blockIdx.x * kIcpThreads + threadIdx.x;
    if (i < n) {
        const float* T = st->T_final;
        const float4 p = src[i];
        const float x = T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3];
        const float y = T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7];
        const float z = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
        const Nn1Best b = nn1_search(g, gc, x, y, z, __int_as_float(0x7f800000));
        if (b.key != kKeyNone) v[0] = (double)b.d2;
    }
    block_sum<1>(v, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = v[0];
}

__global__ void __launch_bounds__(kIcpThreads) icp_fitness_reduce_kernel(const double* __restrict__ partials,
                                                                         uint32_t nblocks, IcpState* __restrict__ st) {
    __shared__ double red[8];
    double v[1] = {0.0};
    for (uint32_t b = threadIdx.x; b < nblocks; b += kIcpThreads) v[0] += partials[b];
    block_sum<1>(v, red);
    if (threadIdx.x == 0) st->fitness_sum = v[0];
}

}  // namespace lvreg
