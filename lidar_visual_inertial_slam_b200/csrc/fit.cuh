// fit.cuh -- per-point edge / plane fitting and the small dense solves of LMOptimization, as
// device functions evaluated in registers / local arrays by ONE thread per problem.
//
//   jacobi_eigen<N>     cv::eigen on symmetric CV_32F (MO:1050 3x3, MO:1268 6x6): classic Jacobi
//                       with row/column max-index caches, eigenvalues descending, vectors as rows
//   corner_residual     cornerOptimization body MO:1025-1092
//   plane_fit_5x3       Eigen colPivHouseholderQr().solve for Matrix<float,5,3> (MO:1128)
//   surf_residual       surfOptimization body MO:1121-1163
//   jacobian_row        LMOptimization rows MO:1222-1255
//   qr_solve6 / lu_solve6 / matvec6   cv::solve(DECOMP_QR) MO:1260, matV.inv()*matV2 MO:1283,
//                       matP*matX2 MO:1290
//   lm_solve            LMOptimization after the normal equations MO:1260-1311
//
// Every function performs the same fp32 (and, where the reference's double literals promote,
// fp64) operations in the same order as the reference's third-party routines; the library is
// built with -fmad=false so nothing is contracted.
#pragma once

#include <float.h>

#include "common.cuh"

namespace lvreg {

struct RegParams {
    float knn_gate_sq, line_eig_ratio, plane_tol, min_weight, degeneracy_eig, conv_deg, conv_cm;
    int min_matches, max_iters, reference_quirks;
};

__device__ __forceinline__ float cv_hypot(float a, float b) {
    a = fabsf(a);
    b = fabsf(b);
    if (a > b) {
        b /= a;
        return a * sqrtf(1.0f + b * b);
    }
    if (b > 0.0f) {
        a /= b;
        return b * sqrtf(1.0f + a * a);
    }
    return 0.0f;
}

template <int N>
__device__ __forceinline__ int row_argmax(const float* A, int r) {
    int m = r + 1;
    float mv = fabsf(A[r * N + m]);
    for (int c = r + 2; c < N; ++c) {
        float v = fabsf(A[r * N + c]);
        if (mv < v) { mv = v; m = c; }
    }
    return m;
}
template <int N>
__device__ __forceinline__ int col_argmax(const float* A, int c) {
    int m = 0;
    float mv = fabsf(A[c]);
    for (int r = 1; r < c; ++r) {
        float v = fabsf(A[r * N + c]);
        if (mv < v) { mv = v; m = r; }
    }
    return m;
}
__device__ __forceinline__ void givens(float& v0, float& v1, float c, float s) {
    float a0 = v0, b0 = v1;
    v0 = a0 * c - b0 * s;
    v1 = a0 * s + b0 * c;
}

// A is destroyed.  W: eigenvalues descending, V: eigenvectors as rows.
template <int N>
__device__ void jacobi_eigen(float* A, float* W, float* V) {
    int indR[N], indC[N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) V[i * N + j] = (i == j) ? 1.0f : 0.0f;
    for (int k = 0; k < N; ++k) {
        W[k] = A[k * N + k];
        indR[k] = 0; indC[k] = 0;
        if (k < N - 1) indR[k] = row_argmax<N>(A, k);
        if (k > 0) indC[k] = col_argmax<N>(A, k);
    }
    const int max_iters = N * N * 30;
    for (int it = 0; it < max_iters; ++it) {
        int k = 0;
        float mv = fabsf(A[indR[0]]);
        for (int i = 1; i < N - 1; ++i) {
            float v = fabsf(A[i * N + indR[i]]);
            if (mv < v) { mv = v; k = i; }
        }
        int l = indR[k];
        for (int i = 1; i < N; ++i) {
            float v = fabsf(A[indC[i] * N + i]);
            if (mv < v) { mv = v; k = indC[i]; l = i; }
        }
        float p = A[k * N + l];
        if (fabsf(p) <= FLT_EPSILON) break;
        float y = (W[l] - W[k]) * 0.5f;
        float t = fabsf(y) + cv_hypot(p, y);
        float s = cv_hypot(p, t);
        float c = t / s;
        s = p / s;
        t = (p / t) * p;
        if (y < 0.0f) { s = -s; t = -t; }
        A[k * N + l] = 0.0f;
        W[k] -= t;
        W[l] += t;
        for (int i = 0; i < k; ++i) givens(A[i * N + k], A[i * N + l], c, s);
        for (int i = k + 1; i < l; ++i) givens(A[k * N + i], A[i * N + l], c, s);
        for (int i = l + 1; i < N; ++i) givens(A[k * N + i], A[l * N + i], c, s);
        for (int i = 0; i < N; ++i) givens(V[k * N + i], V[l * N + i], c, s);
        for (int j = 0; j < 2; ++j) {
            int idx = (j == 0) ? k : l;
            if (idx < N - 1) indR[idx] = row_argmax<N>(A, idx);
            if (idx > 0) indC[idx] = col_argmax<N>(A, idx);
        }
    }
    for (int k = 0; k < N - 1; ++k) {
        int m = k;
        for (int i = k + 1; i < N; ++i)
            if (W[m] < W[i]) m = i;
        if (k != m) {
            float tw = W[m]; W[m] = W[k]; W[k] = tw;
            for (int i = 0; i < N; ++i) {
                float tv = V[m * N + i]; V[m * N + i] = V[k * N + i]; V[k * N + i] = tv;
            }
        }
    }
}

// cv::eigen for the 3x3 covariance of cornerOptimization (MO:1050), entirely in registers: the same
// operation sequence as jacobi_eigen<3> (pivot search through the row / column max-index caches,
// including their staleness, rotation order, descending sort), but the three possible pivots
// (0,1), (0,2), (1,2) are selected with predicated moves instead of indexing local arrays.  For
// N = 3 only the strict upper triangle a01, a02, a12 is ever read; indR[1] = 2 and indC[1] = 0 are
// constants, so the caches reduce to indR0 in {1, 2} and indC2 in {0, 1}.
// Outputs: w0 >= w1 (the two largest eigenvalues) and the eigenvector row of w0.
__device__ __forceinline__ void jacobi_eigen3_top(float a00, float a01, float a02, float a11, float a12,
                                                  float a22, float* w_first, float* w_second,
                                                  float* vx, float* vy, float* vz) {
    float w0 = a00, w1 = a11, w2 = a22;
    float v00 = 1.f, v01 = 0.f, v02 = 0.f, v10 = 0.f, v11 = 1.f, v12 = 0.f, v20 = 0.f, v21 = 0.f, v22 = 1.f;
    int indR0 = (fabsf(a01) < fabsf(a02)) ? 2 : 1;
    int indC2 = (fabsf(a02) < fabsf(a12)) ? 1 : 0;
#pragma unroll 1
    for (int it = 0; it < 3 * 3 * 30; ++it) {
        // pivot: rows 0, 1 through indR, then columns 1, 2 through indC (strict '<' keeps the first maximum)
        float mv = fabsf(indR0 == 1 ? a01 : a02);
        int k = 0;
        float v = fabsf(a12);
        if (mv < v) { mv = v; k = 1; }
        int l = (k == 0) ? indR0 : 2;
        v = fabsf(a01);
        if (mv < v) { mv = v; k = 0; l = 1; }
        v = fabsf(indC2 == 0 ? a02 : a12);
        if (mv < v) { mv = v; k = indC2; l = 2; }
        const bool p01 = (l == 1), p12 = (k == 1);            // else the pivot is (0,2)
        const float p = p01 ? a01 : (p12 ? a12 : a02);
        if (fabsf(p) <= FLT_EPSILON) break;
        const float wk = p12 ? w1 : w0, wl = p01 ? w1 : w2;
        const float y = (wl - wk) * 0.5f;
        float t = fabsf(y) + cv_hypot(p, y);
        float s = cv_hypot(p, t);
        const float c = t / s;
        s = p / s;
        t = (p / t) * p;
        if (y < 0.0f) { s = -s; t = -t; }
        // the two off-diagonal entries that are not the pivot rotate as one pair, in this order:
        // (0,1): (a02, a12)   (0,2): (a01, a12)   (1,2): (a01, a02)
        float ga = p01 ? a02 : a01, gb = p12 ? a02 : a12;
        givens(ga, gb, c, s);
        if (p01) { a01 = 0.0f; a02 = ga; a12 = gb; }
        else if (p12) { a12 = 0.0f; a01 = ga; a02 = gb; }
        else { a02 = 0.0f; a01 = ga; a12 = gb; }
        const float nwk = wk - t, nwl = wl + t;
        if (p12) w1 = nwk; else w0 = nwk;
        if (p01) w1 = nwl; else w2 = nwl;
        // eigenvector rows k and l
        float rk0 = p12 ? v10 : v00, rk1 = p12 ? v11 : v01, rk2 = p12 ? v12 : v02;
        float rl0 = p01 ? v10 : v20, rl1 = p01 ? v11 : v21, rl2 = p01 ? v12 : v22;
        givens(rk0, rl0, c, s);
        givens(rk1, rl1, c, s);
        givens(rk2, rl2, c, s);
        if (p12) { v10 = rk0; v11 = rk1; v12 = rk2; } else { v00 = rk0; v01 = rk1; v02 = rk2; }
        if (p01) { v10 = rl0; v11 = rl1; v12 = rl2; } else { v20 = rl0; v21 = rl1; v22 = rl2; }
        // caches of the touched rows / columns: indR[k] (only row 0 has a choice), indC[l] (only column 2)
        if (k == 0) indR0 = (fabsf(a01) < fabsf(a02)) ? 2 : 1;
        if (l == 2) indC2 = (fabsf(a02) < fabsf(a12)) ? 1 : 0;
    }
    // selection sort, descending, rows of V follow (only the first two values and the first row are used)
    {
        int m = 0;
        if (w0 < w1) m = 1;
        if ((m == 0 ? w0 : w1) < w2) m = 2;
        if (m == 1) {
            float tw = w1; w1 = w0; w0 = tw;
            float t0 = v10, t1 = v11, t2 = v12; v10 = v00; v11 = v01; v12 = v02; v00 = t0; v01 = t1; v02 = t2;
        } else if (m == 2) {
            float tw = w2; w2 = w0; w0 = tw;
            float t0 = v20, t1 = v21, t2 = v22; v20 = v00; v21 = v01; v22 = v02; v00 = t0; v01 = t1; v02 = t2;
        }
        if (w1 < w2) { float tw = w2; w2 = w1; w1 = tw; }
    }
    *w_first = w0; *w_second = w1;
    *vx = v00; *vy = v01; *vz = v02;
}

// cornerOptimization for one point.  nb = the 5 neighbours (x,y,z), sel = pointSel (map frame).
// Returns the acceptance flag; coeff = (s*la, s*lb, s*lc, s*ld2).
__device__ __forceinline__ bool corner_residual(const float (&nbx)[5], const float (&nby)[5],
                                                const float (&nbz)[5], float x0, float y0, float z0,
                                                const RegParams& P, float4* coeff) {
    float cx = 0, cy = 0, cz = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) { cx += nbx[j]; cy += nby[j]; cz += nbz[j]; }
    cx /= 5; cy /= 5; cz /= 5;
    float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        float ax = nbx[j] - cx, ay = nby[j] - cy, az = nbz[j] - cz;
        a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
        a22 += ay * ay; a23 += ay * az;
        a33 += az * az;
    }
    a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
    float D[2], V[3];
    jacobi_eigen3_top(a11, a12, a13, a22, a23, a33, &D[0], &D[1], &V[0], &V[1], &V[2]);
    if (!(D[0] > P.line_eig_ratio * D[1])) return false;

    // `cx + 0.1 * v` is evaluated in double in the reference (0.1 is a double literal)
    float x1 = (float)((double)cx + 0.1 * (double)V[0]);
    float y1 = (float)((double)cy + 0.1 * (double)V[1]);
    float z1 = (float)((double)cz + 0.1 * (double)V[2]);
    float x2 = (float)((double)cx - 0.1 * (double)V[0]);
    float y2 = (float)((double)cy - 0.1 * (double)V[1]);
    float z2 = (float)((double)cz - 0.1 * (double)V[2]);

    float m1 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
    float m2 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
    float m3 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
    float a012 = sqrtf(m1 * m1 + m2 * m2 + m3 * m3);
    float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    float la = ((y1 - y2) * m1 + (z1 - z2) * m2) / a012 / l12;
    float lb = -((x1 - x2) * m1 - (z1 - z2) * m3) / a012 / l12;
    float lc = -((x1 - x2) * m2 + (y1 - y2) * m3) / a012 / l12;
    float ld2 = a012 / l12;
    float s = (float)(1.0 - 0.9 * (double)fabsf(ld2));
    *coeff = make_float4(s * la, s * lb, s * lc, s * ld2);
    return (double)s > (double)P.min_weight;
}

// Eigen 3.4 ColPivHouseholderQR<Matrix<float,5,3>>::solve(b), scalar evaluation order.
__device__ __forceinline__ void plane_fit_5x3(const float (&nbx)[5], const float (&nby)[5],
                                              const float (&nbz)[5], float (&x3)[3]) {
    float qr[5][3];
#pragma unroll
    for (int i = 0; i < 5; ++i) { qr[i][0] = nbx[i]; qr[i][1] = nby[i]; qr[i][2] = nbz[i]; }
    float hcoef[3] = {0.f, 0.f, 0.f};
    int perm[3] = {0, 1, 2};
    float nu[3], nd[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 5; ++i) s += qr[i][k] * qr[i][k];
        nd[k] = nu[k] = sqrtf(s);
    }
    float maxn = nu[0];
    if (nu[1] > maxn) maxn = nu[1];
    if (nu[2] > maxn) maxn = nu[2];
    const float th = maxn * FLT_EPSILON / 5.0f;
    const float threshold_helper = th * th;
    const float downdate_threshold = sqrtf(FLT_EPSILON);
    int nonzero_pivots = 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int big = k;
        float bign = nu[k];
#pragma unroll
        for (int j = k + 1; j < 3; ++j)
            if (nu[j] > bign) { bign = nu[j]; big = j; }
        if (nonzero_pivots == 3 && bign * bign < threshold_helper * (float)(5 - k)) nonzero_pivots = k;
        if (big != k) {
#pragma unroll
            for (int j = k + 1; j < 3; ++j) {
                if (j == big) {
#pragma unroll
                    for (int i = 0; i < 5; ++i) { float tq = qr[i][k]; qr[i][k] = qr[i][j]; qr[i][j] = tq; }
                    float tn = nu[k]; nu[k] = nu[j]; nu[j] = tn;
                    tn = nd[k]; nd[k] = nd[j]; nd[j] = tn;
                    int tp = perm[k]; perm[k] = perm[j]; perm[j] = tp;
                }
            }
        }
        float tail_sq = 0.0f;
#pragma unroll
        for (int i = k + 1; i < 5; ++i) tail_sq += qr[i][k] * qr[i][k];
        const float c0 = qr[k][k];
        float beta, tau;
        if (tail_sq <= FLT_MIN) {
            tau = 0.0f;
            beta = c0;
#pragma unroll
            for (int i = k + 1; i < 5; ++i) qr[i][k] = 0.0f;
        } else {
            beta = sqrtf(c0 * c0 + tail_sq);
            if (c0 >= 0.0f) beta = -beta;
            const float denom = c0 - beta;
#pragma unroll
            for (int i = k + 1; i < 5; ++i) qr[i][k] = qr[i][k] / denom;
            tau = (beta - c0) / beta;
        }
        qr[k][k] = beta;
        hcoef[k] = tau;
        if (tau != 0.0f) {
#pragma unroll
            for (int j = k + 1; j < 3; ++j) {
                float tmp = 0.0f;
#pragma unroll
                for (int i = k + 1; i < 5; ++i) tmp += qr[i][k] * qr[i][j];
                tmp += qr[k][j];
                qr[k][j] -= tau * tmp;
#pragma unroll
                for (int i = k + 1; i < 5; ++i) qr[i][j] -= tau * qr[i][k] * tmp;
            }
        }
#pragma unroll
        for (int j = k + 1; j < 3; ++j) {
            if (nu[j] != 0.0f) {
                float temp = fabsf(qr[k][j]) / nu[j];
                temp = (1.0f + temp) * (1.0f - temp);
                temp = temp < 0.0f ? 0.0f : temp;
                const float ratio = nu[j] / nd[j];
                const float temp2 = temp * (ratio * ratio);
                if (temp2 <= downdate_threshold) {
                    float s = 0.0f;
#pragma unroll
                    for (int i = k + 1; i < 5; ++i) s += qr[i][j] * qr[i][j];
                    nd[j] = sqrtf(s);
                    nu[j] = nd[j];
                } else {
                    nu[j] *= sqrtf(temp);
                }
            }
        }
    }
    float c[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};       // matB0.fill(-1)
    x3[0] = x3[1] = x3[2] = 0.0f;
    if (nonzero_pivots == 0) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k < nonzero_pivots) {
            const float tau = hcoef[k];
            if (tau != 0.0f) {
                float tmp = 0.0f;
#pragma unroll
                for (int i = k + 1; i < 5; ++i) tmp += qr[i][k] * c[i];
                tmp += c[k];
                c[k] -= tau * tmp;
#pragma unroll
                for (int i = k + 1; i < 5; ++i) c[i] -= tau * qr[i][k] * tmp;
            }
        }
    }
#pragma unroll
    for (int i = 2; i >= 0; --i) {
        if (i < nonzero_pivots) {
            c[i] /= qr[i][i];
#pragma unroll
            for (int r = 0; r < i; ++r) c[r] -= qr[r][i] * c[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < nonzero_pivots) {
            // x3[perm[i]] = c[i] without dynamic register indexing
            if (perm[i] == 0) x3[0] = c[i];
            else if (perm[i] == 1) x3[1] = c[i];
            else x3[2] = c[i];
        }
    }
}

// surfOptimization for one point.  ox,oy,oz = pointOri (sensor frame), sx,sy,sz = pointSel.
__device__ __forceinline__ bool surf_residual(const float (&nbx)[5], const float (&nby)[5],
                                              const float (&nbz)[5], float ox, float oy, float oz,
                                              float sx, float sy, float sz, const RegParams& P,
                                              float4* coeff) {
    float X[3];
    plane_fit_5x3(nbx, nby, nbz, X);
    float pa = X[0], pb = X[1], pc = X[2], pd = 1.0f;
    const float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa /= ps; pb /= ps; pc /= ps; pd /= ps;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        float v = fabsf(pa * nbx[j] + pb * nby[j] + pc * nbz[j] + pd);
        if ((double)v > (double)P.plane_tol) return false;
    }
    const float pd2 = pa * sx + pb * sy + pc * sz + pd;
    const float rng = sqrtf(sqrtf(ox * ox + oy * oy + oz * oz));
    const float s = (float)(1.0 - 0.9 * (double)fabsf(pd2) / (double)rng);
    *coeff = make_float4(s * pa, s * pb, s * pc, s * pd2);
    return (double)s > (double)P.min_weight;
}

// sin/cos of the pose angles in the reference's camera-axis naming (MO:1202-1207)
struct Trig { float srx, crx, sry, cry, srz, crz; };

// One row of matA / matB (MO:1222-1255): J = [arz, arx, ary, coeff.z, coeff.x, coeff.y], r = -d.
__device__ __forceinline__ void jacobian_row(const Trig& g, float orix, float oriy, float oriz,
                                             const float4& cf, float (&row)[7]) {
    const float px = oriy, py = oriz, pz = orix;          // lidar -> camera
    const float cx = cf.y, cy = cf.z, cz = cf.x;
    const float srx = g.srx, crx = g.crx, sry = g.sry, cry = g.cry, srz = g.srz, crz = g.crz;
    float arx = (crx * sry * srz * px + crx * crz * sry * py - srx * sry * pz) * cx
              + (-srx * srz * px - crz * srx * py - crx * pz) * cy
              + (crx * cry * srz * px + crx * cry * crz * py - cry * srx * pz) * cz;
    float ary = ((cry * srx * srz - crz * sry) * px
              + (sry * srz + cry * crz * srx) * py + crx * cry * pz) * cx
              + ((-cry * crz - srx * sry * srz) * px
              + (cry * srz - crz * srx * sry) * py - crx * sry * pz) * cz;
    float arz = ((crz * srx * sry - cry * srz) * px + (-cry * crz - srx * sry * srz) * py) * cx
              + (crx * crz * px - crx * srz * py) * cy
              + ((sry * srz + cry * crz * srx) * px + (crz * sry - cry * srx * srz) * py) * cz;
    row[0] = arz; row[1] = arx; row[2] = ary;
    row[3] = cz;  row[4] = cx;  row[5] = cy;
    row[6] = -cf.w;
}

// cv::solve(A, b, x, DECOMP_QR) for 6x6 CV_32F (OpenCV's Householder QR fallback, eps = 10*FLT_EPSILON).
// Returns false (x = 0) when a diagonal of R is below eps, as cv::solve does.
__device__ inline bool qr_solve6(const float* Ain, const float* bin, float* x) {
    constexpr int n = 6;
    float A[36], b[6], vl[6], hf[6];
#pragma unroll
    for (int i = 0; i < 36; ++i) A[i] = Ain[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) b[i] = bin[i];
#pragma unroll
    for (int l = 0; l < n; ++l) {
        const int sz = n - l;
        float nrm = 0.0f;
#pragma unroll
        for (int i = 0; i < sz; ++i) {
            vl[i] = A[(l + i) * n + l];
            nrm += vl[i] * vl[i];
        }
        const float head = vl[0];
        const float sgn = (vl[0] >= 0.0f) ? 1.0f : -1.0f;
        vl[0] = vl[0] + sgn * sqrtf(nrm);
        nrm = sqrtf(nrm + vl[0] * vl[0] - head * head);
#pragma unroll
        for (int i = 0; i < sz; ++i) vl[i] /= nrm;
#pragma unroll
        for (int j = l; j < n; ++j) {
            float dot = 0.0f;
#pragma unroll
            for (int i = l; i < n; ++i) dot += vl[i - l] * A[i * n + j];
#pragma unroll
            for (int i = l; i < n; ++i) A[i * n + j] -= 2 * vl[i - l] * dot;
        }
        hf[l] = vl[0] * vl[0];
#pragma unroll
        for (int i = 1; i < sz; ++i) A[(l + i) * n + l] = vl[i] / vl[0];
    }
#pragma unroll
    for (int l = 0; l < n; ++l) {
        vl[0] = 1.0f;
#pragma unroll
        for (int j = 1; j < n - l; ++j) vl[j] = A[(j + l) * n + l];
        float dot = 0.0f;
#pragma unroll
        for (int i = l; i < n; ++i) dot += vl[i - l] * b[i];
#pragma unroll
        for (int i = l; i < n; ++i) b[i] -= 2 * vl[i - l] * dot * hf[l];
    }
#pragma unroll
    for (int i = n - 1; i >= 0; --i) {
#pragma unroll
        for (int j = n - 1; j > i; --j) b[i] -= b[j] * A[i * n + j];
        if (fabsf(A[i * n + i]) < FLT_EPSILON * 10.0f) {
            for (int k = 0; k < 6; ++k) x[k] = 0.0f;
            return false;
        }
        b[i] /= A[i * n + i];
    }
    for (int i = 0; i < 6; ++i) x[i] = b[i];
    return true;
}

// cv::solve(A, B, X, DECOMP_LU) for 6x6 with 6 right-hand sides (what matV.inv()*matV2 lowers to)
__device__ inline bool lu_solve6(const float* Ain, const float* Bin, float* X) {
    constexpr int n = 6;
    float A[36], b[36];
    for (int i = 0; i < 36; ++i) { A[i] = Ain[i]; b[i] = Bin[i]; }
    for (int i = 0; i < n; ++i) {
        int k = i;
        for (int j = i + 1; j < n; ++j)
            if (fabsf(A[j * n + i]) > fabsf(A[k * n + i])) k = j;
        if (fabsf(A[k * n + i]) < FLT_EPSILON * 10.0f) {
            for (int t = 0; t < 36; ++t) X[t] = 0.0f;
            return false;
        }
        if (k != i) {
            for (int j = i; j < n; ++j) { float t = A[i * n + j]; A[i * n + j] = A[k * n + j]; A[k * n + j] = t; }
            for (int j = 0; j < n; ++j) { float t = b[i * n + j]; b[i * n + j] = b[k * n + j]; b[k * n + j] = t; }
        }
        const float d = -1 / A[i * n + i];
        for (int j = i + 1; j < n; ++j) {
            const float alpha = A[j * n + i] * d;
            for (int c = i + 1; c < n; ++c) A[j * n + c] += alpha * A[i * n + c];
            for (int c = 0; c < n; ++c) b[j * n + c] += alpha * b[i * n + c];
        }
    }
    for (int i = n - 1; i >= 0; --i)
        for (int j = 0; j < n; ++j) {
            float s = b[i * n + j];
            for (int k = i + 1; k < n; ++k) s -= A[i * n + k] * b[k * n + j];
            b[i * n + j] = s / A[i * n + i];
        }
    for (int t = 0; t < 36; ++t) X[t] = b[t];
    return true;
}

// Shortcut for the degeneracy test (MO:1262-1281).  If JtJ - mu*I admits a Cholesky factorisation, then
// lambda_min(JtJ) > mu - |E|, |E| <= 7 eps |JtJ| being the backward error of a 6x6 fp32 Cholesky.  With
// mu = 1.01 * threshold + 128 eps * trace this leaves lambda_min > threshold + 121 eps * trace, far above the error
// of the fp32 Jacobi sweep cv::eigen would run (a few eps * |JtJ|): its eigenvalues would all be >= threshold,
// i.e. isDegenerate = false and matP unused -- the sweep can be skipped.  Returns false when in doubt.  (fp32 on
// purpose: the fp64 version was ~2500 instructions of DDIV / DSQRT expansions executed once per registration, i.e.
// always from a cold instruction cache: 12 us at C3.)
__device__ inline bool clearly_well_conditioned(const float* AtA, float threshold) {
    float tr = 0.0f;
    for (int i = 0; i < 6; ++i) tr += AtA[i * 6 + i];
    if (!(tr > 0.0f) || !(tr < 1e30f)) return false;
    const float mu = threshold * 1.01f + 128.0f * 1.1920929e-07f * tr;
    // rolled on purpose: this runs once per registration, i.e. from a cold instruction cache, where the cost is the
    // number of instruction lines fetched (an unrolled version measured 12 us at C3), not the number executed
    float L[36];
#pragma unroll 1
    for (int j = 0; j < 6; ++j) {
        float d = AtA[j * 6 + j] - mu;
#pragma unroll 1
        for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
        if (!(d > 1e-5f * tr)) return false;
        const float ljj = sqrtf(d);
        L[j * 6 + j] = ljj;
#pragma unroll 1
        for (int i = j + 1; i < 6; ++i) {
            float v = 0.5f * (AtA[i * 6 + j] + AtA[j * 6 + i]);
#pragma unroll 1
            for (int k = 0; k < j; ++k) v -= L[i * 6 + k] * L[j * 6 + k];
            L[i * 6 + j] = v / ljj;
        }
    }
    return true;
}

struct LmState {            // isDegenerate (MO:131) and matP (MO:132) persist across scans
    int is_degenerate;
    float matP[36];
    int last_path;          // diagnostics: how iteration 0 decided degeneracy (1 = Cholesky shortcut, 2 = 6x6 Jacobi)
};

// LMOptimization from the solve onward (MO:1260-1311).  AtA/Atb are the fp32 normal equations.
// Updates pose, returns true when converged.
__device__ inline bool lm_solve(const float* AtA, const float* Atb, int iter, float* pose,
                                LmState* st, const RegParams& P, float* x_out) {
    float X[6];
    qr_solve6(AtA, Atb, X);
    float matP_local[36];
    for (int i = 0; i < 36; ++i) matP_local[i] = 0.0f;     // the shadowing local cv::Mat matP (MO:1220)
    if (iter == 0 && clearly_well_conditioned(AtA, P.degeneracy_eig)) {
        // every eigenvalue is provably above the threshold: cv::eigen would report the same
        // (isDegenerate = false, matP unused), so the 6x6 Jacobi sweep is skipped
        st->is_degenerate = 0;
        st->last_path = 1;
    } else if (iter == 0) {
        st->last_path = 2;
        float A[36], E[6], V[36], V2[36];
        for (int i = 0; i < 36; ++i) A[i] = AtA[i];
        jacobi_eigen<6>(A, E, V);
        for (int i = 0; i < 36; ++i) V2[i] = V[i];
        st->is_degenerate = 0;
        for (int i = 5; i >= 0; --i) {
            if (E[i] < P.degeneracy_eig) {
                for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0.0f;
                st->is_degenerate = 1;
            } else {
                break;
            }
        }
        if (st->is_degenerate) {          // matP only matters when degenerate
            lu_solve6(V, V2, matP_local);
            if (!P.reference_quirks)
                for (int i = 0; i < 36; ++i) st->matP[i] = matP_local[i];
        }
    }
    if (st->is_degenerate) {
        const float* Pm = P.reference_quirks ? matP_local : st->matP;
        float X2[6];
        for (int i = 0; i < 6; ++i) X2[i] = X[i];
        for (int i = 0; i < 6; ++i) {       // cv::gemm: double accumulators
            double s = 0.0;
            for (int t = 0; t < 6; ++t) s += (double)Pm[i * 6 + t] * (double)X2[t];
            X[i] = (float)s;
        }
    }
    for (int i = 0; i < 6; ++i) pose[i] += X[i];
    if (x_out) for (int i = 0; i < 6; ++i) x_out[i] = X[i];
    const float r2d = 57.29578f;           // pcl::rad2deg(float)
    const double r0 = (double)(X[0] * r2d), r1 = (double)(X[1] * r2d), r2 = (double)(X[2] * r2d);
    const double t0 = (double)(X[3] * 100), t1 = (double)(X[4] * 100), t2 = (double)(X[5] * 100);
    const float deltaR = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2);
    const float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    return (double)deltaR < (double)P.conv_deg && (double)deltaT < (double)P.conv_cm;
}

// pose -> affine and trig on the device (iterations >= 1 of the fused loop).  sin/cos are taken
// in double and rounded to float, which reproduces glibc's sinf/cosf except on rare
// last-bit rounding cases (tolerance item).
__device__ inline void pose_to_affine_dev(const float* pose, Affine* T, Trig* g) {
    const float roll = pose[0], pitch = pose[1], yaw = pose[2];
    const float A = (float)cos((double)yaw), B = (float)sin((double)yaw);
    const float C = (float)cos((double)pitch), D = (float)sin((double)pitch);
    const float E = (float)cos((double)roll), F = (float)sin((double)roll);
    const float DE = D * E, DF = D * F;
    T->m[0] = A * C;  T->m[1] = A * DF - B * E;  T->m[2]  = B * F + A * DE;  T->m[3]  = pose[3];
    T->m[4] = B * C;  T->m[5] = A * E + B * DF;  T->m[6]  = B * DE - A * F;  T->m[7]  = pose[4];
    T->m[8] = -D;     T->m[9] = C * F;           T->m[10] = C * E;           T->m[11] = pose[5];
    // MO:1202-1207: srx = sin(pitch), sry = sin(yaw), srz = sin(roll)
    g->srx = D; g->crx = C; g->sry = B; g->cry = A; g->srz = F; g->crz = E;
}

// The same, by the first six lanes of a warp: lane i evaluates one of the six sin / cos values (a double-precision
// sin or cos is ~150 dependent instructions; one after the other they cost ~2 us at the head of every LM
// iteration), lane 0 assembles the matrix.  Identical values: every lane runs the same scalar routine on the same
// argument as pose_to_affine_dev.  Call with the full warp converged.
__device__ inline void pose_to_affine_warp(const float* pose, Affine* T, Trig* g) {
    const int lane = threadIdx.x & 31;
    float v = 0.f;
    if (lane < 6) {
        const double ang = (double)pose[2 - (lane >> 1)];          // lanes 0,1: yaw; 2,3: pitch; 4,5: roll
        v = (lane & 1) ? (float)sin(ang) : (float)cos(ang);
    }
    const float A = __shfl_sync(0xffffffffu, v, 0), B = __shfl_sync(0xffffffffu, v, 1);
    const float C = __shfl_sync(0xffffffffu, v, 2), D = __shfl_sync(0xffffffffu, v, 3);
    const float E = __shfl_sync(0xffffffffu, v, 4), F = __shfl_sync(0xffffffffu, v, 5);
    if (lane == 0) {
        const float DE = D * E, DF = D * F;
        T->m[0] = A * C;  T->m[1] = A * DF - B * E;  T->m[2]  = B * F + A * DE;  T->m[3]  = pose[3];
        T->m[4] = B * C;  T->m[5] = A * E + B * DF;  T->m[6]  = B * DE - A * F;  T->m[7]  = pose[4];
        T->m[8] = -D;     T->m[9] = C * F;           T->m[10] = C * E;           T->m[11] = pose[5];
        g->srx = D; g->crx = C; g->sry = B; g->cry = A; g->srz = F; g->crz = E;
    }
}

}  // namespace lvreg
