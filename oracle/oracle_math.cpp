// oracle_math.cpp -- TEST INFRASTRUCTURE (see oracle.h).  Restatements of the small dense
// routines the reference reaches through OpenCV and Eigen on the hot path:
//   cv::eigen            MO:1050 (3x3), MO:1268 (6x6)   -> orc_jacobi_eigen
//   cv::solve DECOMP_QR  MO:1260                        -> orc_qr_solve
//   matV.inv()*matV2     MO:1283 (lowers to solve/LU)   -> orc_lu_solve
//   cv::Mat products     MO:1257-1259, 1290             -> orc_gemm, orc_normal_equations
//   colPivHouseholderQr  MO:1128                        -> orc_colpiv_qr_solve_5x3
// OpenCV 4.x algorithms (core/src/lapack.cpp, matrix_decomp.cpp, matmul) and Eigen 3.4
// (QR/ColPivHouseholderQR.h, Householder/Householder.h) restated from their published
// sources; neither library is vendored by the reference nor installed as C++ here.
// The first three are checked bit-for-bit against the cv2 wheel in tests/test_oracle_pins.py.
//
// Everything is fp32 with one rounding per operation: build with -ffp-contract=off.
#include "oracle.h"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

namespace {

// OpenCV's own hypot (not libm's): scaled sqrt(1 + r*r).
inline float cv_hypot(float a, float b) {
    a = std::fabs(a);
    b = std::fabs(b);
    if (a > b) {
        b /= a;
        return a * std::sqrt(1.0f + b * b);
    }
    if (b > 0.0f) {
        a /= b;
        return b * std::sqrt(1.0f + a * a);
    }
    return 0.0f;
}

// index of the largest |A[r][c]| for c in (r, n)
inline int row_argmax(const float* A, int n, int r) {
    int m = r + 1;
    float mv = std::fabs(A[r * n + m]);
    for (int c = r + 2; c < n; ++c) {
        float v = std::fabs(A[r * n + c]);
        if (mv < v) { mv = v; m = c; }
    }
    return m;
}
// index of the largest |A[r][c]| for r in [0, c)
inline int col_argmax(const float* A, int n, int c) {
    int m = 0;
    float mv = std::fabs(A[c]);
    for (int r = 1; r < c; ++r) {
        float v = std::fabs(A[r * n + c]);
        if (mv < v) { mv = v; m = r; }
    }
    return m;
}

inline void givens(float& v0, float& v1, float c, float s) {
    float a0 = v0, b0 = v1;
    v0 = a0 * c - b0 * s;
    v1 = a0 * s + b0 * c;
}

}  // namespace

extern "C" void orc_jacobi_eigen(const float* Ain, int n, float* W, float* V) {
    std::vector<float> Abuf(Ain, Ain + (size_t)n * n);
    float* A = Abuf.data();
    const float eps = FLT_EPSILON;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0f : 0.0f;

    std::vector<int> indR(n, 0), indC(n, 0);
    for (int k = 0; k < n; ++k) {
        W[k] = A[k * n + k];
        if (k < n - 1) indR[k] = row_argmax(A, n, k);
        if (k > 0) indC[k] = col_argmax(A, n, k);
    }

    const int max_iters = n * n * 30;
    if (n > 1) {
        for (int it = 0; it < max_iters; ++it) {
            // largest off-diagonal entry via the row / column caches
            int k = 0;
            float mv = std::fabs(A[indR[0]]);
            for (int i = 1; i < n - 1; ++i) {
                float v = std::fabs(A[i * n + indR[i]]);
                if (mv < v) { mv = v; k = i; }
            }
            int l = indR[k];
            for (int i = 1; i < n; ++i) {
                float v = std::fabs(A[indC[i] * n + i]);
                if (mv < v) { mv = v; k = indC[i]; l = i; }
            }

            float p = A[k * n + l];
            if (std::fabs(p) <= eps) break;
            float y = (W[l] - W[k]) * 0.5f;
            float t = std::fabs(y) + cv_hypot(p, y);
            float s = cv_hypot(p, t);
            float c = t / s;
            s = p / s;
            t = (p / t) * p;
            if (y < 0.0f) { s = -s; t = -t; }
            A[k * n + l] = 0.0f;
            W[k] -= t;
            W[l] += t;

            for (int i = 0; i < k; ++i) givens(A[i * n + k], A[i * n + l], c, s);
            for (int i = k + 1; i < l; ++i) givens(A[k * n + i], A[i * n + l], c, s);
            for (int i = l + 1; i < n; ++i) givens(A[k * n + i], A[l * n + i], c, s);
            for (int i = 0; i < n; ++i) givens(V[k * n + i], V[l * n + i], c, s);

            for (int j = 0; j < 2; ++j) {
                int idx = (j == 0) ? k : l;
                if (idx < n - 1) indR[idx] = row_argmax(A, n, idx);
                if (idx > 0) indC[idx] = col_argmax(A, n, idx);
            }
        }
    }

    // selection sort, descending; eigenvector rows follow
    for (int k = 0; k < n - 1; ++k) {
        int m = k;
        for (int i = k + 1; i < n; ++i)
            if (W[m] < W[i]) m = i;
        if (k != m) {
            std::swap(W[m], W[k]);
            for (int i = 0; i < n; ++i) std::swap(V[m * n + i], V[k * n + i]);
        }
    }
}

// Householder QR solve as in OpenCV's hal::QR32f fallback (square case used at MO:1260).
extern "C" int orc_qr_solve(const float* Ain, const float* bin, int n, int nrhs, float* x) {
    const int m = n;
    std::vector<float> Abuf(Ain, Ain + (size_t)m * n), bbuf(bin, bin + (size_t)m * nrhs);
    std::vector<float> vl(m), hf(n);
    float* A = Abuf.data();
    float* b = bbuf.data();
    const float eps = FLT_EPSILON * 10.0f;  // hal::QR32f passes FLT_EPSILON*10

    for (int l = 0; l < n; ++l) {
        int sz = m - l;
        float nrm = 0.0f;
        for (int i = 0; i < sz; ++i) {
            vl[i] = A[(l + i) * n + l];
            nrm += vl[i] * vl[i];
        }
        float head = vl[0];
        float sgn = (vl[0] >= 0.0f) ? 1.0f : -1.0f;
        vl[0] = vl[0] + sgn * std::sqrt(nrm);
        nrm = std::sqrt(nrm + vl[0] * vl[0] - head * head);
        for (int i = 0; i < sz; ++i) vl[i] /= nrm;

        for (int j = l; j < n; ++j) {
            float dot = 0.0f;
            for (int i = l; i < m; ++i) dot += vl[i - l] * A[i * n + j];
            for (int i = l; i < m; ++i) A[i * n + j] -= 2 * vl[i - l] * dot;
        }
        hf[l] = vl[0] * vl[0];
        for (int i = 1; i < sz; ++i) A[(l + i) * n + l] = vl[i] / vl[0];
    }

    for (int l = 0; l < n; ++l) {
        vl[0] = 1.0f;
        for (int j = 1; j < m - l; ++j) vl[j] = A[(j + l) * n + l];
        for (int j = 0; j < nrhs; ++j) {
            float dot = 0.0f;
            for (int i = l; i < m; ++i) dot += vl[i - l] * b[i * nrhs + j];
            for (int i = l; i < m; ++i) b[i * nrhs + j] -= 2 * vl[i - l] * dot * hf[l];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        for (int j = n - 1; j > i; --j)
            for (int p = 0; p < nrhs; ++p) b[i * nrhs + p] -= b[j * nrhs + p] * A[i * n + j];
        if (std::fabs(A[i * n + i]) < eps) {
            std::memset(x, 0, sizeof(float) * (size_t)n * nrhs);  // cv::solve: dst = 0 on failure
            return 0;
        }
        for (int p = 0; p < nrhs; ++p) b[i * nrhs + p] /= A[i * n + i];
    }
    std::memcpy(x, b, sizeof(float) * (size_t)n * nrhs);
    return 1;
}

// LU with partial pivoting as in OpenCV's hal::LU32f fallback.
extern "C" int orc_lu_solve(const float* Ain, const float* Bin, int n, int nrhs, float* X) {
    std::vector<float> Abuf(Ain, Ain + (size_t)n * n), Bbuf(Bin, Bin + (size_t)n * nrhs);
    float* A = Abuf.data();
    float* b = Bbuf.data();
    const float eps = FLT_EPSILON * 10.0f;
    for (int i = 0; i < n; ++i) {
        int k = i;
        for (int j = i + 1; j < n; ++j)
            if (std::fabs(A[j * n + i]) > std::fabs(A[k * n + i])) k = j;
        if (std::fabs(A[k * n + i]) < eps) {
            std::memset(X, 0, sizeof(float) * (size_t)n * nrhs);
            return 0;
        }
        if (k != i) {
            for (int j = i; j < n; ++j) std::swap(A[i * n + j], A[k * n + j]);
            for (int j = 0; j < nrhs; ++j) std::swap(b[i * nrhs + j], b[k * nrhs + j]);
        }
        float d = -1 / A[i * n + i];
        for (int j = i + 1; j < n; ++j) {
            float alpha = A[j * n + i] * d;
            for (int c = i + 1; c < n; ++c) A[j * n + c] += alpha * A[i * n + c];
            for (int c = 0; c < nrhs; ++c) b[j * nrhs + c] += alpha * b[i * nrhs + c];
        }
    }
    for (int i = n - 1; i >= 0; --i)
        for (int j = 0; j < nrhs; ++j) {
            float s = b[i * nrhs + j];
            for (int k = i + 1; k < n; ++k) s -= A[i * n + k] * b[k * nrhs + j];
            b[i * nrhs + j] = s / A[i * n + i];
        }
    std::memcpy(X, b, sizeof(float) * (size_t)n * nrhs);
    return 1;
}

// cv::gemm on CV_32F uses double accumulators (GEMMSingleMul<float,double>), k in order.
extern "C" void orc_gemm(const float* A, const float* B, int m, int k, int n, float* C) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int t = 0; t < k; ++t) s += (double)A[i * k + t] * (double)B[t * n + j];
            C[i * n + j] = (float)s;
        }
}

extern "C" void orc_normal_equations(const float* A, const float* b, int nrows, float* AtA,
                                     float* Atb) {
    double acc[6][6] = {{0}};
    double accb[6] = {0};
    for (int r = 0; r < nrows; ++r) {
        const float* row = A + (size_t)r * 6;
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j < 6; ++j) acc[i][j] += (double)row[i] * (double)row[j];
            accb[i] += (double)row[i] * (double)b[r];
        }
    }
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) AtA[i * 6 + j] = (float)acc[i][j];
        Atb[i] = (float)accb[i];
    }
}

// Eigen 3.4 ColPivHouseholderQR<Matrix<float,5,3>> compute + solve, scalar evaluation order.
extern "C" void orc_colpiv_qr_solve_5x3(const float* Arm, const float* b5, float* x3) {
    const int rows = 5, cols = 3;
    float qr[5][3];
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) qr[i][j] = Arm[i * 3 + j];
    float hcoef[3] = {0, 0, 0};
    int perm[3] = {0, 1, 2};
    float norm_upd[3], norm_dir[3];
    for (int k = 0; k < cols; ++k) {
        float s = 0.0f;
        for (int i = 0; i < rows; ++i) s += qr[i][k] * qr[i][k];
        norm_dir[k] = norm_upd[k] = std::sqrt(s);
    }
    float maxn = norm_upd[0];
    for (int k = 1; k < cols; ++k) if (norm_upd[k] > maxn) maxn = norm_upd[k];
    float th = maxn * FLT_EPSILON / (float)rows;
    const float threshold_helper = th * th;
    const float downdate_threshold = std::sqrt(FLT_EPSILON);
    int nonzero_pivots = cols;

    for (int k = 0; k < cols; ++k) {
        int big = k;
        float bign = norm_upd[k];
        for (int j = k + 1; j < cols; ++j)
            if (norm_upd[j] > bign) { bign = norm_upd[j]; big = j; }
        float big_sq = bign * bign;
        if (nonzero_pivots == cols && big_sq < threshold_helper * (float)(rows - k)) nonzero_pivots = k;
        if (big != k) {
            for (int i = 0; i < rows; ++i) std::swap(qr[i][k], qr[i][big]);
            std::swap(norm_upd[k], norm_upd[big]);
            std::swap(norm_dir[k], norm_dir[big]);
            std::swap(perm[k], perm[big]);
        }
        // makeHouseholderInPlace on qr[k..rows-1][k]
        float tail_sq = 0.0f;
        for (int i = k + 1; i < rows; ++i) tail_sq += qr[i][k] * qr[i][k];
        float c0 = qr[k][k];
        float beta, tau;
        if (tail_sq <= FLT_MIN) {
            tau = 0.0f;
            beta = c0;
            for (int i = k + 1; i < rows; ++i) qr[i][k] = 0.0f;
        } else {
            beta = std::sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0.0f) beta = -beta;
            float denom = c0 - beta;
            for (int i = k + 1; i < rows; ++i) qr[i][k] = qr[i][k] / denom;
            tau = (beta - c0) / beta;
        }
        qr[k][k] = beta;
        hcoef[k] = tau;
        // apply H = I - tau v v^T (v = [1; essential]) to the trailing columns
        if (tau != 0.0f) {
            for (int j = k + 1; j < cols; ++j) {
                float tmp = 0.0f;
                for (int i = k + 1; i < rows; ++i) tmp += qr[i][k] * qr[i][j];
                tmp += qr[k][j];
                qr[k][j] -= tau * tmp;
                for (int i = k + 1; i < rows; ++i) qr[i][j] -= tau * qr[i][k] * tmp;
            }
        }
        for (int j = k + 1; j < cols; ++j) {
            if (norm_upd[j] != 0.0f) {
                float temp = std::fabs(qr[k][j]) / norm_upd[j];
                temp = (1.0f + temp) * (1.0f - temp);
                temp = temp < 0.0f ? 0.0f : temp;
                float ratio = norm_upd[j] / norm_dir[j];
                float temp2 = temp * (ratio * ratio);
                if (temp2 <= downdate_threshold) {
                    float s = 0.0f;
                    for (int i = k + 1; i < rows; ++i) s += qr[i][j] * qr[i][j];
                    norm_dir[j] = std::sqrt(s);
                    norm_upd[j] = norm_dir[j];
                } else {
                    norm_upd[j] *= std::sqrt(temp);
                }
            }
        }
    }

    // solve: c = Q^T b (first nonzero_pivots reflectors), back-substitute, un-permute
    float c[5];
    for (int i = 0; i < rows; ++i) c[i] = b5[i];
    x3[0] = x3[1] = x3[2] = 0.0f;
    if (nonzero_pivots == 0) return;
    for (int k = 0; k < nonzero_pivots; ++k) {
        float tau = hcoef[k];
        if (tau != 0.0f) {
            float tmp = 0.0f;
            for (int i = k + 1; i < rows; ++i) tmp += qr[i][k] * c[i];
            tmp += c[k];
            c[k] -= tau * tmp;
            for (int i = k + 1; i < rows; ++i) c[i] -= tau * qr[i][k] * tmp;
        }
    }
    for (int i = nonzero_pivots - 1; i >= 0; --i) {
        c[i] /= qr[i][i];
        for (int r = 0; r < i; ++r) c[r] -= qr[r][i] * c[i];   // column-oriented back substitution
    }
    for (int i = 0; i < nonzero_pivots; ++i) x3[perm[i]] = c[i];
}
