// oracle_icp.cpp -- CPU restatement of the loop-closure registration:
//   MO:549-628  performLoopClosure      (submaps, pcl::IterativeClosestPoint settings, gates, pose correction)
//   MO:630-661  detectLoopClosureDistance
//   MO:719-741  loopFindNearKeyframes
// MO: = /root/reference/lidar_odometry/src/mapOptimization.cpp.
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// pcl::IterativeClosestPoint<PointXYZI, PointXYZI> (PCL 1.12, not vendored, not installed here) is
// restated from its published algorithm:
//   Registration::align                     source copied, guess = identity
//   IterativeClosestPoint::computeTransformation
//        loop { CorrespondenceEstimation::determineCorrespondences (1-NN in the target, kept when
//               d2 <= max_corr_dist^2);  < 3 correspondences -> not converged;
//               TransformationEstimationSVD (Umeyama without scaling) on the kept pairs;
//               source cloud <- T * source cloud;  final <- T * final;  ++iterations;
//               DefaultConvergenceCriteria }
//   DefaultConvergenceCriteria::hasConverged  iterations >= max -> converged (ITERATIONS);
//        cos_angle = 0.5 (trace(R) - 1) >= 1 - eps_T and |t|^2 <= eps_T -> TRANSFORM;
//        |mse - mse_prev| < 1e-12 -> ABS_MSE;  |mse - mse_prev| / mse_prev < eps_fit -> REL_MSE
//        (max_iterations_similar_transforms = 0, mse_prev starts at DBL_MAX,
//         mse = mean of the kept squared distances)
//   Registration::getFitnessScore            mean squared 1-NN distance of final * source, no range cap
//
// PARITY PIN STATUS: UNPINNED by the reference (no fixtures; PCL / Eigen absent).  Deliberate
// deviation, stated here and in DESIGN.md: Eigen::umeyama works in float with Eigen's (build-
// dependent) vectorised reduction order and Eigen::JacobiSVD; this restatement accumulates the
// means and the cross-covariance as raw moments in double and uses a one-sided Jacobi SVD in double,
// which makes the result independent of the summation order to ~1e-13 (so a CPU sum and a GPU tree
// reduction give the same float transform) and agrees with the float formulation to float rounding.
// Pinned to tolerance against numpy.linalg.svd (tests/test_oracle_pins.py).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "oracle.h"

void orc_kdtree_knn_one(const orc_kdtree* t, const float* q, int k, int32_t* idx, float* d2);

namespace {

// 4x4 float product, coefficient (i,j) accumulated over k in order (Eigen fixed-size product)
void mat4_mul(const float* A, const float* B, float* C) {
    float R[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = A[i * 4 + 0] * B[0 * 4 + j];
            s = s + A[i * 4 + 1] * B[1 * 4 + j];
            s = s + A[i * 4 + 2] * B[2 * 4 + j];
            s = s + A[i * 4 + 3] * B[3 * 4 + j];
            R[i * 4 + j] = s;
        }
    std::memcpy(C, R, sizeof(R));
}

inline void apply4(const float* T, const float* p, float* o) {
    // tr * (x, y, z, 1): columns accumulated left to right
    o[0] = T[0] * p[0] + T[1] * p[1] + T[2] * p[2] + T[3];
    o[1] = T[4] * p[0] + T[5] * p[1] + T[6] * p[2] + T[7];
    o[2] = T[8] * p[0] + T[9] * p[1] + T[10] * p[2] + T[11];
    o[3] = p[3];
}

}  // namespace

// Umeyama (no scaling) from raw moments.  mom = {n, -, sum src (3), sum tgt (3), sum tgt_i * src_j (9, row i)}
// laid out as mom[0] = n, mom[1] unused here (sum d2), mom[2..4], mom[5..7], mom[8..16].  T: row-major 4x4.
extern "C" void orc_umeyama_from_moments(const double* mom, float* T) {
    const double n = mom[0];
    double ms[3], mt[3], A[3][3], V[3][3];
    for (int i = 0; i < 3; ++i) { ms[i] = mom[2 + i] / n; mt[i] = mom[5 + i] / n; }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            A[i][j] = mom[8 + 3 * i + j] / n - mt[i] * ms[j];      // sigma = E[(t - mt)(s - ms)^T]
            V[i][j] = i == j ? 1.0 : 0.0;
        }
    // one-sided Jacobi: A V' = U S, columns of A made orthogonal
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
                if (gamma == 0.0 || std::fabs(gamma) <= 1e-15 * std::sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
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
    for (int j = 0; j < 3; ++j) sv[j] = std::sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
    // order the singular values descending (JacobiSVD's convention: S(2) multiplies the smallest)
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
        // third left vector: A's column when it carries signal, the cross product otherwise
        const double s = sv[ord[2]];
        if (s > 1e-12 * sv[ord[0]] && s > 0.0) {
            for (int i = 0; i < 3; ++i) U[i][2] = A[i][ord[2]] / s;
        } else {
            U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
            U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
            U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
        }
    }
    auto det3 = [](const double M[3][3]) {
        return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
               M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
    };
    const double d = det3(U) * det3(W) < 0.0 ? -1.0 : 1.0;
    double R[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = U[i][0] * W[j][0] + U[i][1] * W[j][1] + d * U[i][2] * W[j][2];
    for (int i = 0; i < 3; ++i) {
        const double ti = mt[i] - (R[i][0] * ms[0] + R[i][1] * ms[1] + R[i][2] * ms[2]);
        T[i * 4 + 0] = (float)R[i][0]; T[i * 4 + 1] = (float)R[i][1]; T[i * 4 + 2] = (float)R[i][2];
        T[i * 4 + 3] = (float)ti;
    }
    T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
}

extern "C" void orc_icp_default_params(orc_icp_params* p) {
    p->max_corr_dist = 30.0f;                // historyKeyframeSearchRadius * 2, MO:580, utility.h:296
    p->max_iterations = 100;                 // MO:581
    p->transformation_epsilon = 1e-6;        // MO:582
    p->euclidean_fitness_epsilon = 1e-6;     // MO:583
    p->num_threads = 1;
}

// 1-NN of every query in the target (exact, ties -> lower index); idx -1 when the target is empty
extern "C" void orc_nn1(const float* tgt, size_t nt, const float* q, size_t nq, int32_t* idx, float* d2,
                        int num_threads) {
    orc_kdtree* tree = orc_kdtree_build(tgt, nt);
    if (num_threads < 1) num_threads = 1;
    const long long cnt = (long long)nq;
#pragma omp parallel for num_threads(num_threads) schedule(static)
    for (long long i = 0; i < cnt; ++i) orc_kdtree_knn_one(tree, q + 4 * i, 1, idx + i, d2 + i);
    orc_kdtree_free(tree);
}

extern "C" void orc_icp_align(const float* src, size_t ns, const float* tgt, size_t nt, const orc_icp_params* P,
                              orc_icp_result* res) {
    std::memset(res, 0, sizeof(*res));
    for (int i = 0; i < 4; ++i) res->final_transformation[i * 5] = 1.f;
    res->fitness = std::numeric_limits<double>::max();
    if (ns == 0 || nt == 0) { res->state = ORC_ICP_NO_INPUT; return; }
    const int nth = P->num_threads < 1 ? 1 : P->num_threads;
    orc_kdtree* tree = orc_kdtree_build(tgt, nt);
    std::vector<float> cur(src, src + 4 * ns);
    std::vector<int32_t> nn(ns);
    std::vector<float> nd(ns);
    const double max_d2 = (double)P->max_corr_dist * (double)P->max_corr_dist;
    const double rot_thr = 1.0 - P->transformation_epsilon, trans_thr = P->transformation_epsilon;
    double mse_prev = std::numeric_limits<double>::max();
    float final_T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    int iterations = 0, state = ORC_ICP_NOT_CONVERGED;
    const long long cnt = (long long)ns;
    for (;;) {
#pragma omp parallel for num_threads(nth) schedule(static)
        for (long long i = 0; i < cnt; ++i) orc_kdtree_knn_one(tree, &cur[4 * i], 1, &nn[i], &nd[i]);
        double mom[17] = {0};
        for (size_t i = 0; i < ns; ++i) {
            if ((double)nd[i] > max_d2) continue;
            const float* s = &cur[4 * i];
            const float* t = tgt + 4 * (size_t)nn[i];
            mom[0] += 1.0;
            mom[1] += (double)nd[i];
            for (int a = 0; a < 3; ++a) { mom[2 + a] += (double)s[a]; mom[5 + a] += (double)t[a]; }
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) mom[8 + 3 * a + b] += (double)t[a] * (double)s[b];
        }
        res->n_correspondences = (int)mom[0];
        if (mom[0] < 3.0) { state = ORC_ICP_NO_CORRESPONDENCES; break; }
        float T[16];
        orc_umeyama_from_moments(mom, T);
#pragma omp parallel for num_threads(nth) schedule(static)
        for (long long i = 0; i < cnt; ++i) {
            float o[4];
            apply4(T, &cur[4 * i], o);
            std::memcpy(&cur[4 * i], o, sizeof(o));
        }
        mat4_mul(T, final_T, final_T);
        ++iterations;
        // DefaultConvergenceCriteria::hasConverged
        if (iterations >= P->max_iterations) { state = ORC_ICP_ITERATIONS; break; }
        const double cos_angle = 0.5 * (double)(T[0] + T[5] + T[10] - 1.0f);
        const double trans_sqr = (double)(T[3] * T[3] + T[7] * T[7] + T[11] * T[11]);
        if (cos_angle >= rot_thr && trans_sqr <= trans_thr) { state = ORC_ICP_TRANSFORM; break; }
        const double mse = mom[1] / mom[0];
        res->mse = mse;
        if (std::fabs(mse - mse_prev) < 1e-12) { state = ORC_ICP_ABS_MSE; break; }
        if (std::fabs(mse - mse_prev) / mse_prev < P->euclidean_fitness_epsilon) { state = ORC_ICP_REL_MSE; break; }
        mse_prev = mse;
    }
    res->iterations = iterations;
    res->state = state;
    res->converged = (state == ORC_ICP_ITERATIONS || state == ORC_ICP_TRANSFORM || state == ORC_ICP_ABS_MSE ||
                      state == ORC_ICP_REL_MSE)
                         ? 1
                         : 0;
    std::memcpy(res->final_transformation, final_T, sizeof(final_T));
    // getFitnessScore(): the ORIGINAL source under the final transformation
    {
        double sum = 0.0;
        size_t nr = 0;
#pragma omp parallel for num_threads(nth) schedule(static)
        for (long long i = 0; i < cnt; ++i) {
            float o[4];
            apply4(final_T, src + 4 * i, o);
            orc_kdtree_knn_one(tree, o, 1, &nn[i], &nd[i]);
        }
        for (size_t i = 0; i < ns; ++i) { sum += (double)nd[i]; ++nr; }
        res->fitness = nr ? sum / (double)nr : std::numeric_limits<double>::max();
    }
    orc_kdtree_free(tree);
}

// tCorrect = correction * pclPointToAffine3f(pose), then pcl::getTranslationAndEulerAngles.  MO:604-609
extern "C" void orc_correct_pose(const float* correction4x4, const float pose[6], float out[6]) {
    float T12[12], W[16], C[16];
    orc_pose_to_affine(pose, T12);
    std::memcpy(W, T12, sizeof(T12));
    W[12] = W[13] = W[14] = 0.f; W[15] = 1.f;
    mat4_mul(correction4x4, W, C);
    out[3] = C[3]; out[4] = C[7]; out[5] = C[11];
    out[0] = std::atan2(C[9], C[10]);
    out[1] = (float)std::asin((double)-C[8]);
    out[2] = std::atan2(C[4], C[0]);
}
