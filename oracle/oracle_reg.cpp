// oracle_reg.cpp -- TEST INFRASTRUCTURE (see oracle.h).
// Restatement of the registration loop of the reference:
//   cornerOptimization MO:1006-1096, surfOptimization MO:1098-1167,
//   combineOptimizationCoeffs MO:1169-1188, LMOptimization MO:1190-1313,
//   scan2MapOptimization MO:1315-1343, transformUpdate MO:1345-1385,
//   extractNearby / extractCloud MO:894-970, downsampleCurrentScan MO:987-999.
// Float / double promotions follow the reference source exactly:
// double literals promote their sub-expression, std::sqrt/fabs/sin/cos on float stay float.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

void orc_kdtree_knn_one(const orc_kdtree* t, const float* q, int k, int32_t* idx, float* d2);

extern "C" void orc_default_params(orc_params* p) {
    p->corner_leaf = 0.2f;
    p->surf_leaf = 0.4f;
    p->edge_min_valid = 10;
    p->surf_min_valid = 100;
    p->max_iters = 20;
    p->knn_gate_sq = 1.0f;
    p->line_eig_ratio = 3.0f;
    p->plane_tol = 0.2f;
    p->min_weight = 0.1f;
    p->min_matches = 50;
    p->degeneracy_eig = 100.0f;
    p->conv_deg = 0.05f;
    p->conv_cm = 0.05f;
    p->reference_quirks = 1;
    p->keyframe_search_radius = 50.0f;
    p->keyframe_density = 2.0f;
    p->rotation_tolerance = 1000.0f;
    p->z_tolerance = 1000.0f;
    p->num_threads = 8;
}

namespace {

inline void apply_affine(const float T[12], const float* pi, float* po) {
    po[0] = T[0] * pi[0] + T[1] * pi[1] + T[2] * pi[2] + T[3];
    po[1] = T[4] * pi[0] + T[5] * pi[1] + T[6] * pi[2] + T[7];
    po[2] = T[8] * pi[0] + T[9] * pi[1] + T[10] * pi[2] + T[11];
    po[3] = pi[3];
}

// one corner feature: returns 1 when the correspondence is accepted
inline int corner_one(const float* map, const int32_t* nn, const float* nd2, const float* ori,
                      const float* sel, const orc_params* P, float* coeff) {
    coeff[0] = coeff[1] = coeff[2] = coeff[3] = 0.0f;
    if (!(nd2[4] < (double)P->knn_gate_sq)) return 0;
    float cx = 0, cy = 0, cz = 0;
    for (int j = 0; j < 5; ++j) {
        cx += map[4 * (size_t)nn[j]];
        cy += map[4 * (size_t)nn[j] + 1];
        cz += map[4 * (size_t)nn[j] + 2];
    }
    cx /= 5; cy /= 5; cz /= 5;
    float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
    for (int j = 0; j < 5; ++j) {
        float ax = map[4 * (size_t)nn[j]] - cx;
        float ay = map[4 * (size_t)nn[j] + 1] - cy;
        float az = map[4 * (size_t)nn[j] + 2] - cz;
        a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
        a22 += ay * ay; a23 += ay * az;
        a33 += az * az;
    }
    a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
    float A[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33};
    float D[3], V[9];
    orc_jacobi_eigen(A, 3, D, V);
    if (!(D[0] > P->line_eig_ratio * D[1])) return 0;

    float x0 = sel[0], y0 = sel[1], z0 = sel[2];
    float x1 = (float)(cx + 0.1 * V[0]);
    float y1 = (float)(cy + 0.1 * V[1]);
    float z1 = (float)(cz + 0.1 * V[2]);
    float x2 = (float)(cx - 0.1 * V[0]);
    float y2 = (float)(cy - 0.1 * V[1]);
    float z2 = (float)(cz - 0.1 * V[2]);

    float m1 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
    float m2 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
    float m3 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
    float a012 = std::sqrt(m1 * m1 + m2 * m2 + m3 * m3);
    float l12 = std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    float la = ((y1 - y2) * m1 + (z1 - z2) * m2) / a012 / l12;
    float lb = -((x1 - x2) * m1 - (z1 - z2) * m3) / a012 / l12;
    float lc = -((x1 - x2) * m2 + (y1 - y2) * m3) / a012 / l12;
    float ld2 = a012 / l12;
    float s = (float)(1 - 0.9 * std::fabs(ld2));
    coeff[0] = s * la;
    coeff[1] = s * lb;
    coeff[2] = s * lc;
    coeff[3] = s * ld2;
    return (s > (double)P->min_weight) ? 1 : 0;
}

inline int surf_one(const float* map, const int32_t* nn, const float* nd2, const float* ori,
                    const float* sel, const orc_params* P, float* coeff) {
    coeff[0] = coeff[1] = coeff[2] = coeff[3] = 0.0f;
    if (!(nd2[4] < (double)P->knn_gate_sq)) return 0;
    float A[15], b[5] = {-1, -1, -1, -1, -1}, X[3];
    for (int j = 0; j < 5; ++j) {
        A[3 * j + 0] = map[4 * (size_t)nn[j]];
        A[3 * j + 1] = map[4 * (size_t)nn[j] + 1];
        A[3 * j + 2] = map[4 * (size_t)nn[j] + 2];
    }
    orc_colpiv_qr_solve_5x3(A, b, X);
    float pa = X[0], pb = X[1], pc = X[2], pd = 1;
    float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
    pa /= ps; pb /= ps; pc /= ps; pd /= ps;
    for (int j = 0; j < 5; ++j) {
        float v = std::fabs(pa * A[3 * j] + pb * A[3 * j + 1] + pc * A[3 * j + 2] + pd);
        if (v > (double)P->plane_tol) return 0;
        if (std::isnan(v)) { /* NaN compares false: the reference keeps the plane */ }
    }
    float pd2 = pa * sel[0] + pb * sel[1] + pc * sel[2] + pd;
    float s = (float)(1 - 0.9 * std::fabs(pd2) /
                              std::sqrt(std::sqrt(ori[0] * ori[0] + ori[1] * ori[1] + ori[2] * ori[2])));
    coeff[0] = s * pa;
    coeff[1] = s * pb;
    coeff[2] = s * pc;
    coeff[3] = s * pd2;
    return (s > (double)P->min_weight) ? 1 : 0;
}

typedef int (*one_fn)(const float*, const int32_t*, const float*, const float*, const float*,
                      const orc_params*, float*);

void residuals(one_fn fn, const float* map, size_t m, const orc_kdtree* tree, const float* pts,
               size_t n, const float pose[6], const orc_params* P, float* coeff, uint8_t* flag,
               int32_t* knn_idx) {
    float T[12];
    orc_pose_to_affine(pose, T);
    int nt = P->num_threads < 1 ? 1 : P->num_threads;
    const long long cnt = (long long)n;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (long long i = 0; i < cnt; ++i) {
        float sel[4];
        int32_t nn[5];
        float nd2[5];
        apply_affine(T, pts + 4 * i, sel);
        orc_kdtree_knn_one(tree, sel, 5, nn, nd2);
        if (knn_idx) std::memcpy(knn_idx + 5 * i, nn, sizeof(nn));
        int ok = 0;
        float c[4] = {0, 0, 0, 0};
        if (m >= 5) ok = fn(map, nn, nd2, pts + 4 * i, sel, P, c);
        // rejected points keep their coefficients zeroed so arrays are comparable
        if (!ok) c[0] = c[1] = c[2] = c[3] = 0.0f;
        std::memcpy(coeff + 4 * i, c, sizeof(c));
        flag[i] = (uint8_t)ok;
    }
}

}  // namespace

extern "C" void orc_corner_residuals(const float* map, size_t m, const orc_kdtree* tree,
                                     const float* pts, size_t n, const float pose[6],
                                     const orc_params* p, float* coeff, uint8_t* flag,
                                     int32_t* knn_idx) {
    residuals(corner_one, map, m, tree, pts, n, pose, p, coeff, flag, knn_idx);
}

extern "C" void orc_surf_residuals(const float* map, size_t m, const orc_kdtree* tree,
                                   const float* pts, size_t n, const float pose[6],
                                   const orc_params* p, float* coeff, uint8_t* flag,
                                   int32_t* knn_idx) {
    residuals(surf_one, map, m, tree, pts, n, pose, p, coeff, flag, knn_idx);
}

extern "C" void orc_jacobian_rows(const float* ori, const float* coeff, size_t n,
                                  const float pose[6], float* A, float* b) {
    // lidar -> camera axis naming of the reference (MO:1202-1207)
    float srx = std::sin(pose[1]), crx = std::cos(pose[1]);
    float sry = std::sin(pose[2]), cry = std::cos(pose[2]);
    float srz = std::sin(pose[0]), crz = std::cos(pose[0]);
    for (size_t i = 0; i < n; ++i) {
        float px = ori[4 * i + 1], py = ori[4 * i + 2], pz = ori[4 * i + 0];
        float cx = coeff[4 * i + 1], cy = coeff[4 * i + 2], cz = coeff[4 * i + 0];
        float ci = coeff[4 * i + 3];
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
        float* row = A + 6 * i;
        row[0] = arz; row[1] = arx; row[2] = ary;
        row[3] = cz;  row[4] = cx;  row[5] = cy;
        b[i] = -ci;
    }
}

extern "C" int orc_lm_step(const float* ori, const float* coeff, size_t n_sel, int iter,
                           float pose[6], orc_lm_state* st, const orc_params* P, float AtA_out[36],
                           float Atb_out[6], float x_out[6]) {
    if ((int)n_sel < P->min_matches) return 0;
    std::vector<float> A(n_sel * 6), b(n_sel);
    orc_jacobian_rows(ori, coeff, n_sel, pose, A.data(), b.data());
    float AtA[36], Atb[6], X[6];
    orc_normal_equations(A.data(), b.data(), (int)n_sel, AtA, Atb);
    orc_qr_solve(AtA, Atb, 6, 1, X);

    float matP_local[36];
    std::memset(matP_local, 0, sizeof(matP_local));     // the shadowing local cv::Mat matP
    if (iter == 0) {
        float E[6], V[36], V2[36];
        orc_jacobi_eigen(AtA, 6, E, V);
        std::memcpy(V2, V, sizeof(V));
        st->is_degenerate = 0;
        for (int i = 5; i >= 0; --i) {
            if (E[i] < P->degeneracy_eig) {
                for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0.0f;
                st->is_degenerate = 1;
            } else {
                break;
            }
        }
        orc_lu_solve(V, V2, 6, 6, matP_local);          // matV.inv() * matV2
        if (!P->reference_quirks) std::memcpy(st->matP, matP_local, sizeof(matP_local));
    }
    if (st->is_degenerate) {
        const float* Pm = P->reference_quirks ? matP_local : st->matP;
        float X2[6];
        std::memcpy(X2, X, sizeof(X));
        orc_gemm(Pm, X2, 6, 6, 1, X);
    }
    for (int i = 0; i < 6; ++i) pose[i] += X[i];
    if (AtA_out) std::memcpy(AtA_out, AtA, sizeof(AtA));
    if (Atb_out) std::memcpy(Atb_out, Atb, sizeof(Atb));
    if (x_out) std::memcpy(x_out, X, sizeof(X));

    const float r2d = 57.29578f;                         // pcl::rad2deg(float)
    double r0 = (double)(X[0] * r2d), r1 = (double)(X[1] * r2d), r2 = (double)(X[2] * r2d);
    double t0 = (double)(X[3] * 100), t1 = (double)(X[4] * 100), t2 = (double)(X[5] * 100);
    float deltaR = (float)std::sqrt(r0 * r0 + r1 * r1 + r2 * r2);
    float deltaT = (float)std::sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    return (deltaR < (double)P->conv_deg && deltaT < (double)P->conv_cm) ? 1 : 0;
}

extern "C" void orc_scan2map(const float* corner_map, size_t mc, const float* surf_map, size_t ms,
                             const float* corner, size_t nc, const float* surf, size_t ns,
                             float pose[6], orc_lm_state* st, const orc_params* P, orc_result* res) {
    std::memset(res, 0, sizeof(*res));
    res->degenerate = st->is_degenerate;
    if (!((int)nc > P->edge_min_valid && (int)ns > P->surf_min_valid)) {
        res->status = 1;
        return;
    }
    orc_kdtree* tc = orc_kdtree_build(corner_map, mc);   // kdtree->setInputCloud MO:1322-1323
    orc_kdtree* ts = orc_kdtree_build(surf_map, ms);
    std::vector<float> ccoef(nc * 4), scoef(ns * 4);
    std::vector<uint8_t> cflag(nc), sflag(ns);
    std::vector<float> ori, coef;
    ori.reserve((nc + ns) * 4);
    coef.reserve((nc + ns) * 4);
    int max_iters = std::min(P->max_iters, 32);
    for (int it = 0; it < max_iters; ++it) {
        orc_corner_residuals(corner_map, mc, tc, corner, nc, pose, P, ccoef.data(), cflag.data(), nullptr);
        orc_surf_residuals(surf_map, ms, ts, surf, ns, pose, P, scoef.data(), sflag.data(), nullptr);
        ori.clear();
        coef.clear();
        for (size_t i = 0; i < nc; ++i)
            if (cflag[i]) {
                ori.insert(ori.end(), corner + 4 * i, corner + 4 * i + 4);
                coef.insert(coef.end(), ccoef.begin() + 4 * i, ccoef.begin() + 4 * i + 4);
            }
        for (size_t i = 0; i < ns; ++i)
            if (sflag[i]) {
                ori.insert(ori.end(), surf + 4 * i, surf + 4 * i + 4);
                coef.insert(coef.end(), scoef.begin() + 4 * i, scoef.begin() + 4 * i + 4);
            }
        size_t nsel = ori.size() / 4;
        int conv = orc_lm_step(ori.data(), coef.data(), nsel, it, pose, st, P, nullptr, nullptr, nullptr);
        res->n_sel[it] = (int)nsel;
        std::memcpy(res->pose_iter[it], pose, sizeof(float) * 6);
        res->iterations = it + 1;
        if (conv) { res->converged = 1; break; }
    }
    res->degenerate = st->is_degenerate;
    orc_kdtree_free(tc);
    orc_kdtree_free(ts);
}

namespace {
struct Quat { double x, y, z, w; };
inline Quat quat_rpy(double roll, double pitch, double yaw) {
    double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
    double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp);
    double cr = std::cos(hr), sr = std::sin(hr);
    return Quat{sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy};
}
inline double qdot(const Quat& a, const Quat& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline Quat slerp(const Quat& a, const Quat& b, double t) {
    double s = std::sqrt(qdot(a, a) * qdot(b, b));
    double d = qdot(a, b);
    double theta = (d < 0 ? std::acos(-d / s) : std::acos(d / s));   // angleShortestPath / 2
    if (theta != 0.0) {
        double inv = 1.0 / std::sin(theta);
        double s0 = std::sin((1.0 - t) * theta), s1 = std::sin(t * theta);
        double sg = d < 0 ? -1.0 : 1.0;
        return Quat{(a.x * s0 + sg * b.x * s1) * inv, (a.y * s0 + sg * b.y * s1) * inv,
                    (a.z * s0 + sg * b.z * s1) * inv, (a.w * s0 + sg * b.w * s1) * inv};
    }
    return a;
}
inline void quat_to_rpy(const Quat& q, double& roll, double& pitch, double& yaw) {
    double d = qdot(q, q), s = 2.0 / d;
    double xs = q.x * s, ys = q.y * s, zs = q.z * s;
    double wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    double xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    double yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
    double m01 = xy - wz, m02 = xz + wy;
    if (std::fabs(m20) >= 1.0) {
        yaw = 0;
        double delta = std::atan2(m01, m02);
        if (m20 < 0) { pitch = M_PI / 2.0; roll = delta; }
        else { pitch = -M_PI / 2.0; roll = delta; }
    } else {
        pitch = -std::asin(m20);
        roll = std::atan2(m21 / std::cos(pitch), m22 / std::cos(pitch));
        yaw = std::atan2(m10 / std::cos(pitch), m00 / std::cos(pitch));
    }
}
inline float clampf(float v, float lim) {
    if (v < -lim) v = -lim;
    if (v > lim) v = lim;
    return v;
}
}  // namespace

extern "C" void orc_transform_update(float pose[6], int imu_available, float imu_roll,
                                     float imu_pitch, float imu_weight, const orc_params* P) {
    if (imu_available && std::fabs(imu_pitch) < 1.4) {
        double r, p, y;
        Quat q = slerp(quat_rpy(pose[0], 0, 0), quat_rpy(imu_roll, 0, 0), imu_weight);
        quat_to_rpy(q, r, p, y);
        pose[0] = (float)r;
        q = slerp(quat_rpy(0, pose[1], 0), quat_rpy(0, imu_pitch, 0), imu_weight);
        quat_to_rpy(q, r, p, y);
        pose[1] = (float)p;
    }
    pose[0] = clampf(pose[0], P->rotation_tolerance);
    pose[1] = clampf(pose[1], P->rotation_tolerance);
    pose[5] = clampf(pose[5], P->z_tolerance);
}

// ---------------------------------------------------------------------------------------
// mapOptimization-like object
// ---------------------------------------------------------------------------------------
struct orc_mo {
    orc_params P;
    std::vector<std::vector<float>> corner_kf, surf_kf;      // cornerCloudKeyFrames / surfCloudKeyFrames
    std::vector<float> poses;                                  // cloudKeyPoses6D: 6 floats {r,p,y,x,y,z}
    std::vector<double> times;
    std::map<int, std::pair<std::vector<float>, std::vector<float>>> cache;   // laserCloudMapContainer
    std::vector<float> corner_map_ds, surf_map_ds;             // laserCloud{Corner,Surf}FromMapDS
    orc_lm_state lm;
    std::vector<float> loop_cloud[2];                          // cureKeyframeCloud / prevKeyframeCloud
};

extern "C" orc_mo* orc_mo_create(const orc_params* p) {
    orc_mo* mo = new orc_mo();
    mo->P = *p;
    std::memset(&mo->lm, 0, sizeof(mo->lm));
    return mo;
}
extern "C" void orc_mo_destroy(orc_mo* mo) { delete mo; }

extern "C" int orc_mo_add_keyframe(orc_mo* mo, const float* corner, size_t nc, const float* surf,
                                   size_t ns, const float pose[6], double time) {
    mo->corner_kf.emplace_back(corner, corner + 4 * nc);
    mo->surf_kf.emplace_back(surf, surf + 4 * ns);
    mo->poses.insert(mo->poses.end(), pose, pose + 6);
    mo->times.push_back(time);
    return (int)mo->times.size() - 1;
}
extern "C" size_t orc_mo_num_keyframes(const orc_mo* mo) { return mo->times.size(); }

extern "C" size_t orc_mo_extract_nearby(orc_mo* mo, double time_now, int32_t* ids, size_t cap) {
    const size_t K = mo->times.size();
    if (K == 0) return 0;
    // cloudKeyPoses3D: {x,y,z,intensity = index}
    std::vector<float> kp(4 * K);
    for (size_t i = 0; i < K; ++i) {
        kp[4 * i] = mo->poses[6 * i + 3];
        kp[4 * i + 1] = mo->poses[6 * i + 4];
        kp[4 * i + 2] = mo->poses[6 * i + 5];
        kp[4 * i + 3] = (float)i;
    }
    orc_kdtree* t = orc_kdtree_build(kp.data(), K);
    std::vector<int32_t> hit(K);
    std::vector<float> hd(K);
    size_t nh = orc_kdtree_radius(t, &kp[4 * (K - 1)], mo->P.keyframe_search_radius, hit.data(), hd.data(), K);
    std::vector<float> sur(4 * nh), surds(4 * nh);
    for (size_t i = 0; i < nh; ++i) std::memcpy(&sur[4 * i], &kp[4 * (size_t)hit[i]], 16);
    int pass = 0;
    size_t nds = orc_voxelgrid(sur.data(), nh, mo->P.keyframe_density, surds.data(), nullptr, nullptr, &pass);
    std::vector<int32_t> out;
    for (size_t i = 0; i < nds; ++i) {
        int32_t nn;
        float d;
        orc_kdtree_knn(t, &surds[4 * i], 1, 1, &nn, &d, 1);
        surds[4 * i + 3] = kp[4 * (size_t)nn + 3];
    }
    // also the keyframes of the last 10 s
    std::vector<float> list(surds.begin(), surds.begin() + 4 * nds);
    for (long long i = (long long)K - 1; i >= 0; --i) {
        if (time_now - mo->times[i] < 10.0) list.insert(list.end(), &kp[4 * i], &kp[4 * i] + 4);
        else break;
    }
    orc_kdtree_free(t);
    // extractCloud's distance filter (MO:938-939) is applied here so that the id list is final
    const float* last = &kp[4 * (K - 1)];
    for (size_t i = 0; i < list.size() / 4; ++i) {
        const float* q = &list[4 * i];
        float dist = std::sqrt((q[0] - last[0]) * (q[0] - last[0]) + (q[1] - last[1]) * (q[1] - last[1]) +
                               (q[2] - last[2]) * (q[2] - last[2]));
        if (dist > mo->P.keyframe_search_radius) continue;
        out.push_back((int32_t)q[3]);
    }
    size_t n = std::min(cap, out.size());
    std::memcpy(ids, out.data(), n * sizeof(int32_t));
    return out.size();
}

extern "C" void orc_mo_build_local_map(orc_mo* mo, const int32_t* ids, size_t n) {
    std::vector<float> cmap, smap;
    for (size_t i = 0; i < n; ++i) {
        int id = ids[i];
        auto it = mo->cache.find(id);
        if (it == mo->cache.end()) {
            float T[12];
            // pclPointToAffine3f(cloudKeyPoses6D[id])
            orc_pose_to_affine(&mo->poses[6 * (size_t)id], T);
            std::vector<float> c(mo->corner_kf[id].size()), s(mo->surf_kf[id].size());
            orc_transform_cloud(mo->corner_kf[id].data(), c.size() / 4, T, c.data(), mo->P.num_threads);
            orc_transform_cloud(mo->surf_kf[id].data(), s.size() / 4, T, s.data(), mo->P.num_threads);
            it = mo->cache.emplace(id, std::make_pair(std::move(c), std::move(s))).first;
        }
        cmap.insert(cmap.end(), it->second.first.begin(), it->second.first.end());
        smap.insert(smap.end(), it->second.second.begin(), it->second.second.end());
    }
    int pass;
    mo->corner_map_ds.resize(cmap.size());
    size_t mc = orc_voxelgrid(cmap.data(), cmap.size() / 4, mo->P.corner_leaf, mo->corner_map_ds.data(), nullptr, nullptr, &pass);
    mo->corner_map_ds.resize(4 * mc);
    mo->surf_map_ds.resize(smap.size());
    size_t ms = orc_voxelgrid(smap.data(), smap.size() / 4, mo->P.surf_leaf, mo->surf_map_ds.data(), nullptr, nullptr, &pass);
    mo->surf_map_ds.resize(4 * ms);
    if (mo->cache.size() > 1000) mo->cache.clear();
}

extern "C" size_t orc_mo_map_size(const orc_mo* mo, int which) {
    return (which == 0 ? mo->corner_map_ds.size() : mo->surf_map_ds.size()) / 4;
}
extern "C" void orc_mo_get_map(const orc_mo* mo, int which, float* out) {
    const std::vector<float>& v = which == 0 ? mo->corner_map_ds : mo->surf_map_ds;
    std::memcpy(out, v.data(), v.size() * sizeof(float));
}

extern "C" void orc_mo_register_scan(orc_mo* mo, const float* corner_raw, size_t nc_raw,
                                     const float* surf_raw, size_t ns_raw, float pose[6],
                                     orc_result* res, size_t* nc_ds, size_t* ns_ds) {
    std::vector<float> cds(4 * nc_raw), sds(4 * ns_raw);
    int pass;
    size_t nc = orc_voxelgrid(corner_raw, nc_raw, mo->P.corner_leaf, cds.data(), nullptr, nullptr, &pass);
    size_t ns = orc_voxelgrid(surf_raw, ns_raw, mo->P.surf_leaf, sds.data(), nullptr, nullptr, &pass);
    if (nc_ds) *nc_ds = nc;
    if (ns_ds) *ns_ds = ns;
    if (mo->times.empty()) {
        std::memset(res, 0, sizeof(*res));
        res->status = 2;
        return;
    }
    orc_scan2map(mo->corner_map_ds.data(), mo->corner_map_ds.size() / 4, mo->surf_map_ds.data(),
                 mo->surf_map_ds.size() / 4, cds.data(), nc, sds.data(), ns, pose, &mo->lm, &mo->P, res);
    if (res->status == 0) orc_transform_update(pose, 0, 0, 0, 0, &mo->P);
}

// ---- loop closure ----------------------------------------------------------------
// loopFindNearKeyframes, MO:719-741: corner then surf of every keyframe in [key - n, key + n], each
// under its own stored pose, concatenated in that order, then downSizeFilterICP (leaf =
// mappingSurfLeafSize, MO:249)
extern "C" size_t orc_mo_loop_find_near_keyframes(orc_mo* mo, int key, int search_num, int slot) {
    std::vector<float> cat;
    const int K = (int)mo->times.size();
    for (int i = -search_num; i <= search_num; ++i) {
        const int k = key + i;
        if (k < 0 || k >= K) continue;
        float T[12];
        orc_pose_to_affine(&mo->poses[6 * (size_t)k], T);
        for (int which = 0; which < 2; ++which) {
            const std::vector<float>& src = which == 0 ? mo->corner_kf[k] : mo->surf_kf[k];
            const size_t at = cat.size();
            cat.resize(at + src.size());
            orc_transform_cloud(src.data(), src.size() / 4, T, cat.data() + at, mo->P.num_threads);
        }
    }
    std::vector<float>& out = mo->loop_cloud[slot ? 1 : 0];
    out.clear();
    if (cat.empty()) return 0;
    out.resize(cat.size());
    int pass;
    const size_t n = orc_voxelgrid(cat.data(), cat.size() / 4, mo->P.surf_leaf, out.data(), nullptr, nullptr, &pass);
    out.resize(4 * n);
    return n;
}
extern "C" void orc_mo_get_loop_cloud(const orc_mo* mo, int slot, float* out) {
    const std::vector<float>& v = mo->loop_cloud[slot ? 1 : 0];
    std::memcpy(out, v.data(), v.size() * sizeof(float));
}

// detectLoopClosureDistance, MO:630-661: nearest-first radius search around the last key pose, first hit
// older than time_diff
extern "C" int orc_mo_detect_loop_closure_distance(orc_mo* mo, double time_cur, float radius, float time_diff,
                                                   int* key_cur, int* key_pre) {
    const size_t K = mo->times.size();
    if (K == 0) return 0;
    std::vector<float> pos(4 * K);
    for (size_t i = 0; i < K; ++i) {
        pos[4 * i + 0] = mo->poses[6 * i + 3];
        pos[4 * i + 1] = mo->poses[6 * i + 4];
        pos[4 * i + 2] = mo->poses[6 * i + 5];
        pos[4 * i + 3] = (float)i;
    }
    orc_kdtree* tree = orc_kdtree_build(pos.data(), K);
    std::vector<int32_t> idx(K);
    std::vector<float> d2(K);
    const size_t found = orc_kdtree_radius(tree, &pos[4 * (K - 1)], radius, idx.data(), d2.data(), K);
    orc_kdtree_free(tree);
    int pre = -1;
    for (size_t i = 0; i < found && i < K; ++i)
        if (std::fabs(mo->times[(size_t)idx[i]] - time_cur) > (double)time_diff) { pre = idx[i]; break; }
    const int cur = (int)K - 1;
    if (pre == -1 || pre == cur) return 0;
    *key_cur = cur;
    *key_pre = pre;
    return 1;
}

// performLoopClosure from the submaps on, MO:566-613
extern "C" void orc_mo_perform_loop_closure(orc_mo* mo, int key_cur, int key_pre, int search_num,
                                            const orc_icp_params* P, float fitness_gate, orc_loop_result* out) {
    std::memset(out, 0, sizeof(*out));
    const size_t ns = orc_mo_loop_find_near_keyframes(mo, key_cur, 0, 0);
    const size_t nt = orc_mo_loop_find_near_keyframes(mo, key_pre, search_num, 1);
    out->n_source = (int)ns;
    out->n_target = (int)nt;
    if (ns < 300 || nt < 1000) { out->status = 1; return; }
    orc_icp_align(mo->loop_cloud[0].data(), ns, mo->loop_cloud[1].data(), nt, P, &out->icp);
    if (!out->icp.converged) { out->status = 2; return; }
    if (out->icp.fitness > (double)fitness_gate) { out->status = 3; return; }
    orc_correct_pose(out->icp.final_transformation, &mo->poses[6 * (size_t)key_cur], out->pose_from);
    std::memcpy(out->pose_to, &mo->poses[6 * (size_t)key_pre], 6 * sizeof(float));
    out->noise = (float)out->icp.fitness;
    out->status = 0;
}

// global map (publishGlobalMap MO:493-508, saveMapService MO:199-231): selected clouds of the listed
// keyframes under their stored poses, list order, one VoxelGrid; result in loop slot 0
extern "C" size_t orc_mo_build_global_map(orc_mo* mo, const int32_t* ids, size_t n_ids, int which, float leaf) {
    std::vector<float> cat;
    for (size_t i = 0; i < n_ids; ++i) {
        const int k = ids[i];
        float T[12];
        orc_pose_to_affine(&mo->poses[6 * (size_t)k], T);
        for (int w = 0; w < 2; ++w) {
            if (!(which & (1 << w))) continue;
            const std::vector<float>& src = w == 0 ? mo->corner_kf[k] : mo->surf_kf[k];
            const size_t at = cat.size();
            cat.resize(at + src.size());
            orc_transform_cloud(src.data(), src.size() / 4, T, cat.data() + at, mo->P.num_threads);
        }
    }
    std::vector<float>& out = mo->loop_cloud[0];
    out.clear();
    if (cat.empty()) return 0;
    out.resize(cat.size());
    int pass;
    const size_t n = orc_voxelgrid(cat.data(), cat.size() / 4, leaf, out.data(), nullptr, nullptr, &pass);
    out.resize(4 * n);
    return n;
}
