// oracle_depth.cpp -- CPU restatement of the visual side's LiDAR depth association:
//   feature_tracker/src/feature_tracker_node.cpp:273-375  lidar_callback: 0.2 m VoxelGrid of the new
//        cloud, camera-view filter, transform into the odometry frame, 5 s queue, fuse, 0.2 m VoxelGrid
//   feature_tracker/src/feature_tracker.h:115-300         DepthRegister::get_depth: cloud into the
//        camera frame, features onto the unit sphere, 360 x 360 range image keeping the closest point
//        per 0.5 deg bin, unit-sphere projection, 3-NN per feature, ray / plane intersection, clamps
// (FT: = feature_tracker/src/feature_tracker.h, FN: = feature_tracker/src/feature_tracker_node.cpp).
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// PARITY PIN STATUS: UNPINNED by the reference (no fixtures; PCL / FLANN / Eigen absent).  One
// deliberate deviation: the reference's atan2 calls resolve to glibc atan2f (<= 1 ulp, not
// necessarily correctly rounded); here -- and on the device -- atan2 is evaluated in double and
// rounded to float, which is the correctly rounded float result except for ~2^-28 of the inputs.
// The bin index computed from it can differ from the reference's only when atan2f itself is off by
// an ulp exactly at a bin boundary.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <deque>
#include <vector>

#include "oracle.h"

struct orc_depth {
    std::deque<std::vector<float>> cloudQueue;      // FN:343
    std::deque<double> timeQueue;                   // FN:344
    std::vector<float> depthCloud;                  // FN:362-371
};

extern "C" orc_depth* orc_depth_create(void) { return new orc_depth(); }
extern "C" void orc_depth_destroy(orc_depth* d) { delete d; }
extern "C" size_t orc_depth_cloud_size(const orc_depth* d) { return d->depthCloud.size() / 4; }
extern "C" void orc_depth_get_cloud(const orc_depth* d, float* out) {
    std::memcpy(out, d->depthCloud.data(), d->depthCloud.size() * sizeof(float));
}

static inline bool in_camera_view(const float* p) {
    // p.x >= 0 && abs(p.y / p.x) <= 10 && abs(p.z / p.x) <= 10   (FN:317, FT:178 negated)
    return p[0] >= 0.f && std::fabs(p[1] / p[0]) <= 10.f && std::fabs(p[2] / p[0]) <= 10.f;
}

// lidar_callback steps 2-9, FN:303-371
extern "C" size_t orc_depth_add_cloud(orc_depth* d, const float* cloud, size_t n, const float T[12], double stamp) {
    std::vector<float> ds(4 * (n ? n : 1));
    int pass;
    const size_t nds = n ? orc_voxelgrid(cloud, n, 0.2f, ds.data(), nullptr, nullptr, &pass) : 0;
    std::vector<float> filt;
    for (size_t i = 0; i < nds; ++i)
        if (in_camera_view(&ds[4 * i])) filt.insert(filt.end(), &ds[4 * i], &ds[4 * i] + 4);
    std::vector<float> glob(filt.size());
    orc_transform_cloud(filt.data(), filt.size() / 4, T, glob.data(), 1);
    d->cloudQueue.push_back(std::move(glob));
    d->timeQueue.push_back(stamp);
    while (!d->timeQueue.empty()) {
        if (stamp - d->timeQueue.front() > 5.0) {
            d->cloudQueue.pop_front();
            d->timeQueue.pop_front();
        } else break;
    }
    std::vector<float> fused;
    for (const auto& c : d->cloudQueue) fused.insert(fused.end(), c.begin(), c.end());
    d->depthCloud.assign(fused.size(), 0.f);
    const size_t m = fused.empty() ? 0 : orc_voxelgrid(fused.data(), fused.size() / 4, 0.2f, d->depthCloud.data(), nullptr, nullptr, &pass);
    d->depthCloud.resize(4 * m);
    return m;
}

static inline float atan2_rounded(float a, float b) { return (float)std::atan2((double)a, (double)b); }

// get_depth from step 0.4 on (FT:150-283).  depth_out[n] (-1 = no depth), feat3d_out (optional, n x 4:
// features_3d_sphere as published), local_out (optional, room for m rows: depth_cloud_local after the
// range-image filter).  Returns the size of the filtered local cloud.
extern "C" size_t orc_get_depth(const float* depth_cloud, size_t m, const float Tinv[12], const float* feat_xyz,
                                size_t n, int num_bins, float* depth_out, float* feat3d_out, float* local_out) {
    for (size_t i = 0; i < n; ++i) depth_out[i] = -1.f;
    std::vector<float> sphere_f(4 * n);
    for (size_t i = 0; i < n; ++i) {
        // Eigen::Vector3f::normalize(): divide by sqrt(squaredNorm)
        const float x = feat_xyz[3 * i], y = feat_xyz[3 * i + 1], z = feat_xyz[3 * i + 2];
        const float nrm = std::sqrt(x * x + y * y + z * z);
        const float fx = x / nrm, fy = y / nrm, fz = z / nrm;
        sphere_f[4 * i + 0] = fz;          // ROS convention, FT:163-165
        sphere_f[4 * i + 1] = -fx;
        sphere_f[4 * i + 2] = -fy;
        sphere_f[4 * i + 3] = -1.f;
    }
    auto finish = [&]() {
        if (feat3d_out) std::memcpy(feat3d_out, sphere_f.data(), sphere_f.size() * sizeof(float));
    };
    if (m == 0) { finish(); return 0; }
    std::vector<float> local(4 * m);
    orc_transform_cloud(depth_cloud, m, Tinv, local.data(), 1);
    // range image: closest point per bin, first one wins among equals (FT:170-196)
    const float bin_res = 180.0 / (float)num_bins;
    std::vector<float> range((size_t)num_bins * num_bins, FLT_MAX);
    std::vector<int> who((size_t)num_bins * num_bins, -1);
    for (size_t i = 0; i < m; ++i) {
        const float* p = &local[4 * i];
        if (p[0] < 0.f || std::fabs(p[1] / p[0]) > 10.f || std::fabs(p[2] / p[0]) > 10.f) continue;
        const float row_angle = (float)((double)atan2_rounded(p[2], std::sqrt(p[0] * p[0] + p[1] * p[1])) * 180.0 / M_PI + 90.0);
        const int row_id = (int)std::round(row_angle / bin_res);
        const float col_angle = (float)((double)atan2_rounded(p[0], p[1]) * 180.0 / M_PI);
        const int col_id = (int)std::round(col_angle / bin_res);
        if (row_id < 0 || row_id >= num_bins || col_id < 0 || col_id >= num_bins) continue;
        const float dist = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        const size_t b = (size_t)row_id * num_bins + col_id;
        if (dist < range[b]) { range[b] = dist; who[b] = (int)i; }
    }
    std::vector<float> kept, unit;
    for (size_t b = 0; b < range.size(); ++b)
        if (range[b] != FLT_MAX) {
            const float* p = &local[4 * (size_t)who[b]];
            kept.insert(kept.end(), p, p + 4);
            const float r = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
            unit.push_back(p[0] / r); unit.push_back(p[1] / r); unit.push_back(p[2] / r); unit.push_back(r);
        }
    const size_t nk = kept.size() / 4;
    if (local_out) std::memcpy(local_out, kept.data(), kept.size() * sizeof(float));
    if (nk < 10) { finish(); return nk; }                                        // FT:224-225
    orc_kdtree* tree = orc_kdtree_build(unit.data(), nk);
    const float thr = (float)std::pow(std::sin(bin_res / 180.0 * M_PI) * 5.0, 2);
    for (size_t i = 0; i < n; ++i) {
        int32_t idx[3];
        float d2[3];
        float* f = &sphere_f[4 * i];
        orc_kdtree_knn(tree, f, 1, 3, idx, d2, 1);
        if (idx[2] < 0 || !(d2[2] < thr)) continue;
        const float* u0 = &unit[4 * (size_t)idx[0]];
        const float* u1 = &unit[4 * (size_t)idx[1]];
        const float* u2 = &unit[4 * (size_t)idx[2]];
        const float r1 = u0[3], r2 = u1[3], r3 = u2[3];
        const float A[3] = {u0[0] * r1, u0[1] * r1, u0[2] * r1};
        const float B[3] = {u1[0] * r2, u1[1] * r2, u1[2] * r2};
        const float Cc[3] = {u2[0] * r3, u2[1] * r3, u2[2] * r3};
        const float ab[3] = {A[0] - B[0], A[1] - B[1], A[2] - B[2]};
        const float bc[3] = {B[0] - Cc[0], B[1] - Cc[1], B[2] - Cc[2]};
        const float N[3] = {ab[1] * bc[2] - ab[2] * bc[1], ab[2] * bc[0] - ab[0] * bc[2], ab[0] * bc[1] - ab[1] * bc[0]};
        float s = (N[0] * A[0] + N[1] * A[1] + N[2] * A[2]) / (N[0] * f[0] + N[1] * f[1] + N[2] * f[2]);
        const float min_depth = std::min(r1, std::min(r2, r3));
        const float max_depth = std::max(r1, std::max(r2, r3));
        if (max_depth - min_depth > 2 || s <= 0.5) continue;
        else if (s - max_depth > 0) s = max_depth;
        else if (s - min_depth < 0) s = min_depth;
        f[0] *= s; f[1] *= s; f[2] *= s;
        f[3] = f[0];
    }
    orc_kdtree_free(tree);
    for (size_t i = 0; i < n; ++i)
        if (sphere_f[4 * i + 3] > 3.0) depth_out[i] = sphere_f[4 * i + 3];
    finish();
    return nk;
}
