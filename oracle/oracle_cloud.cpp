// oracle_cloud.cpp -- TEST INFRASTRUCTURE (see oracle.h).
//   orc_pose_to_affine   pcl::getTransformation as used by trans2Affine3f / pclPointToAffine3f
//                        MO:399-407 (PCL common/impl/eigen.hpp, restated)
//   orc_transform_cloud  transformPointCloud MO:347-385 / pointAssociateToMap MO:339-345
//   orc_voxelgrid        pcl::VoxelGrid<PointXYZI>::applyFilter (PCL 1.12 filters/impl/voxel_grid.hpp,
//                        restated from SURVEY Appendix A.1) as used at MO:959-965, MO:991-997
// PCL is not vendored by the reference nor installed here; the voxel key arithmetic is
// cross-checked against numpy integer arithmetic in tests/test_oracle_pins.py.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

extern "C" void orc_pose_to_affine(const float pose[6], float T[12]) {
    const float roll = pose[0], pitch = pose[1], yaw = pose[2];
    // fp32 throughout: libm sinf/cosf, products rounded one at a time
    float A = std::cos(yaw), B = std::sin(yaw);
    float C = std::cos(pitch), D = std::sin(pitch);
    float E = std::cos(roll), F = std::sin(roll);
    float DE = D * E, DF = D * F;
    T[0] = A * C;  T[1] = A * DF - B * E;  T[2]  = B * F + A * DE;  T[3]  = pose[3];
    T[4] = B * C;  T[5] = A * E + B * DF;  T[6]  = B * DE - A * F;  T[7]  = pose[4];
    T[8] = -D;     T[9] = C * F;           T[10] = C * E;           T[11] = pose[5];
}

extern "C" void orc_transform_cloud(const float* in, size_t n, const float T[12], float* out,
                                    int num_threads) {
    if (num_threads < 1) num_threads = 1;
    const long long cnt = (long long)n;
#pragma omp parallel for num_threads(num_threads)
    for (long long i = 0; i < cnt; ++i) {
        const float x = in[4 * i], y = in[4 * i + 1], z = in[4 * i + 2];
        out[4 * i + 0] = T[0] * x + T[1] * y + T[2] * z + T[3];
        out[4 * i + 1] = T[4] * x + T[5] * y + T[6] * z + T[7];
        out[4 * i + 2] = T[8] * x + T[9] * y + T[10] * z + T[11];
        out[4 * i + 3] = in[4 * i + 3];
    }
}

namespace {
struct KeyIdx {
    uint32_t key;
    uint32_t idx;
};

// stable sort by key: LSD byte radix sort for large inputs (so that the timed CPU baseline is
// not handicapped against PCL 1.12's boost::sort::spreadsort::integer_sort), std::stable_sort
// for small ones.  Both give the same order.
void stable_sort_by_key(std::vector<KeyIdx>& kv) {
    const size_t n = kv.size();
    if (n < 4096) {
        std::stable_sort(kv.begin(), kv.end(),
                         [](const KeyIdx& a, const KeyIdx& b) { return a.key < b.key; });
        return;
    }
    std::vector<KeyIdx> tmp(n);
    KeyIdx* src = kv.data();
    KeyIdx* dst = tmp.data();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        size_t hist[256] = {0};
        for (size_t i = 0; i < n; ++i) ++hist[(src[i].key >> shift) & 255u];
        bool trivial = false;
        for (int b = 0; b < 256; ++b) if (hist[b] == n) trivial = true;
        if (trivial) continue;
        size_t sum = 0;
        for (int b = 0; b < 256; ++b) { size_t c = hist[b]; hist[b] = sum; sum += c; }
        for (size_t i = 0; i < n; ++i) dst[hist[(src[i].key >> shift) & 255u]++] = src[i];
        std::swap(src, dst);
    }
    if (src != kv.data()) std::memcpy(kv.data(), src, n * sizeof(KeyIdx));
}
}  // namespace

extern "C" size_t orc_voxelgrid(const float* in, size_t n, float leaf, float* out,
                                uint32_t* keys_out, uint32_t* out_keys, int* passthrough) {
    if (passthrough) *passthrough = 0;
    if (n == 0) return 0;
    const float inv = 1.0f / leaf;   // Array4f::Ones() / leaf_size_
    float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(),
                   std::numeric_limits<float>::max()};
    float mx[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(),
                   -std::numeric_limits<float>::max()};
    for (size_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            float v = in[4 * i + a];
            mn[a] = std::min(mn[a], v);
            mx[a] = std::max(mx[a], v);
        }
    int64_t d[3];
    for (int a = 0; a < 3; ++a) d[a] = (int64_t)((mx[a] - mn[a]) * inv) + 1;
    if (d[0] * d[1] * d[2] > (int64_t)std::numeric_limits<int32_t>::max()) {
        // "Leaf size is too small for the input dataset": output = input
        std::memcpy(out, in, sizeof(float) * 4 * n);
        if (passthrough) *passthrough = 1;
        if (keys_out) std::memset(keys_out, 0, sizeof(uint32_t) * n);
        return n;
    }
    int min_b[3], max_b[3], div_b[3];
    for (int a = 0; a < 3; ++a) {
        min_b[a] = (int)std::floor(mn[a] * inv);
        max_b[a] = (int)std::floor(mx[a] * inv);
        div_b[a] = max_b[a] - min_b[a] + 1;
    }
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};

    std::vector<KeyIdx> kv(n);
    for (size_t i = 0; i < n; ++i) {
        int ijk[3];
        for (int a = 0; a < 3; ++a)
            ijk[a] = (int)(std::floor(in[4 * i + a] * inv) - (float)min_b[a]);
        int idx = ijk[0] * mul[0] + ijk[1] * mul[1] + ijk[2] * mul[2];
        kv[i].key = (uint32_t)idx;
        kv[i].idx = (uint32_t)i;
        if (keys_out) keys_out[i] = (uint32_t)idx;
    }
    // PCL sorts with an unstable sort; the stable order (ties by input index) is one legal
    // instance and is what the CUDA path reproduces.
    stable_sort_by_key(kv);

    size_t m = 0;
    size_t i = 0;
    while (i < n) {
        size_t j = i;
        float sx = 0.0f, sy = 0.0f, sz = 0.0f, si = 0.0f;
        while (j < n && kv[j].key == kv[i].key) {
            const float* p = in + 4 * (size_t)kv[j].idx;
            sx += p[0]; sy += p[1]; sz += p[2]; si += p[3];
            ++j;
        }
        const float cnt = (float)(j - i);
        out[4 * m + 0] = sx / cnt;
        out[4 * m + 1] = sy / cnt;
        out[4 * m + 2] = sz / cnt;
        out[4 * m + 3] = si / cnt;
        if (out_keys) out_keys[m] = kv[i].key;
        ++m;
        i = j;
    }
    return m;
}
