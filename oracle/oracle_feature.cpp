// oracle_feature.cpp -- TEST INFRASTRUCTURE (see oracle.h).  Restatement of the reference's
// FeatureExtraction node body (lidar_odometry/src/featureExtraction.cpp):
//   calculateSmoothness :87-111, markOccludedPoints :113-148, extractFeatures :150-245
// "next" row 8f-1 of SURVEY.md.  Unspecified behaviour of the reference is pinned as documented
// in oracle.h (sort ties by index; never-initialised entries read as zero).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
struct Smooth { float value; int ind; };
}

extern "C" int orc_extract_features(const float* pts, size_t n_, const float* range, const int32_t* col,
                                    const int32_t* start_ring, const int32_t* end_ring, int n_scan,
                                    float edge_th, float surf_th, float leaf, float* corner_out,
                                    size_t* n_corner, float* surf_out, size_t* n_surf, int32_t* label_out) {
    const int n = (int)n_;
    std::vector<float> curv(n, 0.0f);
    std::vector<int> picked(n, 0), label(n, 0);
    std::vector<Smooth> sm(n);
    for (int i = 0; i < n; ++i) sm[i] = Smooth{0.0f, i};
    // calculateSmoothness
    for (int i = 5; i < n - 5; ++i) {
        float diff = range[i - 2] + range[i - 1] - range[i] * 4 + range[i + 1] + range[i + 2];
        curv[i] = diff * diff;
        sm[i].value = curv[i];
        sm[i].ind = i;
    }
    // markOccludedPoints
    for (int i = 5; i < n - 6; ++i) {
        float depth1 = range[i], depth2 = range[i + 1];
        int cd = std::abs(int(col[i + 1] - col[i]));
        if (cd < 10) {
            if (depth1 - depth2 > 0.3) {
                picked[i - 1] = 1;
                picked[i] = 1;
            } else if (depth2 - depth1 > 0.3) {
                picked[i + 1] = 1;
                picked[i + 2] = 1;
            }
        }
        float diff1 = std::abs(float(range[i - 1] - range[i]));
        float diff2 = std::abs(float(range[i + 1] - range[i]));
        if (diff1 > 0.1 * range[i] && diff2 > 0.1 * range[i]) picked[i] = 1;
    }
    // extractFeatures
    size_t nc = 0, ns = 0;
    std::vector<float> ring_surf, ring_ds;
    for (int i = 0; i < n_scan; ++i) {
        ring_surf.clear();
        for (int j = 0; j < 6; ++j) {
            int sp = (start_ring[i] * (6 - j) + end_ring[i] * j) / 6;
            int ep = (start_ring[i] * (5 - j) + end_ring[i] * (j + 1)) / 6 - 1;
            if (sp >= ep) continue;
            std::stable_sort(sm.begin() + sp, sm.begin() + ep, [](const Smooth& a, const Smooth& b) {
                return a.value < b.value || (a.value == b.value && a.ind < b.ind);
            });
            int largest = 0;
            for (int k = ep; k >= sp; --k) {
                int ind = sm[k].ind;
                if (picked[ind] == 0 && curv[ind] > edge_th) {
                    ++largest;
                    if (largest <= 40) {
                        label[ind] = 1;
                        std::memcpy(corner_out + 4 * nc, pts + 4 * (size_t)ind, 16);
                        ++nc;
                    } else {
                        break;
                    }
                    picked[ind] = 1;
                    for (int l = 1; l <= 5; ++l) {
                        if (ind + l >= n) break;                       // the reference would read past the arrays
                        int cd = std::abs(int(col[ind + l] - col[ind + l - 1]));
                        if (cd > 10) break;
                        picked[ind + l] = 1;
                    }
                    for (int l = -1; l >= -5; --l) {
                        if (ind + l < 0) break;
                        int cd = std::abs(int(col[ind + l] - col[ind + l + 1]));
                        if (cd > 10) break;
                        picked[ind + l] = 1;
                    }
                }
            }
            for (int k = sp; k <= ep; ++k) {
                int ind = sm[k].ind;
                if (picked[ind] == 0 && curv[ind] < surf_th) {
                    label[ind] = -1;
                    picked[ind] = 1;
                    for (int l = 1; l <= 5; ++l) {
                        if (ind + l >= n) break;
                        int cd = std::abs(int(col[ind + l] - col[ind + l - 1]));
                        if (cd > 10) break;
                        picked[ind + l] = 1;
                    }
                    for (int l = -1; l >= -5; --l) {
                        if (ind + l < 0) break;
                        int cd = std::abs(int(col[ind + l] - col[ind + l + 1]));
                        if (cd > 10) break;
                        picked[ind + l] = 1;
                    }
                }
            }
            for (int k = sp; k <= ep; ++k)
                if (label[k] <= 0) ring_surf.insert(ring_surf.end(), pts + 4 * (size_t)k, pts + 4 * (size_t)k + 4);
        }
        const size_t m = ring_surf.size() / 4;
        ring_ds.resize(ring_surf.size());
        int pass = 0;
        size_t mds = orc_voxelgrid(ring_surf.data(), m, leaf, ring_ds.data(), nullptr, nullptr, &pass);
        std::memcpy(surf_out + 4 * ns, ring_ds.data(), mds * 16);
        ns += mds;
    }
    *n_corner = nc;
    *n_surf = ns;
    if (label_out) for (int i = 0; i < n; ++i) label_out[i] = label[i];
    return 0;
}
