// oracle_kdtree.cpp -- TEST INFRASTRUCTURE (see oracle.h).
// Exact k-nearest-neighbour search standing in for pcl::KdTreeFLANN<PointXYZI>
// (setInputCloud MO:1322-1323, nearestKSearch MO:1019/1111, radiusSearch MO:903).
// FLANN 1.9 KDTreeSingleIndex (leaf_max_size 15, reorder, L2_Simple) is restated from its
// published algorithm: middle split on the widest dimension, bounding-
// box lower bounds while descending, near child first.  One deliberate difference: FLANN
// keeps the first-visited of equal distances (traversal dependent); here ties are ordered
// by (d2, index) so results are independent of tree shape and comparable with the CUDA grid.
// d2 = ((dx*dx) + dy*dy) + dz*dz in fp32, one rounding per operation (L2_Simple on x86-64).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace {

constexpr int kLeafMax = 15;
constexpr int kMaxK = 8;

struct Node {
    int left, right;       // leaf: [left, right) into the reordered arrays
    int divfeat;           // -1 for leaf
    float divlow, divhigh;
    int child1, child2;
};

struct BBox { float lo[3], hi[3]; };

inline float sqdist3(const float* a, const float* b) {
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    float r = dx * dx;
    r += dy * dy;
    r += dz * dz;
    return r;
}

// fixed-capacity sorted result set ordered by (d2, index)
struct TopK {
    int k, count;
    float d[kMaxK];
    int32_t id[kMaxK];
    explicit TopK(int k_) : k(k_), count(0) {}
    inline float worst() const { return count < k ? std::numeric_limits<float>::infinity() : d[k - 1]; }
    inline void offer(float dist, int32_t index) {
        if (count == k) {
            if (dist > d[k - 1] || (dist == d[k - 1] && index > id[k - 1])) return;
        }
        int pos = (count < k) ? count : k - 1;
        while (pos > 0 && (d[pos - 1] > dist || (d[pos - 1] == dist && id[pos - 1] > index))) {
            d[pos] = d[pos - 1];
            id[pos] = id[pos - 1];
            --pos;
        }
        d[pos] = dist;
        id[pos] = index;
        if (count < k) ++count;
    }
};

}  // namespace

struct orc_kdtree {
    size_t n = 0;
    std::vector<float> pts;      // reordered xyz (3 floats per point)
    std::vector<int32_t> vind;   // reordered position -> original index
    std::vector<Node> nodes;
    BBox root_bbox;
    int root = -1;

    void minmax(const int32_t* ind, int count, int dim, float& mn, float& mx) const {
        mn = mx = src[(size_t)ind[0] * 4 + dim];
        for (int i = 1; i < count; ++i) {
            float v = src[(size_t)ind[i] * 4 + dim];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
    }
    void plane_split(int32_t* ind, int count, int dim, float cut, int& lim1, int& lim2) const {
        int left = 0, right = count - 1;
        for (;;) {
            while (left <= right && src[(size_t)ind[left] * 4 + dim] < cut) ++left;
            while (left <= right && src[(size_t)ind[right] * 4 + dim] >= cut) --right;
            if (left > right) break;
            std::swap(ind[left], ind[right]);
            ++left; --right;
        }
        lim1 = left;
        right = count - 1;
        for (;;) {
            while (left <= right && src[(size_t)ind[left] * 4 + dim] <= cut) ++left;
            while (left <= right && src[(size_t)ind[right] * 4 + dim] > cut) --right;
            if (left > right) break;
            std::swap(ind[left], ind[right]);
            ++left; --right;
        }
        lim2 = left;
    }
    int divide(int left, int right, BBox& bbox) {
        int me = (int)nodes.size();
        nodes.push_back(Node());
        if (right - left <= kLeafMax) {
            nodes[me].divfeat = -1;
            nodes[me].left = left;
            nodes[me].right = right;
            nodes[me].child1 = nodes[me].child2 = -1;
            for (int a = 0; a < 3; ++a) minmax(&tmp_ind[left], right - left, a, bbox.lo[a], bbox.hi[a]);
            return me;
        }
        int32_t* ind = &tmp_ind[left];
        const int count = right - left;
        const float EPS = 0.00001f;
        float max_span = bbox.hi[0] - bbox.lo[0];
        for (int a = 1; a < 3; ++a) max_span = std::max(max_span, bbox.hi[a] - bbox.lo[a]);
        float max_spread = -1.0f;
        int cutfeat = 0;
        for (int a = 0; a < 3; ++a) {
            float span = bbox.hi[a] - bbox.lo[a];
            if (span > (1.0f - EPS) * max_span) {
                float mn, mx;
                minmax(ind, count, a, mn, mx);
                if (mx - mn > max_spread) { cutfeat = a; max_spread = mx - mn; }
            }
        }
        float split = (bbox.lo[cutfeat] + bbox.hi[cutfeat]) / 2;
        float mn, mx;
        minmax(ind, count, cutfeat, mn, mx);
        float cutval = split < mn ? mn : (split > mx ? mx : split);
        int lim1, lim2, idx;
        plane_split(ind, count, cutfeat, cutval, lim1, lim2);
        if (lim1 > count / 2) idx = lim1;
        else if (lim2 < count / 2) idx = lim2;
        else idx = count / 2;

        BBox lb = bbox, rb = bbox;
        lb.hi[cutfeat] = cutval;
        int c1 = divide(left, left + idx, lb);
        rb.lo[cutfeat] = cutval;
        int c2 = divide(left + idx, right, rb);
        Node& nd = nodes[me];
        nd.divfeat = cutfeat;
        nd.child1 = c1;
        nd.child2 = c2;
        nd.divlow = lb.hi[cutfeat];
        nd.divhigh = rb.lo[cutfeat];
        nd.left = nd.right = 0;
        for (int a = 0; a < 3; ++a) {
            bbox.lo[a] = std::min(lb.lo[a], rb.lo[a]);
            bbox.hi[a] = std::max(lb.hi[a], rb.hi[a]);
        }
        return me;
    }

    template <class Visit>
    void search(int ni, const float* q, float mindist, float* dists, Visit& vis) const {
        const Node& nd = nodes[ni];
        if (nd.divfeat < 0) {
            for (int i = nd.left; i < nd.right; ++i) {
                float d = sqdist3(q, &pts[(size_t)i * 3]);
                vis.offer(d, vind[i]);
            }
            return;
        }
        const int f = nd.divfeat;
        const float val = q[f];
        const float diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
        int best, other;
        float cut;
        if (diff1 + diff2 < 0) { best = nd.child1; other = nd.child2; cut = diff2 * diff2; }
        else { best = nd.child2; other = nd.child1; cut = diff1 * diff1; }
        search(best, q, mindist, dists, vis);
        float saved = dists[f];
        float md = mindist + cut - saved;
        dists[f] = cut;
        // 0.9999: the fp32 lower bound may exceed the true bound by a few ulp; never prune on that
        if (md * 0.9999f <= vis.worst()) search(other, q, md, dists, vis);
        dists[f] = saved;
    }
    float root_dists(const float* q, float* dists) const {
        float s = 0.0f;
        for (int a = 0; a < 3; ++a) {
            dists[a] = 0.0f;
            if (q[a] < root_bbox.lo[a]) { float t = q[a] - root_bbox.lo[a]; dists[a] = t * t; }
            if (q[a] > root_bbox.hi[a]) { float t = q[a] - root_bbox.hi[a]; dists[a] = t * t; }
            s += dists[a];
        }
        return s;
    }

    const float* src = nullptr;       // only valid during build
    std::vector<int32_t> tmp_ind;
};

extern "C" orc_kdtree* orc_kdtree_build(const float* map, size_t m) {
    orc_kdtree* t = new orc_kdtree();
    t->n = m;
    if (m == 0) return t;
    t->src = map;
    t->tmp_ind.resize(m);
    for (size_t i = 0; i < m; ++i) t->tmp_ind[i] = (int32_t)i;
    for (int a = 0; a < 3; ++a) t->minmax(t->tmp_ind.data(), (int)m, a, t->root_bbox.lo[a], t->root_bbox.hi[a]);
    t->nodes.reserve(m / 4 + 16);
    BBox bb = t->root_bbox;
    t->root = t->divide(0, (int)m, bb);
    t->root_bbox = bb;
    t->vind.swap(t->tmp_ind);
    t->pts.resize(m * 3);
    for (size_t i = 0; i < m; ++i) {
        const float* p = map + (size_t)t->vind[i] * 4;
        t->pts[3 * i] = p[0]; t->pts[3 * i + 1] = p[1]; t->pts[3 * i + 2] = p[2];
    }
    t->src = nullptr;
    return t;
}

extern "C" void orc_kdtree_free(orc_kdtree* t) { delete t; }

// single query, no OpenMP (called from inside the callers' parallel loops)
void orc_kdtree_knn_one(const orc_kdtree* t, const float* q, int k, int32_t* idx, float* d2) {
    if (k > kMaxK) k = kMaxK;
    TopK top(k);
    if (t->n > 0) {
        float dists[3];
        float md = t->root_dists(q, dists);
        t->search(t->root, q, md, dists, top);
    }
    for (int j = 0; j < k; ++j) {
        idx[j] = j < top.count ? top.id[j] : -1;
        d2[j] = j < top.count ? top.d[j] : std::numeric_limits<float>::infinity();
    }
}

extern "C" void orc_kdtree_knn(const orc_kdtree* t, const float* queries, size_t nq, int k,
                               int32_t* idx, float* d2, int num_threads) {
    if (k > kMaxK) k = kMaxK;
    if (num_threads < 1) num_threads = 1;
    const long long cnt = (long long)nq;
#pragma omp parallel for num_threads(num_threads) schedule(static)
    for (long long i = 0; i < cnt; ++i)
        orc_kdtree_knn_one(t, queries + 4 * i, k, idx + i * k, d2 + i * k);
}

namespace {
struct RadiusSet {
    float r2;
    std::vector<std::pair<float, int32_t>> hits;
    inline float worst() const { return r2; }
    inline void offer(float d, int32_t i) { if (d < r2) hits.emplace_back(d, i); }
};
}  // namespace

extern "C" size_t orc_kdtree_radius(const orc_kdtree* t, const float q[3], float radius,
                                    int32_t* idx, float* d2, size_t cap) {
    RadiusSet rs;
    rs.r2 = radius * radius;
    if (t->n > 0) {
        float dists[3];
        float md = t->root_dists(q, dists);
        t->search(t->root, q, md, dists, rs);
    }
    std::sort(rs.hits.begin(), rs.hits.end());
    size_t n = std::min(cap, rs.hits.size());
    for (size_t i = 0; i < n; ++i) { d2[i] = rs.hits[i].first; idx[i] = rs.hits[i].second; }
    return rs.hits.size();
}

extern "C" void orc_knn5_brute(const float* map, size_t m, const float* queries, size_t nq,
                               int32_t* idx, float* d2, int num_threads) {
    if (num_threads < 1) num_threads = 1;
    const long long cnt = (long long)nq;
#pragma omp parallel for num_threads(num_threads) schedule(static)
    for (long long i = 0; i < cnt; ++i) {
        TopK top(5);
        const float* q = queries + 4 * i;
        for (size_t j = 0; j < m; ++j) top.offer(sqdist3(q, map + 4 * j), (int32_t)j);
        for (int j = 0; j < 5; ++j) {
            idx[i * 5 + j] = j < top.count ? top.id[j] : -1;
            d2[i * 5 + j] = j < top.count ? top.d[j] : std::numeric_limits<float>::infinity();
        }
    }
}
