// synth.cpp -- see synth.hpp.
#include "synth.hpp"

#include <algorithm>
#include <cmath>
#include <thread>

namespace lvreg_host {

namespace {
inline uint64_t splitmix64(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
inline uint64_t hash2(uint64_t a, uint64_t b) {
    uint64_t x = a * 0x9e3779b97f4a7c15ull + b;
    return splitmix64(x);
}
inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

struct Mat3 { double m[9]; };
Mat3 rot_rpy(double roll, double pitch, double yaw) {
    double A = std::cos(yaw), B = std::sin(yaw), C = std::cos(pitch), D = std::sin(pitch);
    double E = std::cos(roll), F = std::sin(roll);
    Mat3 R;
    R.m[0] = A * C; R.m[1] = A * D * F - B * E; R.m[2] = B * F + A * D * E;
    R.m[3] = B * C; R.m[4] = A * E + B * D * F; R.m[5] = B * D * E - A * F;
    R.m[6] = -D;    R.m[7] = C * F;             R.m[8] = C * E;
    return R;
}
inline void mul(const Mat3& R, const double v[3], double o[3]) {
    for (int i = 0; i < 3; ++i) o[i] = R.m[3 * i] * v[0] + R.m[3 * i + 1] * v[1] + R.m[3 * i + 2] * v[2];
}
inline void mulT(const Mat3& R, const double v[3], double o[3]) {
    for (int i = 0; i < 3; ++i) o[i] = R.m[i] * v[0] + R.m[3 + i] * v[1] + R.m[6 + i] * v[2];
}

// nearest intersection of o + t d with any box (slab test); returns t or +inf
inline double hit_boxes(const World& w, const double o[3], const double d[3], double tmax) {
    double best = tmax;
    for (const Box& b : w.boxes) {
        double t0 = 0.0, t1 = best;
        bool ok = true;
        for (int a = 0; a < 3 && ok; ++a) {
            if (std::fabs(d[a]) < 1e-12) {
                if (o[a] < b.lo[a] || o[a] > b.hi[a]) ok = false;
            } else {
                double inv = 1.0 / d[a];
                double ta = (b.lo[a] - o[a]) * inv, tb = (b.hi[a] - o[a]) * inv;
                if (ta > tb) std::swap(ta, tb);
                if (ta > t0) t0 = ta;
                if (tb < t1) t1 = tb;
                if (t0 > t1) ok = false;
            }
        }
        if (ok && t0 > 1e-6 && t0 < best) best = t0;
    }
    return best;
}

struct CornerCand { int ring; float az; float p[4]; };
}  // namespace

Rng::Rng(uint64_t seed) {
    uint64_t x = seed;
    for (int i = 0; i < 4; ++i) s[i] = splitmix64(x);
}
uint64_t Rng::next() {       // xoshiro256**
    const uint64_t result = rotl(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return result;
}
double Rng::uniform() { return u01(next()); }
double Rng::uniform(double lo, double hi) { return lo + (hi - lo) * uniform(); }
double Rng::normal() {
    if (has_spare) { has_spare = false; return spare; }
    double u1 = uniform(), u2 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    double r = std::sqrt(-2.0 * std::log(u1)), th = 6.283185307179586 * u2;
    spare = r * std::sin(th);
    has_spare = true;
    return r * std::cos(th);
}

SensorSpec sensor_128beam() {
    SensorSpec s;
    s.rings = 128; s.cols = 2048;
    s.elev_lo = -22.5f * 0.017453292f; s.elev_hi = 22.5f * 0.017453292f;
    s.max_range = 260.f; s.min_range = 1.0f; s.range_noise = 0.02f; s.corner_cap = 40;
    return s;
}
SensorSpec sensor_mid360() {      // ~20k returns per scan, FOV 360 x [-7, +52] deg
    SensorSpec s;
    s.rings = 32; s.cols = 625;
    s.elev_lo = -7.f * 0.017453292f; s.elev_hi = 52.f * 0.017453292f;
    s.max_range = 40.f; s.min_range = 0.5f; s.range_noise = 0.02f; s.corner_cap = 40;
    return s;
}
WorldSpec world_urban() {
    WorldSpec w;
    w.half_extent = 330.f; w.n_boxes = 320; w.box_min = 4.f; w.box_max = 18.f;
    w.height_min = 4.f; w.height_max = 30.f; w.n_poles = 4800; w.pole_height = 8.f;
    w.street_half_width = 4.f; w.ceiling_z = -1.f; w.n_lanes = 5; w.lane_spacing = 20.f; w.sensor_height = 3.0f;
    w.n_walls = 0; w.wall_len_min = 8.f; w.wall_len_max = 30.f; w.wall_h_min = 1.0f; w.wall_h_max = 3.0f;
    return w;
}
WorldSpec world_indoor() {
    WorldSpec w;
    w.half_extent = 45.f; w.n_boxes = 60; w.box_min = 1.0f; w.box_max = 6.f;
    w.height_min = 1.0f; w.height_max = 5.5f; w.n_poles = 60; w.pole_height = 5.5f;
    w.street_half_width = 2.0f; w.ceiling_z = 6.0f; w.n_lanes = 1; w.lane_spacing = 0.f; w.sensor_height = 1.2f;
    w.n_walls = 0; w.wall_len_min = w.wall_len_max = w.wall_h_min = w.wall_h_max = 0.f;
    return w;
}

World make_world(uint64_t seed, const WorldSpec& spec) {
    World w;
    w.seed = seed;
    w.ground_z = -spec.sensor_height;          // the sensor rides at z = 0
    w.ceiling_z = spec.ceiling_z;
    Rng rng(seed);
    int guard = 0;
    while ((int)w.boxes.size() < spec.n_boxes && guard++ < spec.n_boxes * 200) {
        float sx = (float)rng.uniform(spec.box_min, spec.box_max);
        float sy = (float)rng.uniform(spec.box_min, spec.box_max);
        float cx = (float)rng.uniform(-spec.half_extent, spec.half_extent);
        float cy = (float)rng.uniform(-spec.half_extent, spec.half_extent);
        float hz = (float)rng.uniform(spec.height_min, spec.height_max);
        Box b;
        b.lo[0] = cx - sx / 2; b.hi[0] = cx + sx / 2;
        b.lo[1] = cy - sy / 2; b.hi[1] = cy + sy / 2;
        b.lo[2] = w.ground_z;  b.hi[2] = w.ground_z + hz;
        bool on_street = false;                                 // keep the streets free
        for (int l = 0; l < spec.n_lanes; ++l) {
            const float yc = spec.lane_spacing * (l - 0.5f * (spec.n_lanes - 1));
            if (b.lo[1] < yc + spec.street_half_width && b.hi[1] > yc - spec.street_half_width) on_street = true;
        }
        if (on_street) continue;
        bool overlap = false;
        for (const Box& o : w.boxes)
            if (b.lo[0] < o.hi[0] + 1.f && b.hi[0] > o.lo[0] - 1.f && b.lo[1] < o.hi[1] + 1.f && b.hi[1] > o.lo[1] - 1.f)
                overlap = true;
        if (!overlap) w.boxes.push_back(b);
    }
    // thin low walls: extra planar + edge structure that hides little behind it
    guard = 0;
    int walls = 0;
    while (walls < spec.n_walls && guard++ < spec.n_walls * 200) {
        const float len = (float)rng.uniform(spec.wall_len_min, spec.wall_len_max);
        const float hz = (float)rng.uniform(spec.wall_h_min, spec.wall_h_max);
        const float cx = (float)rng.uniform(-spec.half_extent, spec.half_extent);
        const float cy = (float)rng.uniform(-spec.half_extent, spec.half_extent);
        const bool along_x = rng.uniform() < 0.5;
        Box b;
        const float hx = along_x ? len / 2 : 0.15f, hy = along_x ? 0.15f : len / 2;
        b.lo[0] = cx - hx; b.hi[0] = cx + hx;
        b.lo[1] = cy - hy; b.hi[1] = cy + hy;
        b.lo[2] = w.ground_z; b.hi[2] = w.ground_z + hz;
        bool bad = false;
        for (int l = 0; l < spec.n_lanes; ++l) {
            const float yc = spec.lane_spacing * (l - 0.5f * (spec.n_lanes - 1));
            if (b.lo[1] < yc + spec.street_half_width && b.hi[1] > yc - spec.street_half_width) bad = true;
        }
        for (const Box& o : w.boxes)
            if (b.lo[0] < o.hi[0] + 0.5f && b.hi[0] > o.lo[0] - 0.5f && b.lo[1] < o.hi[1] + 0.5f && b.hi[1] > o.lo[1] - 0.5f)
                bad = true;
        if (bad) continue;
        w.boxes.push_back(b);
        ++walls;
    }
    guard = 0;
    while ((int)w.poles.size() < spec.n_poles && guard++ < spec.n_poles * 200) {
        Pole p;
        p.x = (float)rng.uniform(-spec.half_extent, spec.half_extent);
        p.y = (float)rng.uniform(-spec.half_extent, spec.half_extent);
        bool on_street = false;
        for (int l = 0; l < spec.n_lanes; ++l) {
            const float yc = spec.lane_spacing * (l - 0.5f * (spec.n_lanes - 1));
            if (std::fabs(p.y - yc) < spec.street_half_width * 0.5f) on_street = true;
        }
        if (on_street) continue;
        bool inside = false;
        for (const Box& o : w.boxes)
            if (p.x > o.lo[0] - 0.5f && p.x < o.hi[0] + 0.5f && p.y > o.lo[1] - 0.5f && p.y < o.hi[1] + 0.5f) inside = true;
        if (inside) continue;
        p.z0 = w.ground_z;
        p.z1 = w.ground_z + spec.pole_height;
        w.poles.push_back(p);
    }
    return w;
}

void generate_scan(const World& w, const SensorSpec& s, const float pose[6], uint64_t seed,
                   std::vector<float>& corner, std::vector<float>& surf, int threads) {
    const Mat3 R = rot_rpy(pose[0], pose[1], pose[2]);
    const double o[3] = {pose[3], pose[4], pose[5]};
    if (threads < 1) threads = 1;
    if (threads > s.rings) threads = s.rings;
    const double de = (double)(s.elev_hi - s.elev_lo) / s.rings;

    // ---- surf: ray casting, ring-major output ----
    std::vector<std::vector<float>> part(threads);
    auto cast = [&](int tid) {
        std::vector<float>& out = part[tid];
        const int r0 = (int)((long long)s.rings * tid / threads), r1 = (int)((long long)s.rings * (tid + 1) / threads);
        out.reserve((size_t)(r1 - r0) * s.cols * 4);
        for (int r = r0; r < r1; ++r) {
            const double e = s.elev_lo + de * (r + 0.5);
            const double ce = std::cos(e), se = std::sin(e);
            for (int c = 0; c < s.cols; ++c) {
                const double a = 6.283185307179586 * (c + 0.37 * r) / s.cols;
                const double ds[3] = {ce * std::cos(a), ce * std::sin(a), se};
                double dw[3];
                mul(R, ds, dw);
                double t = s.max_range;
                if (dw[2] < -1e-9) t = std::min(t, (w.ground_z - o[2]) / dw[2]);
                if (w.ceiling_z > 0 && dw[2] > 1e-9) t = std::min(t, (w.ceiling_z - o[2]) / dw[2]);
                t = hit_boxes(w, o, dw, t);
                if (!(t < s.max_range) || t < s.min_range) continue;
                const uint64_t hsh = hash2(seed, (uint64_t)r * 1000003ull + c);
                uint64_t st = hsh;
                double u1 = u01(splitmix64(st)), u2 = u01(splitmix64(st)), u3 = u01(splitmix64(st));
                if (u1 < 1e-300) u1 = 1e-300;
                const double nz = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
                const double rg = t + s.range_noise * nz;
                out.push_back((float)(ds[0] * rg));
                out.push_back((float)(ds[1] * rg));
                out.push_back((float)(ds[2] * rg));
                out.push_back((float)(255.0 * u3));
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < threads; ++t) th.emplace_back(cast, t);
        cast(0);
        for (auto& x : th) x.join();
    }
    surf.clear();
    for (int t = 0; t < threads; ++t) surf.insert(surf.end(), part[t].begin(), part[t].end());

    // ---- corner: edges sampled by the laser rings / columns ----
    std::vector<CornerCand> cand;
    auto to_sensor = [&](const double P[3], double ps[3]) {
        const double d[3] = {P[0] - o[0], P[1] - o[1], P[2] - o[2]};
        mulT(R, d, ps);
    };
    auto elev_of = [&](const double P[3]) {
        double ps[3];
        to_sensor(P, ps);
        return std::atan2(ps[2], std::sqrt(ps[0] * ps[0] + ps[1] * ps[1]));
    };
    auto try_add = [&](const double P[3], uint64_t tag) {
        double ps[3];
        to_sensor(P, ps);
        const double dist = std::sqrt(ps[0] * ps[0] + ps[1] * ps[1] + ps[2] * ps[2]);
        if (dist < s.min_range || dist > s.max_range) return;
        const double e = std::asin(ps[2] / dist);
        if (e < s.elev_lo || e >= s.elev_hi) return;
        double dw[3] = {(P[0] - o[0]) / dist, (P[1] - o[1]) / dist, (P[2] - o[2]) / dist};
        if (hit_boxes(w, o, dw, dist - 0.05) < dist - 0.05) return;       // occluded
        int ring = (int)((e - s.elev_lo) / de);
        if (ring >= s.rings) ring = s.rings - 1;
        uint64_t st = hash2(seed ^ 0xC0FFEEull, tag);
        double u1 = u01(splitmix64(st)), u2 = u01(splitmix64(st)), u3 = u01(splitmix64(st));
        if (u1 < 1e-300) u1 = 1e-300;
        const double nz = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
        const double k = (dist + s.range_noise * nz) / dist;
        CornerCand cc;
        cc.ring = ring;
        cc.az = (float)std::atan2(ps[1], ps[0]);
        cc.p[0] = (float)(ps[0] * k); cc.p[1] = (float)(ps[1] * k); cc.p[2] = (float)(ps[2] * k);
        cc.p[3] = (float)(255.0 * u3);
        cand.push_back(cc);
    };
    // vertical edges: one sample per ring crossing
    auto vertical = [&](double x, double y, double z0, double z1, uint64_t id) {
        double P0[3] = {x, y, z0}, P1[3] = {x, y, z1};
        double e0 = elev_of(P0), e1 = elev_of(P1);
        if (e0 > e1) return;
        for (int r = 0; r < s.rings; ++r) {
            const double e = s.elev_lo + de * (r + 0.5);
            if (e < e0 || e > e1) continue;
            double lo = z0, hi = z1;
            for (int it = 0; it < 30; ++it) {
                double mid = 0.5 * (lo + hi);
                double Pm[3] = {x, y, mid};
                if (elev_of(Pm) < e) lo = mid; else hi = mid;
            }
            double P[3] = {x, y, 0.5 * (lo + hi)};
            try_add(P, id * 4099ull + (uint64_t)r);
        }
    };
    // horizontal edges: one sample per azimuth column crossing
    auto horizontal = [&](const double A[3], const double B[3], uint64_t id) {
        double As[3], Bs[3];
        to_sensor(A, As);
        to_sensor(B, Bs);
        double a0 = std::atan2(As[1], As[0]), a1 = std::atan2(Bs[1], Bs[0]);
        double da = a1 - a0;
        if (da > M_PI) da -= 2 * M_PI;
        if (da < -M_PI) da += 2 * M_PI;
        const double step = 6.283185307179586 / s.cols;
        const int n = (int)(std::fabs(da) / step);
        for (int k = 0; k <= n; ++k) {
            const double a = a0 + (da >= 0 ? 1 : -1) * step * (k + 0.5);
            const double dx = std::cos(a), dy = std::sin(a);
            const double ex = Bs[0] - As[0], ey = Bs[1] - As[1];
            const double den = ex * dy - ey * dx;
            if (std::fabs(den) < 1e-9) continue;
            const double u = -(As[0] * dy - As[1] * dx) / den;
            if (u < 0.0 || u > 1.0) continue;
            double P[3] = {A[0] + u * (B[0] - A[0]), A[1] + u * (B[1] - A[1]), A[2] + u * (B[2] - A[2])};
            try_add(P, id * 8209ull + (uint64_t)k);
        }
    };
    uint64_t id = 1;
    for (const Box& b : w.boxes) {
        const double xs[2] = {b.lo[0], b.hi[0]}, ys[2] = {b.lo[1], b.hi[1]};
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) vertical(xs[i], ys[j], b.lo[2], b.hi[2], id++);
        const double z = b.hi[2];
        const double c00[3] = {xs[0], ys[0], z}, c10[3] = {xs[1], ys[0], z}, c11[3] = {xs[1], ys[1], z}, c01[3] = {xs[0], ys[1], z};
        horizontal(c00, c10, id++);
        horizontal(c10, c11, id++);
        horizontal(c11, c01, id++);
        horizontal(c01, c00, id++);
    }
    for (const Pole& p : w.poles) vertical(p.x, p.y, p.z0, p.z1, id++);

    // cap: at most corner_cap per ring per 1/6 sector (featureExtraction.cpp:158-185)
    std::stable_sort(cand.begin(), cand.end(), [](const CornerCand& a, const CornerCand& b) {
        if (a.ring != b.ring) return a.ring < b.ring;
        return a.az < b.az;
    });
    corner.clear();
    size_t i = 0;
    while (i < cand.size()) {
        const int ring = cand[i].ring;
        const int sector = std::min(5, (int)((cand[i].az + M_PI) / (2 * M_PI) * 6));
        size_t j = i;
        while (j < cand.size() && cand[j].ring == ring &&
               std::min(5, (int)((cand[j].az + M_PI) / (2 * M_PI) * 6)) == sector) ++j;
        const size_t cnt = j - i;
        const size_t keep = std::min(cnt, (size_t)s.corner_cap);
        for (size_t k = 0; k < keep; ++k) {
            const CornerCand& cc = cand[i + k * cnt / keep];
            corner.insert(corner.end(), cc.p, cc.p + 4);
        }
        i = j;
    }
}

}  // namespace lvreg_host
