// synth.hpp -- deterministic synthetic LiDAR worlds and scans for the replay harness and bench.py
// (in-repo PRNG, splitmix64 -> xoshiro256**, Box-Muller; no std::*_distribution).
// The reference ships no data; feature extraction (featureExtraction.cpp) is out of scope, so
// the generator emits the two feature clouds the hot path consumes directly:
//   surf   = ray-cast hits on planar structure (ground, building faces)
//   corner = laser-ring samples of edge structure (building edges, roof lines, poles), capped
//            at 40 per ring per 1/6 sector like featureExtraction.cpp:158-185
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

namespace lvreg_host {

struct Rng {
    uint64_t s[4];
    explicit Rng(uint64_t seed);
    uint64_t next();
    double uniform();                       // [0,1)
    double uniform(double lo, double hi);
    double normal();                        // N(0,1), Box-Muller
  private:
    bool has_spare = false;
    double spare = 0.0;
};

struct Box { float lo[3], hi[3]; };
struct Pole { float x, y, z0, z1; };

struct WorldSpec {
    float half_extent;      // buildings are placed in [-half_extent, half_extent]^2
    int   n_boxes;
    float box_min, box_max; // footprint edge range (m)
    float height_min, height_max;
    int   n_poles;
    float pole_height;
    float street_half_width;   // keep |y| < this free of buildings along the x axis (the trajectory)
    float ceiling_z;           // > 0: add a ceiling plane (indoor), <= 0: none
    int   n_lanes;             // parallel free streets at y = lane_spacing * (i - (n_lanes-1)/2)
    float lane_spacing;
    float sensor_height;       // height of the sensor above the ground plane
    int   n_walls;             // thin low walls / fences (0.3 m thick), axis aligned
    float wall_len_min, wall_len_max, wall_h_min, wall_h_max;
};

struct SensorSpec {
    int   rings;            // N_SCAN
    int   cols;             // Horizon_SCAN
    float elev_lo, elev_hi; // radians
    float max_range;
    float min_range;
    float range_noise;      // sigma, metres
    int   corner_cap;       // per ring per 1/6 sector (40 in the reference)
};

struct World {
    std::vector<Box> boxes;
    std::vector<Pole> poles;
    float ground_z = 0.f;
    float ceiling_z = -1.f;
    uint64_t seed = 0;
};

World make_world(uint64_t seed, const WorldSpec& spec);

// 128-beam spinning sensor (BASELINE C3) and a MID360-like pattern (C1/C2)
SensorSpec sensor_128beam();
SensorSpec sensor_mid360();
WorldSpec  world_urban();
WorldSpec  world_indoor();

// One scan from `pose` = {roll,pitch,yaw,x,y,z}; points are returned in the SENSOR frame as
// packed rows {x,y,z,intensity}.  Deterministic in (world, spec, pose, seed); threads only split
// the work, the output order is fixed.
void generate_scan(const World& w, const SensorSpec& s, const float pose[6], uint64_t seed,
                   std::vector<float>& corner_xyzi, std::vector<float>& surf_xyzi, int threads = 8);

}  // namespace lvreg_host
