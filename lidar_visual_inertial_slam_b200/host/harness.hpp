// harness.hpp -- synthetic sequences for the BASELINE configs (C1..C5) on top of synth.hpp and
// the mapOptimization mirror.  Shared by lvreg_replay (replay_main.cpp) and the C exports that
// bench.py / the tests load through ctypes (capi.cpp).
#pragma once

#include <cstdint>
#include <vector>

#include "map_optimization.hpp"
#include "synth.hpp"

namespace lvreg_host {

struct SequenceSpec {
    int sensor;            // 0 = MID360-like (C1/C2), 1 = 128-beam (C3)
    uint64_t seed;         // 0x5EED0000 + sequence id
    int n_scans;
    double scan_period;    // seconds between scans (0.2 s keeps every scan above the 0.15 s throttle)
    double speed;          // m/s along +x
    float guess_trans;     // initial-guess perturbation, metres  (uniform in [-g, g] per axis)
    float guess_rot;       // radians
};

// ground-truth pose of scan k: 1 m/s forward along the street, sinusoidal yaw
void truth_pose(const SequenceSpec& s, int k, float pose[6]);
// truth perturbed by the deterministic initial-guess error of scan k
void guess_pose(const SequenceSpec& s, int k, const float truth[6], float guess[6]);

struct ReplayStats {
    int scans = 0, registered = 0, keyframes = 0, converged = 0;
    long long iterations = 0, queries = 0;      // queries = (Nc + Ns) * iterations
    double wall_s = 0.0, device_ms = 0.0;
    double max_pos_err = 0.0, max_rot_err = 0.0;
    long long launches = 0;
};

// Replays one synthetic sequence through the mirror: scan 0 bootstraps the first keyframe at the
// truth pose, scans 1.. are registered against the local map.  Scans are generated up front
// (outside the timed region).
// `reuse`: an existing mirror object (one per GPU / robot slot) to replay on after reset(), so that a batch of
// sequences does not pay the first-use device allocations once per sequence; nullptr creates a fresh one
ReplayStats replay_sequence(const SequenceSpec& s, int device, int gen_threads, class mapOptimization* reuse = nullptr);

}  // namespace lvreg_host
