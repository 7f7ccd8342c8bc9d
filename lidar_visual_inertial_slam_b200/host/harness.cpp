// harness.cpp -- see harness.hpp.
#include "harness.hpp"

#include <chrono>
#include <memory>
#include <cmath>
#include <cstring>

namespace lvreg_host {

void truth_pose(const SequenceSpec& s, int k, float pose[6]) {
    const double t = k * s.scan_period;
    pose[0] = (float)(0.01 * std::sin(0.7 * t));
    pose[1] = (float)(0.01 * std::cos(0.5 * t));
    pose[2] = (float)(0.15 * std::sin(0.1 * 6.283185307179586 * t / 10.0));
    // forward along the street at `speed`, turning back smoothly before the end of the world
    const double R = s.sensor == 1 ? 200.0 : 30.0;
    pose[3] = (float)(R * std::sin(s.speed * t / R));
    pose[4] = (float)(0.8 * std::sin(0.05 * t * 6.283185307179586));
    pose[5] = (float)(0.02 * std::sin(0.3 * t));
}

void guess_pose(const SequenceSpec& s, int k, const float truth[6], float guess[6]) {
    Rng rng(s.seed * 1315423911ull + 77ull * (uint64_t)k + 5ull);
    for (int i = 0; i < 3; ++i) guess[i] = truth[i] + (float)rng.uniform(-s.guess_rot, s.guess_rot);
    for (int i = 3; i < 6; ++i) guess[i] = truth[i] + (float)rng.uniform(-s.guess_trans, s.guess_trans);
}

ReplayStats replay_sequence(const SequenceSpec& s, int device, int gen_threads, mapOptimization* reuse) {
    ReplayStats st;
    const SensorSpec sensor = s.sensor == 1 ? sensor_128beam() : sensor_mid360();
    const World world = make_world(s.seed, s.sensor == 1 ? world_urban() : world_indoor());
    std::vector<Cloud> corners(s.n_scans), surfs(s.n_scans);
    std::vector<std::vector<float>> truths(s.n_scans, std::vector<float>(6));
    for (int k = 0; k < s.n_scans; ++k) {
        truth_pose(s, k, truths[k].data());
        std::vector<float> c, f;
        generate_scan(world, sensor, truths[k].data(), s.seed + 1000003ull * (uint64_t)k, c, f, gen_threads);
        corners[k] = cloud_from_xyzi(c.data(), c.size() / 4);
        surfs[k] = cloud_from_xyzi(f.data(), f.size() / 4);
    }
    std::unique_ptr<mapOptimization> own;
    if (!reuse) own.reset(new mapOptimization(ParamServer(), device));
    mapOptimization& mo = reuse ? *reuse : *own;
    if (reuse) mo.reset();
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < s.n_scans; ++k) {
        float guess[6];
        if (k == 0) std::memcpy(guess, truths[0].data(), sizeof(guess));
        else guess_pose(s, k, truths[k].data(), guess);
        if (!mo.laserCloudInfoHandler(corners[k], surfs[k], k * s.scan_period, guess)) continue;
        ++st.scans;
        st.launches += mo.lastTimings.kernel_launches;
        if (k > 0 && mo.lastStatus == LVREG_OK) {
            ++st.registered;
            st.converged += mo.lastResult.converged;
            st.iterations += mo.lastResult.iterations;
            st.queries += (long long)(mo.lastResult.n_corner_ds + mo.lastResult.n_surf_ds) * mo.lastResult.iterations;
            st.device_ms += mo.lastTimings.upload_ms + mo.lastTimings.downsample_ms + mo.lastTimings.map_build_ms +
                            mo.lastTimings.grid_build_ms + mo.lastTimings.register_ms;
            for (int i = 0; i < 3; ++i) {
                st.max_rot_err = std::fmax(st.max_rot_err, std::fabs((double)mo.transformTobeMapped[i] - truths[k][i]));
                st.max_pos_err = std::fmax(st.max_pos_err, std::fabs((double)mo.transformTobeMapped[3 + i] - truths[k][3 + i]));
            }
        }
    }
    st.wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    st.keyframes = (int)mo.cloudKeyPoses3D.size();
    return st;
}

}  // namespace lvreg_host
