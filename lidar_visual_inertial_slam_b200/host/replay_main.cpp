// lvreg_replay -- standalone replay harness (BASELINE configs C2/C3/C5): synthetic sequences are
// generated, then replayed through the mapOptimization mirror on one or more GPUs (one host
// thread per GPU, sequences partitioned round-robin, no NCCL, results gathered on the host).
//   lvreg_replay --sensor 0|1 --scans N --sequences S --gpus G [--seed X]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "harness.hpp"
#include "map_optimization.hpp"

using namespace lvreg_host;

int main(int argc, char** argv) {
    int sensor = 0, scans = 50, sequences = 1, gpus = 1;
    uint64_t seed = 0x5EED0000ull;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "--sensor")) sensor = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--scans")) scans = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--sequences")) sequences = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--gpus")) gpus = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--seed")) seed = std::strtoull(argv[i + 1], nullptr, 0);
    }
    std::vector<ReplayStats> stats(sequences);
    std::vector<int> failed(gpus, 0);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; ++g)
        th.emplace_back([&, g]() {
            // one long-lived mapOptimization per GPU, as on a robot: sequences are replayed on it one after the
            // other (reset() in between), so only the first one pays the device allocations
            std::unique_ptr<mapOptimization> mo;
            try {       // untimed warm-up: CUDA context, lazy kernel loading, first allocations
                ParamServer ps;                       // capacity hints: no device allocation inside the timed replay
                if (sensor == 1) { ps.reserveMapCorner = 3000000; ps.reserveMapSurf = 12000000; ps.reserveScanCorner = 65536;
                                   ps.reserveScanSurf = 400000; ps.reserveGridCells = (size_t)1 << 25; }
                else { ps.reserveMapCorner = 200000; ps.reserveMapSurf = 800000; ps.reserveScanCorner = 20000;
                       ps.reserveScanSurf = 60000; ps.reserveGridCells = (size_t)1 << 22; }
                mo.reset(new mapOptimization(ps, g));
                SequenceSpec w{sensor, seed + 7777ull, 8, 0.2, 1.0, 0.10f, 0.035f};
                replay_sequence(w, g, 2, mo.get());
            } catch (const std::exception& e) {
                std::fprintf(stderr, "warm-up on gpu %d failed: %s\n", g, e.what());
            }
            for (int q = g; q < sequences; q += gpus) {
                SequenceSpec s{sensor, seed + (uint64_t)q, scans, 0.2, 1.0, 0.10f, 0.035f};
                try {
                    stats[q] = replay_sequence(s, g, 8 / gpus > 0 ? 8 / gpus : 1, mo.get());
                } catch (const std::exception& e) {
                    std::fprintf(stderr, "sequence %d on gpu %d failed: %s\n", q, g, e.what());
                    failed[g] = 1;
                }
            }
        });
    for (auto& t : th) t.join();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long long reg = 0, conv = 0, iters = 0, queries = 0, launches = 0;
    double timed = 0, dev_ms = 0, perr = 0, rerr = 0;
    int kf = 0;
    for (const ReplayStats& s : stats) {
        reg += s.registered; conv += s.converged; iters += s.iterations; queries += s.queries; launches += s.launches;
        timed = timed > s.wall_s ? timed : s.wall_s;
        dev_ms += s.device_ms; kf += s.keyframes;
        perr = perr > s.max_pos_err ? perr : s.max_pos_err;
        rerr = rerr > s.max_rot_err ? rerr : s.max_rot_err;
    }
    double replay_s = 0;
    for (const ReplayStats& s : stats) replay_s += s.wall_s;
    int bad = 0;
    for (int f : failed) bad += f;
    std::printf("{\"harness\": \"lvreg_replay\", \"sensor\": %d, \"sequences\": %d, \"scans_per_sequence\": %d, \"gpus\": %d, "
                "\"registrations\": %lld, \"converged\": %lld, \"keyframes\": %d, \"iterations\": %lld, \"knn_queries\": %lld, "
                "\"replay_seconds_sum\": %.6f, \"registrations_per_s\": %.3f, \"device_ms_sum\": %.3f, "
                "\"max_pos_err_vs_truth_m\": %.5f, \"max_rot_err_vs_truth_rad\": %.6f, \"kernel_launches\": %lld, "
                "\"wall_seconds_incl_generation\": %.3f, \"failed\": %d}\n",
                sensor, sequences, scans, gpus, reg, conv, kf, iters, queries, replay_s,
                replay_s > 0 ? reg / (replay_s / gpus) : 0.0, dev_ms, perr, rerr, launches, wall, bad);
    return bad ? 1 : 0;
}
