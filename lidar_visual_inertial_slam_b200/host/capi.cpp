// capi.cpp -- C exports of the host library (liblvreg_host.so) for ctypes callers (bench.py,
// tests): the synthetic generator and the mapOptimization mirror.  Packed {x,y,z,intensity} rows.
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>

#include "front_end.hpp"
#include "harness.hpp"
#include "wire_formats.hpp"

using namespace lvreg_host;

namespace {
thread_local std::string g_err;
struct GenCtx {
    World world;
    SensorSpec sensor;
    std::vector<float> corner, surf;
};
}  // namespace

extern "C" {

const char* lvh_last_error() { return g_err.c_str(); }

// sensor: 0 MID360-like + indoor world, 1 128-beam + urban world
void* lvh_gen_create(int sensor, uint64_t seed) {
    GenCtx* g = new GenCtx();
    g->sensor = sensor == 1 ? sensor_128beam() : sensor_mid360();
    g->world = make_world(seed, sensor == 1 ? world_urban() : world_indoor());
    return g;
}
void lvh_gen_destroy(void* p) { delete (GenCtx*)p; }

// generates one scan; sizes are returned, data is fetched with lvh_gen_fetch
void lvh_gen_scan(void* p, const float pose[6], uint64_t seed, int threads, size_t* n_corner, size_t* n_surf) {
    GenCtx* g = (GenCtx*)p;
    generate_scan(g->world, g->sensor, pose, seed, g->corner, g->surf, threads);
    *n_corner = g->corner.size() / 4;
    *n_surf = g->surf.size() / 4;
}
void lvh_gen_fetch(void* p, float* corner_out, float* surf_out) {
    GenCtx* g = (GenCtx*)p;
    if (corner_out && !g->corner.empty()) std::memcpy(corner_out, g->corner.data(), g->corner.size() * 4);
    if (surf_out && !g->surf.empty()) std::memcpy(surf_out, g->surf.data(), g->surf.size() * 4);
}

void lvh_truth_pose(int sensor, uint64_t seed, double scan_period, double speed, int k, float pose[6]) {
    SequenceSpec s{sensor, seed, 0, scan_period, speed, 0.f, 0.f};
    truth_pose(s, k, pose);
}
void lvh_guess_pose(uint64_t seed, float guess_trans, float guess_rot, int k, const float truth[6], float guess[6]) {
    SequenceSpec s{0, seed, 0, 0.0, 0.0, guess_trans, guess_rot};
    guess_pose(s, k, truth, guess);
}

// ---- mapOptimization mirror ----
void* lvh_mo_create(const lvreg_params* p, int device) {
    try {
        ParamServer ps;
        if (p) ps.lv = *p;
        return new mapOptimization(ps, device);
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
// the same with capacity hints {map corner, map surf, scan corner, scan surf points, grid cells}: every per-call
// device buffer is sized up front (lvreg_reserve), so that no call of a timed replay allocates
void* lvh_mo_create_reserved(const lvreg_params* p, int device, const size_t reserve[5]) {
    try {
        ParamServer ps;
        if (p) ps.lv = *p;
        if (reserve) {
            ps.reserveMapCorner = reserve[0]; ps.reserveMapSurf = reserve[1]; ps.reserveScanCorner = reserve[2];
            ps.reserveScanSurf = reserve[3]; ps.reserveGridCells = reserve[4];
        }
        return new mapOptimization(ps, device);
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void lvh_mo_destroy(void* mo) { delete (mapOptimization*)mo; }

// laserCloudInfoHandler for one scan (packed rows).  Returns the lvreg status of scan2map, or -1
// when throttled, -2 on exception.
int lvh_mo_handle_scan(void* p, const float* corner, size_t nc, const float* surf, size_t ns, double stamp,
                       const float guess[6], float pose_out[6], lvreg_result* res, lvreg_timings* tim,
                       int* n_keyframes) {
    mapOptimization* mo = (mapOptimization*)p;
    try {
        Cloud c = cloud_from_xyzi(corner, nc), s = cloud_from_xyzi(surf, ns);
        if (!mo->laserCloudInfoHandler(c, s, stamp, guess)) return -1;
        std::memcpy(pose_out, mo->transformTobeMapped, 6 * sizeof(float));
        if (res) *res = mo->lastResult;
        if (tim) *tim = mo->lastTimings;
        if (n_keyframes) *n_keyframes = (int)mo->cloudKeyPoses3D.size();
        return mo->lastStatus;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -2;
    }
}
size_t lvh_mo_selection(void* p, int32_t* ids, size_t cap) {
    mapOptimization* mo = (mapOptimization*)p;
    const std::vector<int32_t>& v = mo->lastKeyframeSelection();
    size_t n = v.size() < cap ? v.size() : cap;
    if (ids && n) std::memcpy(ids, v.data(), n * sizeof(int32_t));
    return v.size();
}
void* lvh_mo_handle(void* p) { return ((mapOptimization*)p)->handle(); }

// performLoopClosure (MO:549-628) on the mirror: 1 = constraint queued, 0 = no candidate / gated, -2 = exception.
// cur / pre receive the candidate pair of detectLoopClosureDistance (-1 when none), out the device result.
int lvh_mo_perform_loop_closure(void* p, int* cur, int* pre, lvreg_loop_result* out) {
    mapOptimization* mo = (mapOptimization*)p;
    try {
        int c = -1, q = -1;
        const bool have = mo->detectLoopClosureDistance(&c, &q);
        if (cur) *cur = have ? c : -1;
        if (pre) *pre = have ? q : -1;
        const bool queued = mo->performLoopClosure();
        if (out) *out = mo->lastLoop;
        return queued ? 1 : 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -2;
    }
}

// Front-end mirrors (ImageProjection + FeatureExtraction) on the handle of a mapOptimization mirror:
// raw Livox-layout points (n x 32 bytes) in, registered pose out, nothing but the raw scan uploaded.
// Returns the lvreg status of the registration, -2 on exception.  n_out = {extracted, corner, surf}.
int lvh_mo_raw_scan_to_pose(void* p, const void* raw_points, size_t n, int n_scan, int horizon, int sensor,
                            float min_range, float max_range, int deskew, double time_scan_cur,
                            const double* imu_time, const double* imu_rot_xyz /* k x 3 */, int k,
                            float edge_threshold, const int32_t* ids, size_t n_ids, float pose[6], lvreg_result* res,
                            size_t n_out[3]) {
    mapOptimization* mo = (mapOptimization*)p;
    try {
        ImageProjection ip(mo->handle());
        ip.N_SCAN = n_scan;
        ip.Horizon_SCAN = horizon;
        ip.sensor = (SensorType)sensor;
        ip.lidarMinRange = min_range;
        ip.lidarMaxRange = max_range;
        ip.timeScanCur = time_scan_cur;
        ip.cloudInfo.imu_available = deskew ? 1 : 0;
        ip.laserCloudIn.assign((const PointXYZIRT*)raw_points, (const PointXYZIRT*)raw_points + n);
        if (deskew) {
            for (int i = 0; i < k && i < queueLength; ++i) {
                ip.imuTime[i] = imu_time[i];
                ip.imuRotX[i] = imu_rot_xyz[3 * i];
                ip.imuRotY[i] = imu_rot_xyz[3 * i + 1];
                ip.imuRotZ[i] = imu_rot_xyz[3 * i + 2];
            }
            ip.imuPointerCur = (k < queueLength ? k : queueLength) - 1;
        }
        ip.projectPointCloud(false);
        FeatureExtraction fe(mo->handle());
        fe.N_SCAN = n_scan;
        fe.edgeThreshold = edge_threshold;
        fe.laserCloudInfoHandlerOnDevice(false);
        lvreg_cloud c, s;
        if (lvreg_get_feature_clouds(mo->handle(), &c, &s) != LVREG_OK) throw std::runtime_error("no feature clouds");
        if (n_out) { n_out[0] = ip.extractedCloudSize(); n_out[1] = fe.numCorner(); n_out[2] = fe.numSurface(); }
        return lvreg_register_scan(mo->handle(), &c, &s, ids, n_ids, pose, res);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -2;
    }
}

// whole-sequence replay (what lvreg_replay runs per sequence)
int lvh_replay(int sensor, uint64_t seed, int n_scans, double period, double speed, float gt, float gr, int device,
               int gen_threads, double* out /*10 doubles*/) {
    try {
        SequenceSpec s{sensor, seed, n_scans, period, speed, gt, gr};
        ReplayStats st = replay_sequence(s, device, gen_threads);
        out[0] = st.scans; out[1] = st.registered; out[2] = st.keyframes; out[3] = st.converged;
        out[4] = (double)st.iterations; out[5] = (double)st.queries; out[6] = st.wall_s; out[7] = st.device_ms;
        out[8] = st.max_pos_err; out[9] = st.max_rot_err;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -2;
    }
}

// ---- wire / disk formats ----
// CloudInfo from flat arrays -> CDR bytes.  Returns the size of the message; writes it when it fits in `cap`.
// scalars_i = {imu_available, odom_available, odom_reset_id}; scalars_f = {imu_roll_init, imu_pitch_init,
// imu_yaw_init, initial_guess_x, y, z, roll, pitch, yaw}; clouds as packed {x,y,z,intensity} rows.
size_t lvh_cloudinfo_serialize(double stamp, const char* frame_id, const int32_t* start_ring, const int32_t* end_ring,
                               size_t n_ring, const int32_t* col, const float* range, size_t n_pts,
                               const int64_t scalars_i[3], const float scalars_f[9], const float* deskewed, size_t nd,
                               const float* corner, size_t nc, const float* surf, size_t ns, uint8_t* out, size_t cap) {
    CloudInfo m;
    m.stamp = stamp;
    m.start_ring_index.assign(start_ring, start_ring + n_ring);
    m.end_ring_index.assign(end_ring, end_ring + n_ring);
    m.point_col_ind.assign(col, col + n_pts);
    m.point_range.assign(range, range + n_pts);
    m.imu_available = scalars_i[0]; m.odom_available = scalars_i[1]; m.odom_reset_id = scalars_i[2];
    m.imu_roll_init = scalars_f[0]; m.imu_pitch_init = scalars_f[1]; m.imu_yaw_init = scalars_f[2];
    m.initial_guess_x = scalars_f[3]; m.initial_guess_y = scalars_f[4]; m.initial_guess_z = scalars_f[5];
    m.initial_guess_roll = scalars_f[6]; m.initial_guess_pitch = scalars_f[7]; m.initial_guess_yaw = scalars_f[8];
    m.cloud_deskewed = cloud_from_xyzi(deskewed, nd);
    m.cloud_corner = cloud_from_xyzi(corner, nc);
    m.cloud_surface = cloud_from_xyzi(surf, ns);
    const std::vector<uint8_t> b = serialize_cloud_info(m, frame_id ? frame_id : "");
    if (out && b.size() <= cap) std::memcpy(out, b.data(), b.size());
    return b.size();
}

// CDR bytes -> CloudInfo -> CDR bytes (must reproduce the input); summary = {stamp, rings, points, imu_available,
// odom_available, odom_reset_id, 9 floats, n_deskewed, n_corner, n_surface, sum of all cloud coordinates}.
// Returns the re-serialised size, 0 on a malformed message (lvh_last_error says why).
size_t lvh_cloudinfo_roundtrip(const uint8_t* in, size_t len, uint8_t* out, size_t cap, double summary[19]) {
    try {
        std::string frame;
        const CloudInfo m = deserialize_cloud_info(in, len, &frame);
        if (summary) {
            summary[0] = m.stamp; summary[1] = (double)m.start_ring_index.size(); summary[2] = (double)m.point_range.size();
            summary[3] = (double)m.imu_available; summary[4] = (double)m.odom_available; summary[5] = (double)m.odom_reset_id;
            const float f[9] = {m.imu_roll_init, m.imu_pitch_init, m.imu_yaw_init, m.initial_guess_x, m.initial_guess_y,
                                m.initial_guess_z, m.initial_guess_roll, m.initial_guess_pitch, m.initial_guess_yaw};
            for (int i = 0; i < 9; ++i) summary[6 + i] = f[i];
            summary[15] = (double)m.cloud_deskewed.size(); summary[16] = (double)m.cloud_corner.size();
            summary[17] = (double)m.cloud_surface.size();
            double sum = 0;
            for (const Cloud* c : {&m.cloud_deskewed, &m.cloud_corner, &m.cloud_surface})
                for (const PointType& p : *c) sum += (double)p.x + (double)p.y + (double)p.z + (double)p.intensity;
            summary[18] = sum;
        }
        const std::vector<uint8_t> b = serialize_cloud_info(m, frame);
        if (out && b.size() <= cap) std::memcpy(out, b.data(), b.size());
        return b.size();
    } catch (const std::exception& e) {
        g_err = e.what();
        return 0;
    }
}

// saveMapService (MO:179-236) on the mirror: trajectory / transformations / CornerMap / SurfMap / GlobalMap .pcd
// into `directory` (must exist).  Returns 1 on success, 0 on an I/O failure, -2 on exception.
int lvh_mo_save_map(void* p, const char* directory, float resolution) {
    try {
        return ((mapOptimization*)p)->saveMap(directory, resolution) ? 1 : 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -2;
    }
}

// binary PCD (x y z intensity ...) -> packed rows; returns the number of points, or -1
long long lvh_load_pcd_xyzi(const char* path, float* out, size_t cap_points) {
    Cloud c;
    if (!load_pcd_binary_xyzi(path, &c)) return -1;
    if (out)
        for (size_t i = 0; i < c.size() && i < cap_points; ++i) {
            out[4 * i] = c[i].x; out[4 * i + 1] = c[i].y; out[4 * i + 2] = c[i].z; out[4 * i + 3] = c[i].intensity;
        }
    return (long long)c.size();
}

// replay of one sequence on an existing mirror (reset() first): what lvreg_replay does per sequence and GPU
int lvh_replay_on(void* mo, int sensor, uint64_t seed, int n_scans, double period, double speed, float gt, float gr,
                  int gen_threads, double* out /*11 doubles*/) {
    try {
        SequenceSpec s{sensor, seed, n_scans, period, speed, gt, gr};
        ReplayStats st = replay_sequence(s, 0, gen_threads, (mapOptimization*)mo);
        out[0] = st.scans; out[1] = st.registered; out[2] = st.keyframes; out[3] = st.converged;
        out[4] = (double)st.iterations; out[5] = (double)st.queries; out[6] = st.wall_s; out[7] = st.device_ms;
        out[8] = st.max_pos_err; out[9] = st.max_rot_err; out[10] = (double)st.launches;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -2;
    }
}

}  // extern "C"
