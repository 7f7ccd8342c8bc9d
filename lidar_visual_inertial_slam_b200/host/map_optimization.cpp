// map_optimization.cpp -- see map_optimization.hpp.  Host logic only (keyframe selection,
// bookkeeping); all point-level work is done by liblvreg on the GPU.
#include "map_optimization.hpp"
#include "wire_formats.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

namespace lvreg_host {

PointType make_point(float x, float y, float z, float intensity) {
    PointType p;
    p.x = x; p.y = y; p.z = z; p.data3 = 1.0f;
    p.intensity = intensity; p.pad0 = p.pad1 = p.pad2 = 0.0f;
    return p;
}

Cloud cloud_from_xyzi(const float* xyzi, size_t n) {
    Cloud c(n);
    for (size_t i = 0; i < n; ++i) c[i] = make_point(xyzi[4 * i], xyzi[4 * i + 1], xyzi[4 * i + 2], xyzi[4 * i + 3]);
    return c;
}

lvreg_cloud as_lvreg_cloud(const Cloud& c) {
    lvreg_cloud lc;
    lc.data = c.empty() ? nullptr : c.data();
    lc.n = c.size();
    lc.stride = sizeof(PointType);
    lc.intensity_offset = offsetof(PointType, intensity);
    lc.on_device = 0;
    lc.reserved = 0;
    return lc;
}

static lvreg_cloud_out as_lvreg_out(Cloud& c) {
    lvreg_cloud_out lo;
    lo.data = c.empty() ? nullptr : c.data();
    lo.capacity = c.size();
    lo.stride = sizeof(PointType);
    lo.intensity_offset = offsetof(PointType, intensity);
    lo.on_device = 0;
    lo.reserved = 0;
    return lo;
}

mapOptimization::mapOptimization(const ParamServer& params, int device) : P_(params) {
    std::memset(&lastResult, 0, sizeof(lastResult));
    std::memset(&lastTimings, 0, sizeof(lastTimings));
    int st = lvreg_create(&P_.lv, device, nullptr, &h_);
    if (st != LVREG_OK)
        throw std::runtime_error(std::string("lvreg_create failed: ") + lvreg_status_string(st) +
                                 " (a CUDA device is required; there is no CPU fallback)");
    if (P_.reserveMapCorner || P_.reserveMapSurf || P_.reserveScanCorner || P_.reserveScanSurf) {
        st = lvreg_reserve(h_, P_.reserveMapCorner, P_.reserveMapSurf, P_.reserveScanCorner, P_.reserveScanSurf,
                           P_.reserveGridCells);
        if (st != LVREG_OK) throw std::runtime_error(std::string("lvreg_reserve: ") + lvreg_last_error(h_));
    }
}

mapOptimization::~mapOptimization() { lvreg_destroy(h_); }

void mapOptimization::reset() {
    cornerCloudKeyFrames.clear();
    surfCloudKeyFrames.clear();
    cloudKeyPoses3D.clear();
    cloudKeyPoses6D.clear();
    loopIndexContainer.clear();
    loopQueue.clear();
    lastIds_.clear();
    mapDirty_ = true;
    timeLastProcessing_ = -1;
    scanDownsampled_ = false;
    isDegenerate = false;
    std::memset(transformTobeMapped, 0, sizeof(transformTobeMapped));
    lvreg_clear_keyframes(h_);
    lvreg_reset_lm_state(h_);
}

// pointDistance (utility.h:409-412)
static inline float pointDistance(const PointType& a, const PointType& b) {
    return std::sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z));
}

// pcl::VoxelGrid over a few hundred KEY POSES (downSizeFilterSurroundingKeyPoses MO:910-911,
// downSizeFilterGlobalMapKeyPoses MO:484-486): keyframe-selection logic, host-side by design.  Same
// algorithm as the device filter: PCL's fp32 bounds and idx arithmetic, a stable order inside a voxel,
// one output per voxel in ascending idx with sequential fp32 sums and a true division; "leaf too small" returns
// the input.  A device call here would cost a launch sequence and a host synchronisation per incoming scan.
static void voxelgrid_key_poses(const Cloud& in, float leaf, Cloud* out) {
    out->clear();
    if (in.empty()) return;
    float mn[3] = {in[0].x, in[0].y, in[0].z}, mx[3] = {in[0].x, in[0].y, in[0].z};
    for (const PointType& p : in) {
        mn[0] = std::fmin(mn[0], p.x); mn[1] = std::fmin(mn[1], p.y); mn[2] = std::fmin(mn[2], p.z);
        mx[0] = std::fmax(mx[0], p.x); mx[1] = std::fmax(mx[1], p.y); mx[2] = std::fmax(mx[2], p.z);
    }
    const float inv = 1.0f / leaf;
    int64_t d[3];
    for (int a = 0; a < 3; ++a) d[a] = (int64_t)((mx[a] - mn[a]) * inv) + 1;
    if (d[0] * d[1] * d[2] > (int64_t)INT32_MAX) { *out = in; return; }
    int min_b[3], div_b[3];
    for (int a = 0; a < 3; ++a) {
        min_b[a] = (int)std::floor(mn[a] * inv);
        div_b[a] = (int)std::floor(mx[a] * inv) - min_b[a] + 1;
    }
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    std::vector<std::pair<uint32_t, uint32_t>> keyed(in.size());
    for (size_t i = 0; i < in.size(); ++i) {
        const int ix = (int)(std::floor(in[i].x * inv) - (float)min_b[0]);
        const int iy = (int)(std::floor(in[i].y * inv) - (float)min_b[1]);
        const int iz = (int)(std::floor(in[i].z * inv) - (float)min_b[2]);
        keyed[i] = {(uint32_t)(ix * mul[0] + iy * mul[1] + iz * mul[2]), (uint32_t)i};
    }
    std::sort(keyed.begin(), keyed.end());               // (idx, input index): the stable order
    for (size_t b = 0; b < keyed.size();) {
        size_t e = b;
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        for (; e < keyed.size() && keyed[e].first == keyed[b].first; ++e) {
            const PointType& p = in[keyed[e].second];
            sx += p.x; sy += p.y; sz += p.z; si += p.intensity;
        }
        const float c = (float)(e - b);
        out->push_back(make_point(sx / c, sy / c, sz / c, si / c));
        b = e;
    }
}

std::vector<int32_t> mapOptimization::extractNearby() {
    std::vector<int32_t> ids;
    const size_t K = cloudKeyPoses3D.size();
    if (K == 0) return ids;
    const PointType& last = cloudKeyPoses3D.back();
    // kdtreeSurroundingKeyPoses->radiusSearch(back, 50 m): strict d2 < r2, sorted by (d2, index)  MO:902-903
    const float r2 = P_.surroundingKeyframeSearchRadius * P_.surroundingKeyframeSearchRadius;
    std::vector<std::pair<float, int>> hits;
    for (size_t i = 0; i < K; ++i) {
        const PointType& p = cloudKeyPoses3D[i];
        float dx = last.x - p.x, dy = last.y - p.y, dz = last.z - p.z;
        float d2 = dx * dx;
        d2 += dy * dy;
        d2 += dz * dz;
        if (d2 < r2) hits.emplace_back(d2, (int)i);
    }
    std::sort(hits.begin(), hits.end());
    Cloud surroundingKeyPoses(hits.size());
    for (size_t i = 0; i < hits.size(); ++i) surroundingKeyPoses[i] = cloudKeyPoses3D[hits[i].second];
    // downSizeFilterSurroundingKeyPoses (2.0 m VoxelGrid) MO:910-911
    Cloud surroundingKeyPosesDS;
    voxelgrid_key_poses(surroundingKeyPoses, P_.surroundingKeyframeDensity, &surroundingKeyPosesDS);
    // 1-NN back to a real key pose to recover the integer id (MO:912-916)
    for (PointType& pt : surroundingKeyPosesDS) {
        float best = INFINITY;
        int besti = 0;
        for (size_t i = 0; i < K; ++i) {
            const PointType& p = cloudKeyPoses3D[i];
            float dx = pt.x - p.x, dy = pt.y - p.y, dz = pt.z - p.z;
            float d2 = dx * dx;
            d2 += dy * dy;
            d2 += dz * dz;
            if (d2 < best) { best = d2; besti = (int)i; }
        }
        pt.intensity = cloudKeyPoses3D[besti].intensity;
    }
    // keyframes of the last 10 s, newest first (MO:919-926); may duplicate ids
    for (long long i = (long long)K - 1; i >= 0; --i) {
        if (timeLaserInfoCur - cloudKeyPoses6D[i].time < 10.0) surroundingKeyPosesDS.push_back(cloudKeyPoses3D[i]);
        else break;
    }
    // extractCloud's filter (MO:938-939)
    for (const PointType& pt : surroundingKeyPosesDS) {
        if (pointDistance(pt, last) > P_.surroundingKeyframeSearchRadius) continue;
        ids.push_back((int32_t)pt.intensity);
    }
    return ids;
}

// detectLoopClosureDistance, MO:630-661: nearest-first radius search around the newest key pose, first
// hit whose time differs by more than historyKeyframeSearchTimeDiff
bool mapOptimization::detectLoopClosureDistance(int* latestID, int* closestID) {
    const int loopKeyCur = (int)cloudKeyPoses3D.size() - 1;
    int loopKeyPre = -1;
    if (loopKeyCur < 0) return false;
    if (loopIndexContainer.find(loopKeyCur) != loopIndexContainer.end()) return false;   // MO:636-638
    const PointType& last = cloudKeyPoses3D.back();
    const float r2 = P_.historyKeyframeSearchRadius * P_.historyKeyframeSearchRadius;
    std::vector<std::pair<float, int>> hits;
    for (size_t i = 0; i < cloudKeyPoses3D.size(); ++i) {
        const PointType& p = cloudKeyPoses3D[i];
        float dx = last.x - p.x, dy = last.y - p.y, dz = last.z - p.z;
        float d2 = dx * dx;
        d2 += dy * dy;
        d2 += dz * dz;
        if (d2 < r2) hits.emplace_back(d2, (int)i);
    }
    std::sort(hits.begin(), hits.end());
    for (const auto& hit : hits) {
        if (std::fabs(cloudKeyPoses6D[hit.second].time - timeLaserInfoCur) > P_.historyKeyframeSearchTimeDiff) {
            loopKeyPre = hit.second;
            break;
        }
    }
    if (loopKeyPre == -1 || loopKeyCur == loopKeyPre) return false;
    *latestID = loopKeyCur;
    *closestID = loopKeyPre;
    return true;
}

// performLoopClosure, MO:549-628.  Submaps, ICP, gates and the pose correction run on the device
// (lvreg_perform_loop_closure); the constraint is queued exactly like loopIndexQueue / loopPoseQueue /
// loopNoiseQueue.
bool mapOptimization::performLoopClosure() {
    if (cloudKeyPoses3D.empty()) return false;
    int loopKeyCur, loopKeyPre;
    if (!detectLoopClosureDistance(&loopKeyCur, &loopKeyPre)) return false;
    lvreg_icp_params icp;
    lvreg_icp_default_params(&icp);
    icp.max_corr_dist = P_.historyKeyframeSearchRadius * 2;          // MO:580
    int st = lvreg_perform_loop_closure(h_, loopKeyCur, loopKeyPre, P_.historyKeyframeSearchNum, &icp,
                                        P_.historyKeyframeFitnessScore, &lastLoop);
    if (st != LVREG_OK) throw std::runtime_error(std::string("lvreg_perform_loop_closure: ") + lvreg_last_error(h_));
    if (lastLoop.status != LVREG_LOOP_OK) return false;
    LoopConstraint c;
    c.loopKeyCur = loopKeyCur;
    c.loopKeyPre = loopKeyPre;
    std::memcpy(c.poseFrom, lastLoop.pose_from, sizeof(c.poseFrom));
    std::memcpy(c.poseTo, lastLoop.pose_to, sizeof(c.poseTo));
    c.noiseScore = lastLoop.noise;
    loopQueue.push_back(c);
    loopIndexContainer[loopKeyCur] = loopKeyPre;                     // MO:626
    return true;
}

// publishGlobalMap, MO:460-510: key poses within the visualisation radius, thinned by a pose-density
// VoxelGrid, their corner + surf clouds under the stored poses, one VoxelGrid -- all point work on the device
bool mapOptimization::saveMap(const std::string& directory, float resolution) {
    std::vector<int32_t> ids(cloudKeyPoses3D.size());
    for (size_t i = 0; i < ids.size(); ++i) ids[i] = (int32_t)i;
    auto device_map = [&](int which, float leaf) {                 // all keyframes, list order, under their stored poses
        Cloud out;
        if (ids.empty()) return out;
        size_t n = 0;
        if (lvreg_build_global_map(h_, ids.data(), ids.size(), which, leaf, &n) != LVREG_OK)
            throw std::runtime_error(std::string("lvreg_build_global_map: ") + lvreg_last_error(h_));
        out.resize(n);
        lvreg_cloud_out o = as_lvreg_out(out);
        if (n && lvreg_icp_get_cloud(h_, 0, &o, &n) != LVREG_OK)
            throw std::runtime_error(std::string("lvreg_icp_get_cloud: ") + lvreg_last_error(h_));
        return out;
    };
    bool ok = save_pcd_binary(directory + "/trajectory.pcd", cloudKeyPoses3D);            // MO:192
    ok = save_pcd_binary(directory + "/transformations.pcd", cloudKeyPoses6D) && ok;     // MO:193
    const Cloud globalCornerCloud = device_map(1, 0.f), globalSurfCloud = device_map(2, 0.f);      // MO:200-205
    if (resolution != 0.f) {                                                              // MO:206-218
        ok = save_pcd_binary(directory + "/CornerMap.pcd", device_map(1, resolution)) && ok;
        ok = save_pcd_binary(directory + "/SurfMap.pcd", device_map(2, resolution)) && ok;
    } else {                                                                              // MO:219-225
        ok = save_pcd_binary(directory + "/CornerMap.pcd", globalCornerCloud) && ok;
        ok = save_pcd_binary(directory + "/SurfMap.pcd", globalSurfCloud) && ok;
    }
    Cloud globalMapCloud = globalCornerCloud;                                             // MO:227-229
    globalMapCloud.insert(globalMapCloud.end(), globalSurfCloud.begin(), globalSurfCloud.end());
    return save_pcd_binary(directory + "/GlobalMap.pcd", globalMapCloud) && ok;
}

Cloud mapOptimization::publishGlobalMap() {
    Cloud out;
    const size_t K = cloudKeyPoses3D.size();
    if (K == 0) return out;
    const PointType& last = cloudKeyPoses3D.back();
    const float r2 = P_.globalMapVisualizationSearchRadius * P_.globalMapVisualizationSearchRadius;
    std::vector<std::pair<float, int>> hits;
    for (size_t i = 0; i < K; ++i) {
        const PointType& p = cloudKeyPoses3D[i];
        float dx = last.x - p.x, dy = last.y - p.y, dz = last.z - p.z;
        float d2 = dx * dx;
        d2 += dy * dy;
        d2 += dz * dz;
        if (d2 < r2) hits.emplace_back(d2, (int)i);
    }
    std::sort(hits.begin(), hits.end());
    Cloud globalMapKeyPoses(hits.size());
    for (size_t i = 0; i < hits.size(); ++i) globalMapKeyPoses[i] = cloudKeyPoses3D[hits[i].second];
    Cloud globalMapKeyPosesDS(globalMapKeyPoses.size());
    size_t nds = 0;
    {
        lvreg_cloud in = as_lvreg_cloud(globalMapKeyPoses);
        lvreg_cloud_out o = as_lvreg_out(globalMapKeyPosesDS);
        int pt = 0;
        int st = lvreg_voxelgrid(h_, &in, P_.globalMapVisualizationPoseDensity, &o, &nds, nullptr, &pt);
        if (st != LVREG_OK) throw std::runtime_error(std::string("lvreg_voxelgrid: ") + lvreg_last_error(h_));
    }
    globalMapKeyPosesDS.resize(nds);
    std::vector<int32_t> ids;
    for (PointType& pt : globalMapKeyPosesDS) {               // 1-NN back to a real key pose, MO:487-491
        float best = INFINITY;
        int besti = 0;
        for (size_t i = 0; i < K; ++i) {
            const PointType& p = cloudKeyPoses3D[i];
            float dx = pt.x - p.x, dy = pt.y - p.y, dz = pt.z - p.z;
            float d2 = dx * dx;
            d2 += dy * dy;
            d2 += dz * dz;
            if (d2 < best) { best = d2; besti = (int)i; }
        }
        pt.intensity = cloudKeyPoses3D[besti].intensity;
        if (pointDistance(pt, last) > P_.globalMapVisualizationSearchRadius) continue;   // MO:496-497
        ids.push_back((int32_t)pt.intensity);
    }
    size_t n = 0;
    int st = lvreg_build_global_map(h_, ids.data(), ids.size(), 3, P_.globalMapVisualizationLeafSize, &n);
    if (st != LVREG_OK) throw std::runtime_error(std::string("lvreg_build_global_map: ") + lvreg_last_error(h_));
    out.resize(n);
    lvreg_cloud_out o = as_lvreg_out(out);
    st = lvreg_icp_get_cloud(h_, 0, &o, &n);
    if (st != LVREG_OK) throw std::runtime_error(std::string("lvreg_icp_get_cloud: ") + lvreg_last_error(h_));
    return out;
}

void mapOptimization::extractSurroundingKeyFrames() {
    if (cloudKeyPoses3D.empty()) return;                     // MO:974-975
    std::vector<int32_t> ids = extractNearby();
    // The reference rebuilds the map on every scan (MO:318); the map is a pure function of the id
    // list and the keyframe poses, so it is rebuilt only when one of them changed.
    if (!mapDirty_ && ids == lastIds_) return;
    lvreg_map_info info;
    int st = lvreg_build_local_map(h_, ids.data(), ids.size(), &info);
    if (st != LVREG_OK) throw std::runtime_error(std::string("lvreg_build_local_map: ") + lvreg_last_error(h_));
    laserCloudCornerFromMapDSNum = (int)info.n_corner_ds;
    laserCloudSurfFromMapDSNum = (int)info.n_surf_ds;
    lastIds_ = ids;
    mapDirty_ = false;
    lvreg_timings t;
    lvreg_get_timings(h_, &t);
    lastTimings.map_build_ms = t.map_build_ms;
    lastTimings.grid_build_ms = t.grid_build_ms;
    lastTimings.kernel_launches += t.kernel_launches;
}

void mapOptimization::downsampleCurrentScan() {
    lvreg_cloud c = as_lvreg_cloud(laserCloudCornerLast), s = as_lvreg_cloud(laserCloudSurfLast);
    size_t nc = 0, ns = 0;
    int st = lvreg_downsample_scan(h_, &c, &s, &nc, &ns);
    if (st != LVREG_OK) throw std::runtime_error(std::string("lvreg_downsample_scan: ") + lvreg_last_error(h_));
    laserCloudCornerLastDSNum = (int)nc;
    laserCloudSurfLastDSNum = (int)ns;
    scanDownsampled_ = true;
    lvreg_timings t;
    lvreg_get_timings(h_, &t);
    lastTimings.upload_ms = t.upload_ms;
    lastTimings.downsample_ms = t.downsample_ms;
    lastTimings.kernel_launches += t.kernel_launches;
}

void mapOptimization::scan2MapOptimization() {
    if (cloudKeyPoses3D.empty()) {                           // MO:1317-1318
        lastStatus = LVREG_ERR_NO_KEYFRAMES;
        return;
    }
    lvreg_set_imu_prior(h_, imuAvailable ? 1 : 0, imuRollInit, imuPitchInit);      // consumed by transformUpdate, MO:1339
    lastStatus = lvreg_scan2map(h_, transformTobeMapped, &lastResult);
    if (lastStatus == LVREG_ERR_NOT_ENOUGH_FEATURES) {
        std::fprintf(stderr, "Not enough features! Only %d edge and %d planar features available.\n",
                     laserCloudCornerLastDSNum, laserCloudSurfLastDSNum);          // MO:1341
        return;
    }
    if (lastStatus != LVREG_OK) throw std::runtime_error(std::string("lvreg_scan2map: ") + lvreg_last_error(h_));
    isDegenerate = lastResult.degenerate != 0;        // transformUpdate (MO:1345-1375) ran inside lvreg_scan2map
    lvreg_timings t;
    lvreg_get_timings(h_, &t);
    lastTimings.register_ms = t.register_ms;
    lastTimings.kernel_launches += t.kernel_launches;
}

static void pose_affine(const float x, const float y, const float z, const float roll, const float pitch,
                        const float yaw, float T[12]) {
    const float pose[6] = {roll, pitch, yaw, x, y, z};
    lvreg_pose_to_affine(pose, T);
}

bool mapOptimization::saveFrame() {
    if (cloudKeyPoses3D.empty()) return true;
    const PointTypePose& b = cloudKeyPoses6D.back();
    if (P_.sensorIsLivox && timeLaserInfoCur - b.time > 1.0) return true;      // MO:1392-1396
    float S[12], F[12];
    pose_affine(b.x, b.y, b.z, b.roll, b.pitch, b.yaw, S);
    pose_affine(transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5], transformTobeMapped[0],
                transformTobeMapped[1], transformTobeMapped[2], F);
    // transBetween = transStart.inverse() * transFinal (rigid inverse = [R^T | -R^T t])
    float B[12];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            B[4 * i + j] = S[0 + i] * F[0 + j] + S[4 + i] * F[4 + j] + S[8 + i] * F[8 + j];
        B[4 * i + 3] = S[0 + i] * (F[3] - S[3]) + S[4 + i] * (F[7] - S[7]) + S[8 + i] * (F[11] - S[11]);
    }
    // pcl::getTranslationAndEulerAngles
    const float x = B[3], y = B[7], z = B[11];
    const float roll = std::atan2(B[9], B[10]), pitch = std::asin(-B[8]), yaw = std::atan2(B[4], B[0]);
    if (std::fabs(roll) < P_.surroundingkeyframeAddingAngleThreshold &&
        std::fabs(pitch) < P_.surroundingkeyframeAddingAngleThreshold &&
        std::fabs(yaw) < P_.surroundingkeyframeAddingAngleThreshold &&
        std::sqrt(x * x + y * y + z * z) < P_.surroundingkeyframeAddingDistThreshold)
        return false;
    return true;
}

void mapOptimization::saveKeyFramesAndFactor() {
    if (!saveFrame()) return;
    // No iSAM2 here: the optimised estimate is the LM pose itself.
    PointType p3 = make_point(transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5],
                              (float)cloudKeyPoses3D.size());
    PointTypePose p6;
    std::memset(&p6, 0, sizeof(p6));
    p6.x = p3.x; p6.y = p3.y; p6.z = p3.z; p6.data3 = 1.0f;
    p6.intensity = p3.intensity;
    p6.roll = transformTobeMapped[0]; p6.pitch = transformTobeMapped[1]; p6.yaw = transformTobeMapped[2];
    p6.time = timeLaserInfoCur;
    // the keyframe clouds are the DS feature clouds of this scan (MO:1600-1606); they are already
    // on the device, so they are handed over without a host round trip
    if (!scanDownsampled_) downsampleCurrentScan();
    int32_t id = -1;
    int st = lvreg_add_keyframe_from_scan(h_, transformTobeMapped, &id);
    if (st != LVREG_OK) throw std::runtime_error(std::string("lvreg_add_keyframe_from_scan: ") + lvreg_last_error(h_));
    cloudKeyPoses3D.push_back(p3);
    cloudKeyPoses6D.push_back(p6);
    if (keepHostKeyframeCopies) {
        cornerCloudKeyFrames.push_back(getLaserCloudLastDS(LVREG_CORNER));
        surfCloudKeyFrames.push_back(getLaserCloudLastDS(LVREG_SURF));
    }
}

void mapOptimization::correctPoses(const std::vector<PointTypePose>& corrected) {
    std::vector<float> poses(6 * corrected.size());
    for (size_t i = 0; i < corrected.size() && i < cloudKeyPoses6D.size(); ++i) {
        cloudKeyPoses6D[i] = corrected[i];
        cloudKeyPoses3D[i].x = corrected[i].x; cloudKeyPoses3D[i].y = corrected[i].y; cloudKeyPoses3D[i].z = corrected[i].z;
        float* p = &poses[6 * i];
        p[0] = corrected[i].roll; p[1] = corrected[i].pitch; p[2] = corrected[i].yaw;
        p[3] = corrected[i].x; p[4] = corrected[i].y; p[5] = corrected[i].z;
    }
    lvreg_update_keyframe_poses(h_, poses.data(), std::min(corrected.size(), cloudKeyPoses6D.size()));
    mapDirty_ = true;                                       // laserCloudMapContainer.clear() MO:1623
}

Cloud mapOptimization::transformPointCloud(const Cloud& cloudIn, const PointTypePose& t) {
    Cloud out(cloudIn.size());
    const float pose[6] = {t.roll, t.pitch, t.yaw, t.x, t.y, t.z};
    lvreg_cloud in = as_lvreg_cloud(cloudIn);
    lvreg_cloud_out lo = as_lvreg_out(out);
    int st = lvreg_transform_cloud(h_, &in, pose, &lo);
    if (st != LVREG_OK) throw std::runtime_error(std::string("lvreg_transform_cloud: ") + lvreg_last_error(h_));
    return out;
}

Cloud mapOptimization::getLaserCloudLastDS(int which) {
    size_t n = 0;
    lvreg_get_scan_ds(h_, which, nullptr, &n);
    Cloud c(n);
    lvreg_cloud_out lo = as_lvreg_out(c);
    int st = lvreg_get_scan_ds(h_, which, &lo, &n);
    if (st != LVREG_OK) throw std::runtime_error(std::string("lvreg_get_scan_ds: ") + lvreg_last_error(h_));
    return c;
}

Cloud mapOptimization::getLaserCloudFromMapDS(int which) {
    size_t n = 0;
    if (lvreg_get_local_map(h_, which, nullptr, &n) != LVREG_OK) return Cloud();
    Cloud c(n);
    lvreg_cloud_out lo = as_lvreg_out(c);
    int st = lvreg_get_local_map(h_, which, &lo, &n);
    if (st != LVREG_OK) throw std::runtime_error(std::string("lvreg_get_local_map: ") + lvreg_last_error(h_));
    return c;
}

mapOptimization::PinnedCloud::~PinnedCloud() {
    if (p) lvreg_host_free(p);
}
void mapOptimization::PinnedCloud::assign(const Cloud& c) {
    if (c.size() > cap) {
        if (p) lvreg_host_free(p);
        cap = c.size() + c.size() / 2 + 1024;
        p = (PointType*)lvreg_host_alloc(cap * sizeof(PointType));
        if (!p) { cap = 0; throw std::runtime_error("lvreg_host_alloc failed"); }
    }
    n = c.size();
    if (n) std::memcpy(p, c.data(), n * sizeof(PointType));
}

bool mapOptimization::laserCloudInfoHandler(const Cloud& corner, const Cloud& surf, double stamp, const float* guess) {
    timeLaserInfoCur = stamp;
    const bool fused = fusedHandler && !cloudKeyPoses3D.empty();
    if (fused) {                          // the one copy of the message (pcl::fromROSMsg, MO:305-307) lands in pinned memory
        pinCorner_.assign(corner);
        pinSurf_.assign(surf);
    } else {
        laserCloudCornerLast = corner;
        laserCloudSurfLast = surf;
    }
    if (!(timeLaserInfoCur - timeLastProcessing_ >= P_.mappingProcessInterval)) return false;   // MO:311-314
    timeLastProcessing_ = timeLaserInfoCur;
    std::memset(&lastTimings, 0, sizeof(lastTimings));
    scanDownsampled_ = false;
    if (guess) std::memcpy(transformTobeMapped, guess, sizeof(transformTobeMapped));   // updateInitialGuess
    if (!fusedHandler || cloudKeyPoses3D.empty()) {
        extractSurroundingKeyFrames();
        downsampleCurrentScan();
        scan2MapOptimization();
        saveKeyFramesAndFactor();
        return true;
    }
    // MO:318-322 as ONE library call: the local-map rebuild (when the selection changed) and the scan
    // down-sampling run concurrently on the device and share their host synchronisations
    std::vector<int32_t> ids = extractNearby();
    const bool rebuild = mapDirty_ || ids != lastIds_;
    lvreg_set_imu_prior(h_, imuAvailable ? 1 : 0, imuRollInit, imuPitchInit);
    lvreg_cloud c, s;
    c.data = pinCorner_.p; c.n = pinCorner_.n; c.stride = sizeof(PointType); c.intensity_offset = 16; c.on_device = 0; c.reserved = 0;
    s = c;
    s.data = pinSurf_.p; s.n = pinSurf_.n;
    lastStatus = lvreg_register_scan(h_, &c, &s, rebuild ? ids.data() : nullptr, ids.size(), transformTobeMapped, &lastResult);
    if (lastStatus != LVREG_OK && lastStatus != LVREG_ERR_NOT_ENOUGH_FEATURES)
        throw std::runtime_error(std::string("lvreg_register_scan: ") + lvreg_last_error(h_));
    if (rebuild) {
        lastIds_ = ids;
        mapDirty_ = false;
    }
    laserCloudCornerLastDSNum = lastResult.n_corner_ds;
    laserCloudSurfLastDSNum = lastResult.n_surf_ds;
    laserCloudCornerFromMapDSNum = lastResult.n_corner_map;
    laserCloudSurfFromMapDSNum = lastResult.n_surf_map;
    scanDownsampled_ = true;
    lvreg_timings t;
    lvreg_get_timings(h_, &t);
    lastTimings = t;
    if (lastStatus == LVREG_ERR_NOT_ENOUGH_FEATURES) {
        std::fprintf(stderr, "Not enough features! Only %d edge and %d planar features available.\n",
                     laserCloudCornerLastDSNum, laserCloudSurfLastDSNum);          // MO:1341
    } else {
        isDegenerate = lastResult.degenerate != 0;        // transformUpdate (MO:1339) ran inside the call (IMU prior set above)
    }
    saveKeyFramesAndFactor();
    return true;
}

}  // namespace lvreg_host
