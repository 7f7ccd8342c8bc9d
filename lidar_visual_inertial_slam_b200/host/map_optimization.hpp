// map_optimization.hpp -- C++ host mirror of the reference's `class mapOptimization`
// (lidar_odometry/src/mapOptimization.cpp:49-1782) for the scan-to-map path only.  Member
// functions and data members keep the reference's names and meaning so that a maintainer can
// swap the bodies one for one; every data-parallel step goes through the C ABI
// of liblvreg (include/lvreg.h).  Not mirrored (out of scope): iSAM2 factors,
// GPS, ROS publishing -- saveKeyFramesAndFactor takes the LM pose as the keyframe pose, which is
// what the replay harness needs.  Loop closure is mirrored up to the pose constraint:
// detectLoopClosureDistance + performLoopClosure queue {poseFrom, poseTo, noise}; the gtsam
// BetweenFactor built from them (MO:610-619, 1488-1527) stays with the caller.
#pragma once

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/lvreg.h"

namespace lvreg_host {

// pcl::PointXYZI layout (utility.h:64): 32 bytes, {x,y,z,1 | intensity,0,0,0}
struct PointType {
    float x, y, z, data3;
    float intensity, pad0, pad1, pad2;
};
static_assert(sizeof(PointType) == 32, "pcl::PointXYZI is 32 bytes");

// PointXYZIRPYT (MO:29-46): 48 bytes, time at offset 32
struct PointTypePose {
    float x, y, z, data3;
    float intensity, roll, pitch, yaw;
    double time;
    double pad;
};
static_assert(sizeof(PointTypePose) == 48, "PointXYZIRPYT is 48 bytes");

typedef std::vector<PointType> Cloud;

// the ParamServer fields the path reads (utility.h:119-140, 256-303)
struct ParamServer {
    lvreg_params lv;
    float surroundingKeyframeSearchRadius = 50.0f;          // utility.h:285
    float surroundingKeyframeDensity = 2.0f;                // utility.h:287
    float surroundingkeyframeAddingDistThreshold = 1.0f;    // utility.h:279
    float surroundingkeyframeAddingAngleThreshold = 0.2f;   // utility.h:281
    bool  sensorIsLivox = true;                             // MO:1392-1396: keyframe every > 1.0 s
    double mappingProcessInterval = 0.15;                   // utility.h:283, MO:311-314
    float globalMapVisualizationSearchRadius = 1000.0f;     // utility.h:306
    float globalMapVisualizationPoseDensity = 10.0f;        // utility.h:308
    float globalMapVisualizationLeafSize = 1.0f;            // utility.h:310
    float historyKeyframeSearchRadius = 15.0f;              // utility.h:296
    float historyKeyframeSearchTimeDiff = 30.0f;            // utility.h:298
    int   historyKeyframeSearchNum = 25;                    // utility.h:300
    float historyKeyframeFitnessScore = 0.3f;               // utility.h:302
    // optional capacity hints (0 = grow on demand): points entering the local-map VoxelGrid per class, raw feature
    // points per scan, cells of a search grid -- see lvreg_reserve
    size_t reserveMapCorner = 0, reserveMapSurf = 0, reserveScanCorner = 0, reserveScanSurf = 0, reserveGridCells = 0;
    ParamServer() { lvreg_default_params(&lv); }
};

class mapOptimization {
  public:
    explicit mapOptimization(const ParamServer& params = ParamServer(), int device = 0);
    ~mapOptimization();
    mapOptimization(const mapOptimization&) = delete;
    mapOptimization& operator=(const mapOptimization&) = delete;

    // ---- data members that are de-facto API in the reference ----
    std::vector<Cloud> cornerCloudKeyFrames, surfCloudKeyFrames;     // MO:83-84 (host copies)
    std::vector<PointType> cloudKeyPoses3D;                          // MO:86 (intensity = index)
    std::vector<PointTypePose> cloudKeyPoses6D;                      // MO:87
    Cloud laserCloudCornerLast, laserCloudSurfLast;                  // MO:91-92
    int laserCloudCornerLastDSNum = 0, laserCloudSurfLastDSNum = 0;
    int laserCloudCornerFromMapDSNum = 0, laserCloudSurfFromMapDSNum = 0;   // MO:134-135
    float transformTobeMapped[6] = {0, 0, 0, 0, 0, 0};               // MO:126
    bool isDegenerate = false;                                       // MO:131
    double timeLaserInfoCur = 0.0;
    bool imuAvailable = false;                                       // cloudInfo.imu_available
    float imuRollInit = 0.f, imuPitchInit = 0.f;
    lvreg_result lastResult;                                         // iterations, n_sel, ...
    lvreg_timings lastTimings;
    int lastStatus = LVREG_OK;
    bool keepHostKeyframeCopies = false;
    bool fusedHandler = true;        // laserCloudInfoHandler issues MO:318-322 as one lvreg_register_scan call

    // forget the session (key poses, keyframe clouds, local map, LM and loop state); device buffers are kept,
    // so a new sequence starts without re-allocating anything
    void reset();

    // ---- the path, same names as the reference ----
    void extractSurroundingKeyFrames();      // MO:972-985 -> extractNearby + extractCloud
    void downsampleCurrentScan();            // MO:987-999
    void scan2MapOptimization();             // MO:1315-1343
    bool saveFrame();                        // MO:1387-1412
    void saveKeyFramesAndFactor();           // MO:1529-1613 without GTSAM
    void correctPoses(const std::vector<PointTypePose>& corrected);   // MO:1615-1646
    Cloud transformPointCloud(const Cloud& cloudIn, const PointTypePose& transformIn);   // MO:347-366
    // laserCloudInfoHandler body MO:316-326 for one incoming scan; `guess` replaces
    // updateInitialGuess (MO:806-877, out of scope).  Returns false when throttled (MO:311-314).
    bool laserCloudInfoHandler(const Cloud& corner, const Cloud& surf, double stamp, const float* guess);

    // ---- loop closure (MO:549-661), same names as the reference ----
    struct LoopConstraint {                  // one entry of loopIndexQueue / loopPoseQueue / loopNoiseQueue
        int loopKeyCur, loopKeyPre;
        float poseFrom[6], poseTo[6];        // {roll,pitch,yaw,x,y,z}; the factor is poseFrom.between(poseTo)
        float noiseScore;
    };
    std::map<int, int> loopIndexContainer;                           // MO:107
    std::vector<LoopConstraint> loopQueue;                           // MO:108-110
    lvreg_loop_result lastLoop;
    bool detectLoopClosureDistance(int* latestID, int* closestID);   // MO:630-661
    bool performLoopClosure();                                       // MO:549-628; true when a constraint was queued

    // publishGlobalMap (MO:460-510) without the ROS publisher: returns globalMapKeyFramesDS
    Cloud publishGlobalMap();
    // the lio_sam/save_map service (MO:179-236): trajectory.pcd, transformations.pcd, CornerMap.pcd, SurfMap.pcd and
    // GlobalMap.pcd (binary PCD v0.7, as pcl::io::savePCDFileBinary writes them) into `directory`, which must exist;
    // resolution != 0 down-samples the corner / surf maps with that leaf (MO:206-218).  The transform + concatenate
    // (+ VoxelGrid) of all keyframe clouds runs on the device.
    bool saveMap(const std::string& directory, float resolution);

    // read-backs of device-resident clouds (laserCloud*LastDS / *FromMapDS)
    Cloud getLaserCloudLastDS(int which);
    Cloud getLaserCloudFromMapDS(int which);
    const std::vector<int32_t>& lastKeyframeSelection() const { return lastIds_; }
    lvreg_handle* handle() { return h_; }

    // extractNearby's id list (MO:894-929 + the distance filter of MO:938-939)
    std::vector<int32_t> extractNearby();

  private:
    // page-locked landing buffers of the incoming feature clouds: pcl::fromROSMsg (MO:305-307) copies the message
    // into laserCloud*Last anyway; copying into pinned memory instead makes the upload an asynchronous DMA transfer
    struct PinnedCloud {
        PointType* p = nullptr;
        size_t cap = 0, n = 0;
        ~PinnedCloud();
        void assign(const Cloud& c);
    };
    PinnedCloud pinCorner_, pinSurf_;
    ParamServer P_;
    lvreg_handle* h_ = nullptr;
    std::vector<int32_t> lastIds_;
    bool mapDirty_ = true;
    double timeLastProcessing_ = -1;
    bool scanDownsampled_ = false;
};

// helpers shared with the harness
PointType make_point(float x, float y, float z, float intensity);
Cloud cloud_from_xyzi(const float* xyzi, size_t n);
lvreg_cloud as_lvreg_cloud(const Cloud& c);

}  // namespace lvreg_host
