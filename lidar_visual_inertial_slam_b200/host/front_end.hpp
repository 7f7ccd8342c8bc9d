// front_end.hpp -- C++ host mirrors of the reference's LiDAR front-end classes for the rows 8f-4 / 8f-1:
//   class ImageProjection   (lidar_odometry/src/imageProjection.cpp:56-660): projectPointCloud,
//                           cloudExtraction, deskewPoint / findRotation inputs (imuTime, imuRot*)
//   class FeatureExtraction (lidar_odometry/src/featureExtraction.cpp:16-262): laserCloudInfoHandler =
//                           calculateSmoothness + markOccludedPoints + extractFeatures
// Member names follow the reference so that the bodies can be swapped one for one;
// every data-parallel step goes through the C ABI of liblvreg.  Not mirrored: ROS subscriptions, the
// message conversion (cachePointCloud), the IMU / odometry queue handling (imuDeskewInfo,
// odomDeskewInfo: sequential host code whose results -- imuTime, imuRot*, imuPointerCur -- are inputs).
#pragma once

#include <cstdint>
#include <vector>

#include "../../include/lvreg.h"
#include "map_optimization.hpp"

namespace lvreg_host {

// PointXYZIRT = LiovxPointCustomMsg (imageProjection.cpp:17-29, 48): 32 bytes
struct PointXYZIRT {
    float x, y, z, data3;
    float intensity;
    float time;
    uint16_t ring;
    uint16_t tag;
    float pad;
};
static_assert(sizeof(PointXYZIRT) == 32, "LiovxPointCustomMsg is 32 bytes");

enum class SensorType { VELODYNE = 0, OUSTER = 1, LIVOX = 2 };    // utility.h:62

// lidar_odometry/msg/CloudInfo.msg with the clouds as host vectors (what pcl::fromROSMsg yields)
struct CloudInfo {
    double stamp = 0.0;
    std::vector<int32_t> start_ring_index, end_ring_index;
    std::vector<int32_t> point_col_ind;
    std::vector<float> point_range;
    int64_t imu_available = 0, odom_available = 0;
    float imu_roll_init = 0, imu_pitch_init = 0, imu_yaw_init = 0;
    float initial_guess_x = 0, initial_guess_y = 0, initial_guess_z = 0;
    float initial_guess_roll = 0, initial_guess_pitch = 0, initial_guess_yaw = 0;
    int64_t odom_reset_id = 0;
    Cloud cloud_deskewed, cloud_corner, cloud_surface;
};

const int queueLength = 2000;                                     // imageProjection.cpp:50

class ImageProjection {
  public:
    // ParamServer fields the class reads (utility.h:97-101, 213-222)
    int N_SCAN = 4, Horizon_SCAN = 6000, downsampleRate = 1;
    float lidarMinRange = 1.0f, lidarMaxRange = 1000.0f;
    SensorType sensor = SensorType::LIVOX;

    // state that deskewInfo() leaves behind (imageProjection.cpp:77-91)
    std::vector<double> imuTime, imuRotX, imuRotY, imuRotZ;
    int imuPointerCur = 0;
    int deskewFlag = 0;
    double timeScanCur = 0.0;
    std::vector<PointXYZIRT> laserCloudIn;
    CloudInfo cloudInfo;

    explicit ImageProjection(lvreg_handle* shared_handle);        // the handle of the mapOptimization mirror, or an own one
    // projectPointCloud() + cloudExtraction() (imageProjection.cpp:571-647) in one device call.
    // downloadCloudInfo = false keeps everything on the device for a FeatureExtraction in the same process.
    void projectPointCloud(bool downloadCloudInfo = true);
    void resetParameters();                                       // imageProjection.cpp:167-187
    size_t extractedCloudSize() const { return nExtracted_; }

  private:
    lvreg_handle* h_;
    size_t nExtracted_ = 0;
};

class FeatureExtraction {
  public:
    float edgeThreshold = 1.0f, surfThreshold = 0.1f, odometrySurfLeafSize = 0.4f;   // utility.h:255-265
    int N_SCAN = 4;
    Cloud cornerCloud, surfaceCloud;

    explicit FeatureExtraction(lvreg_handle* shared_handle);
    // laserCloudInfoHandler (featureExtraction.cpp:72-85) on a CloudInfo that carries host arrays
    void laserCloudInfoHandler(CloudInfo& cloudInfo, bool downloadFeatures = true);
    // the same on the device-resident result of ImageProjection::projectPointCloud(false): no host copy of
    // the deskewed cloud in between; the features stay on the device for lvreg_register_scan
    void laserCloudInfoHandlerOnDevice(bool downloadFeatures = false);
    size_t numCorner() const { return nCorner_; }
    size_t numSurface() const { return nSurf_; }

  private:
    lvreg_handle* h_;
    size_t nCorner_ = 0, nSurf_ = 0;
};

}  // namespace lvreg_host
