// front_end.cpp -- see front_end.hpp
#include "front_end.hpp"

#include <cstddef>
#include <stdexcept>
#include <string>

namespace lvreg_host {

static void check(lvreg_handle* h, int st, const char* what) {
    if (st != LVREG_OK) throw std::runtime_error(std::string(what) + ": " + lvreg_last_error(h));
}

static lvreg_cloud_out out_desc(Cloud& c) {
    lvreg_cloud_out o;
    o.data = c.data();
    o.capacity = c.size();
    o.stride = sizeof(PointType);
    o.intensity_offset = offsetof(PointType, intensity);
    o.on_device = 0;
    o.reserved = 0;
    return o;
}

ImageProjection::ImageProjection(lvreg_handle* shared_handle) : h_(shared_handle) {
    if (!h_) throw std::runtime_error("ImageProjection needs a liblvreg handle (a CUDA device is required)");
    resetParameters();
}

void ImageProjection::resetParameters() {
    laserCloudIn.clear();
    imuPointerCur = 0;
    imuTime.assign(queueLength, 0.0);
    imuRotX.assign(queueLength, 0.0);
    imuRotY.assign(queueLength, 0.0);
    imuRotZ.assign(queueLength, 0.0);
}

void ImageProjection::projectPointCloud(bool downloadCloudInfo) {
    lvreg_raw_cloud raw;
    raw.data = laserCloudIn.data();
    raw.n = laserCloudIn.size();
    raw.stride = sizeof(PointXYZIRT);
    raw.intensity_offset = offsetof(PointXYZIRT, intensity);
    raw.ring_offset = offsetof(PointXYZIRT, ring);
    raw.time_offset = offsetof(PointXYZIRT, time);
    raw.on_device = 0;
    raw.reserved = 0;
    lvreg_projection_params pp;
    pp.n_scan = N_SCAN;
    pp.horizon_scan = Horizon_SCAN;
    pp.downsample_rate = downsampleRate;
    pp.sensor = (int)sensor;
    pp.lidar_min_range = lidarMinRange;
    pp.lidar_max_range = lidarMaxRange;
    pp.deskew = (deskewFlag != -1 && cloudInfo.imu_available) ? 1 : 0;      // imageProjection.cpp:540
    pp.imu_pointer_cur = imuPointerCur;
    pp.time_scan_cur = timeScanCur;
    pp.imu_time = imuTime.data();
    pp.imu_rot_x = imuRotX.data();
    pp.imu_rot_y = imuRotY.data();
    pp.imu_rot_z = imuRotZ.data();
    check(h_, lvreg_project_cloud(h_, &raw, &pp, &nExtracted_), "lvreg_project_cloud");
    cloudInfo.start_ring_index.assign(N_SCAN, 0);
    cloudInfo.end_ring_index.assign(N_SCAN, 0);
    if (!downloadCloudInfo) {
        check(h_, lvreg_download_projection(h_, nullptr, nullptr, nullptr, cloudInfo.start_ring_index.data(),
                                            cloudInfo.end_ring_index.data(), nullptr), "lvreg_download_projection");
        return;
    }
    cloudInfo.point_col_ind.assign(nExtracted_, 0);
    cloudInfo.point_range.assign(nExtracted_, 0.f);
    cloudInfo.cloud_deskewed.assign(nExtracted_, PointType());
    lvreg_cloud_out o = out_desc(cloudInfo.cloud_deskewed);
    check(h_, lvreg_download_projection(h_, &o, cloudInfo.point_range.data(), cloudInfo.point_col_ind.data(),
                                        cloudInfo.start_ring_index.data(), cloudInfo.end_ring_index.data(), nullptr),
          "lvreg_download_projection");
    for (PointType& p : cloudInfo.cloud_deskewed) p.data3 = 1.0f;           // PCL_ADD_POINT4D
}

FeatureExtraction::FeatureExtraction(lvreg_handle* shared_handle) : h_(shared_handle) {
    if (!h_) throw std::runtime_error("FeatureExtraction needs a liblvreg handle (a CUDA device is required)");
}

void FeatureExtraction::laserCloudInfoHandler(CloudInfo& ci, bool downloadFeatures) {
    lvreg_cloud in = as_lvreg_cloud(ci.cloud_deskewed);
    lvreg_scan_info info;
    info.start_ring_index = ci.start_ring_index.data();
    info.end_ring_index = ci.end_ring_index.data();
    info.n_scan = (int32_t)ci.start_ring_index.size();
    info.reserved = 0;
    info.point_col_ind = ci.point_col_ind.data();
    info.point_range = ci.point_range.data();
    cornerCloud.assign(downloadFeatures ? ci.cloud_deskewed.size() : 0, PointType());
    surfaceCloud.assign(downloadFeatures ? ci.cloud_deskewed.size() : 0, PointType());
    lvreg_cloud_out co = out_desc(cornerCloud), so = out_desc(surfaceCloud);
    check(h_, lvreg_extract_features(h_, &in, &info, edgeThreshold, surfThreshold, odometrySurfLeafSize,
                                     downloadFeatures ? &co : nullptr, &nCorner_, downloadFeatures ? &so : nullptr, &nSurf_,
                                     nullptr), "lvreg_extract_features");
    if (downloadFeatures) {
        cornerCloud.resize(nCorner_);
        surfaceCloud.resize(nSurf_);
        ci.cloud_corner = cornerCloud;                               // publishFeatureCloud, featureExtraction.cpp:253-260
        ci.cloud_surface = surfaceCloud;
    }
}

void FeatureExtraction::laserCloudInfoHandlerOnDevice(bool downloadFeatures) {
    lvreg_cloud in;
    lvreg_scan_info info;
    if (lvreg_get_projection(h_, &in, &info) != LVREG_OK) throw std::runtime_error("no projected scan on the device");
    cornerCloud.assign(downloadFeatures ? in.n : 0, PointType());
    surfaceCloud.assign(downloadFeatures ? in.n : 0, PointType());
    lvreg_cloud_out co = out_desc(cornerCloud), so = out_desc(surfaceCloud);
    check(h_, lvreg_extract_features(h_, &in, &info, edgeThreshold, surfThreshold, odometrySurfLeafSize,
                                     downloadFeatures ? &co : nullptr, &nCorner_, downloadFeatures ? &so : nullptr, &nSurf_,
                                     nullptr), "lvreg_extract_features");
    if (downloadFeatures) {
        cornerCloud.resize(nCorner_);
        surfaceCloud.resize(nSurf_);
    }
}

}  // namespace lvreg_host
