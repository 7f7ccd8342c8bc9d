// wire_formats.hpp -- the disk / wire formats on either side of the scan-to-map path:
//   * lidar_odometry/msg/CloudInfo.msg:1-32 as a ROS 2 message body: CDR (XCDR1, little endian, 4-byte
//     encapsulation header), i.e. what rmw hands to / takes from the transport and what rosbag2 stores.  The three
//     sensor_msgs/PointCloud2 members carry pcl::PointXYZI clouds exactly as pcl::toROSMsg lays them out
//     (fields x, y, z at 0 / 4 / 8, intensity at 16, point_step 32).
//   * binary PCD v0.7 as pcl::io::savePCDFileBinary writes it for the save-map service (MO:179-236):
//     trajectory.pcd (PointXYZI), transformations.pcd (PointXYZIRPYT), CornerMap / SurfMap / GlobalMap.pcd.
// No ROS, no PCL: plain byte buffers in and out.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "front_end.hpp"

namespace lvreg_host {

// ---- CloudInfo <-> CDR ----
std::vector<uint8_t> serialize_cloud_info(const CloudInfo& msg, const std::string& frame_id);
// throws std::runtime_error on a truncated / malformed buffer
CloudInfo deserialize_cloud_info(const uint8_t* data, size_t size, std::string* frame_id = nullptr);

// ---- PCD ----
// pcl::io::savePCDFileBinary for PointXYZI / PointXYZIRPYT: header + the fields only (PCL drops the padding)
bool save_pcd_binary(const std::string& path, const Cloud& cloud);
bool save_pcd_binary(const std::string& path, const std::vector<PointTypePose>& poses);
// reads back a binary PCD with float fields (x y z intensity [...]); used by the tests and the replay tools
bool load_pcd_binary_xyzi(const std::string& path, Cloud* cloud);

}  // namespace lvreg_host
