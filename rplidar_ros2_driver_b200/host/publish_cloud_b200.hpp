// publish_cloud_b200.hpp -- host side of the PointCloud2 publisher: turns the
// [n][x, y, z, intensity] float32 points the cloud path leaves (rpl_cloud_batch / rpl_cloud_batch_dev) into a
// sensor_msgs::msg::PointCloud2 with the layout laser_geometry::LaserProjection produces for its default
// channels (fields x, y, z, intensity as FLOAT32, point_step 16, height 1).  The reference publishes
// LaserScan only (src/rplidar_node.cpp:679); this is the north star's PointCloud2 extension, so there is
// no reference line to mirror beyond the message definition.  A template so that this header compiles
// without ROS (tests use plain structs with the same members); in the node it is instantiated with
// sensor_msgs::msg::PointCloud2 / PointField.  When the bytes on the wire are all that is needed,
// rpl_pointcloud2_cdr_batch_dev serialises the same message on the device instead.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

namespace rplidar_b200 {

constexpr uint8_t kPointFieldFloat32 = 7;  // sensor_msgs/PointField.FLOAT32

template <class PointCloud2Msg, class Stamp>
void fill_pointcloud2_msg(PointCloud2Msg& msg, const float* xyzi, uint32_t n_points, const Stamp& stamp,
                          const std::string& frame_id) {
  msg.header.stamp = stamp;
  msg.header.frame_id = frame_id;
  msg.height = 1;  // unorganised cloud
  msg.width = n_points;
  static const char* const kNames[4] = {"x", "y", "z", "intensity"};
  msg.fields.resize(4);
  for (uint32_t f = 0; f < 4; ++f) {
    msg.fields[f].name = kNames[f];
    msg.fields[f].offset = 4 * f;
    msg.fields[f].datatype = kPointFieldFloat32;
    msg.fields[f].count = 1;
  }
  msg.is_bigendian = false;
  msg.point_step = 16;
  msg.row_step = 16 * n_points;
  msg.data.resize(static_cast<size_t>(16) * n_points);
  if (n_points) std::memcpy(msg.data.data(), xyzi, msg.data.size());
  msg.is_dense = true;  // the cloud path drops unmeasured points: no NaN rows
}

}  // namespace rplidar_b200
