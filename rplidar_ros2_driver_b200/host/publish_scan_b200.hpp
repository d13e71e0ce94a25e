// publish_scan_b200.hpp -- the tail of RPlidarNode::publish_scan (reference
// src/rplidar_node.cpp:613-679) once the arrays come from the GPU: copy LaserScanData into a
// sensor_msgs::msg::LaserScan.  A template so this header compiles without ROS (tests use a plain
// struct with the same members); in the node it is instantiated with the real message type.
#pragma once
#include <string>
#include <utility>

#include "cuda_scan_pipeline.hpp"

namespace rplidar_b200 {

template <class LaserScanMsg, class Stamp>
bool fill_laserscan_msg(LaserScanMsg& msg, LaserScanData&& d, const Stamp& stamp, const std::string& frame_id) {
  if (!d.publish) return false;  // reference returns without publishing (:558, :609-611)
  msg.header.stamp = stamp;      // :618
  msg.header.frame_id = frame_id;
  msg.angle_min = d.angle_min;
  msg.angle_max = d.angle_max;
  msg.angle_increment = d.angle_increment;
  msg.time_increment = d.time_increment;
  msg.scan_time = d.scan_time;
  msg.range_min = d.range_min;
  msg.range_max = d.range_max;
  msg.ranges = std::move(d.ranges);
  msg.intensities = std::move(d.intensities);
  return true;
}

}  // namespace rplidar_b200
