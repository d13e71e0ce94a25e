// rplidar_b200_node_accel.hpp -- what the patched RPlidarNode owns (ros2_overlay/patches/0002): the B200 side of
// publish_scan plus the PointCloud2 publisher of the north star.
//
// Replaces the compute body of RPlidarNode::publish_scan (reference src/rplidar_node.cpp:566-677: filter +
// fixed-point unpack :581-600, sort :605-607, Mode A :630-660, Mode B :661-677) with one call into
// librplidar_b200 and keeps everything the node decides on the host exactly where it was: protocol detection
// (:575-579), header scalars (:616-625), the lifecycle publisher (:679).  There is no CPU fallback: when the
// CUDA call fails the scan is dropped and the error is logged, as a failed grab is in scan_loop.
//
// Additive parameter surface (all default to "off", so an existing launch file behaves as before):
//   b200_device          int     0       CUDA ordinal
//   publish_pointcloud   bool    false   also publish sensor_msgs/PointCloud2 on `cloud_topic`
//   cloud_topic          string  "cloud"
//   cloud_range_min/max  double  0.15 / (the scan's range_max)   laser_geometry-style range window
//   cloud_intensity_min  double  0.0
//   cloud_voxel_size     double  0.0     metres, 0 = no voxel grid
//   cloud_sor_k          int     0       0 = no statistical outlier removal
//   cloud_sor_alpha      double  1.0
#pragma once
#include <memory>
#include <string>
#include <vector>

#include <rclcpp/rclcpp.hpp>
#include <rclcpp_lifecycle/lifecycle_node.hpp>
#include <rclcpp_lifecycle/lifecycle_publisher.hpp>
#include <sensor_msgs/msg/laser_scan.hpp>
#include <sensor_msgs/msg/point_cloud2.hpp>

#include "cuda_scan_pipeline.hpp"
#include "publish_cloud_b200.hpp"
#include "publish_scan_b200.hpp"

namespace rplidar_b200 {

class NodeAccel {
 public:
  explicit NodeAccel(rclcpp_lifecycle::LifecycleNode& node) : node_(node) {
    int device = 0;
    declare("b200_device", 0, device);
    declare("publish_pointcloud", false, publish_cloud_);
    declare("cloud_topic", std::string("cloud"), cloud_topic_);
    declare("cloud_range_min", 0.15, cloud_range_min_);
    declare("cloud_range_max", 0.0, cloud_range_max_);  // 0 = follow the scan's range_max
    declare("cloud_intensity_min", 0.0, cloud_intensity_min_);
    declare("cloud_voxel_size", 0.0, cloud_voxel_);
    declare("cloud_sor_k", 0, cloud_sor_k_);
    declare("cloud_sor_alpha", 1.0, cloud_sor_alpha_);
    pipeline_ = std::make_unique<CudaScanPipeline>(device, 8192, 1);  // throws without a B200: no CPU fallback
    if (publish_cloud_) {
      cloud_pub_ = node_.create_publisher<sensor_msgs::msg::PointCloud2>(cloud_topic_, rclcpp::SensorDataQoS());
      cloud_pub_->on_activate();
    }
  }

  // the body of RPlidarNode::publish_scan after its empty-input check (:558)
  void publish_scan(const std::vector<sl_lidar_response_measurement_node_hq_t>& nodes, const rclcpp::Time& start_time,
                    double scan_duration, bool is_new_protocol, bool scan_processing, bool inverted, float range_max,
                    const std::string& frame_id,
                    rclcpp_lifecycle::LifecyclePublisher<sensor_msgs::msg::LaserScan>& scan_pub) {
    LaserScanData d;
    if (!pipeline_->laserscan(nodes, is_new_protocol, scan_processing, inverted, scan_duration, range_max, d)) {
      RCLCPP_ERROR(node_.get_logger(), "librplidar_b200: %s", pipeline_->last_error());
      return;
    }
    sensor_msgs::msg::LaserScan scan_msg;
    if (fill_laserscan_msg(scan_msg, std::move(d), start_time, frame_id)) scan_pub.publish(scan_msg);
    if (!cloud_pub_) return;
    rpl_cloud_params cp{};
    cp.range_min = static_cast<float>(cloud_range_min_);
    cp.range_max = cloud_range_max_ > 0.0 ? static_cast<float>(cloud_range_max_) : range_max;
    cp.intensity_min = static_cast<float>(cloud_intensity_min_);
    cp.voxel_size = static_cast<float>(cloud_voxel_);
    cp.sor_k = static_cast<uint32_t>(cloud_sor_k_ < 0 ? 0 : cloud_sor_k_);
    cp.sor_alpha = static_cast<float>(cloud_sor_alpha_);
    cp.is_new_protocol = is_new_protocol;
    uint32_t n_points = 0;
    if (!pipeline_->cloud(nodes, cp, xyzi_, n_points)) {
      RCLCPP_ERROR(node_.get_logger(), "librplidar_b200 (cloud): %s", pipeline_->last_error());
      return;
    }
    sensor_msgs::msg::PointCloud2 cloud_msg;
    fill_pointcloud2_msg(cloud_msg, xyzi_.data(), n_points, builtin_interfaces::msg::Time(start_time), frame_id);
    cloud_pub_->publish(cloud_msg);
  }

  CudaScanPipeline& pipeline() { return *pipeline_; }

 private:
  template <class T>
  void declare(const std::string& name, const T& def, T& out) {
    out = node_.has_parameter(name) ? out : node_.declare_parameter<T>(name, def);
    node_.get_parameter(name, out);
  }

  rclcpp_lifecycle::LifecycleNode& node_;
  std::unique_ptr<CudaScanPipeline> pipeline_;
  rclcpp_lifecycle::LifecyclePublisher<sensor_msgs::msg::PointCloud2>::SharedPtr cloud_pub_;
  std::vector<float> xyzi_;
  bool publish_cloud_ = false;
  std::string cloud_topic_;
  double cloud_range_min_ = 0.15, cloud_range_max_ = 0.0, cloud_intensity_min_ = 0.0, cloud_voxel_ = 0.0,
         cloud_sor_alpha_ = 1.0;
  int cloud_sor_k_ = 0;
};

}  // namespace rplidar_b200
