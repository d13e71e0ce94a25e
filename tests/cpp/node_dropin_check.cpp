// node_dropin_check.cpp -- the INTEGRATION.md section 3 replacement of RPlidarNode::publish_scan, compiled
// against the reference's own node class and run next to the reference's own publish_scan.
//
// The reference's node translation unit is included where it lies under /root/reference (nothing is copied)
// and built against the ROS 2 API stubs of oracle/ros_stubs/ (rclcpp is not installed in this image).
// publish_scan_b200() below is the body INTEGRATION.md proposes, written as a free function over the node
// (in the patched driver it is the member itself and `pipeline` is a member of the node).
//   node_dropin_check cpu                    builds the node, the drop-in must refuse to run without a B200
//   node_dropin_check gpu <golden_raw.bin>   16 captured dummy scans: for every scan and every
//                                            (protocol, mode, inverted) the message published by the drop-in
//                                            must equal the message published by the reference, bit for bit
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>

#include "ros_stub_core.hpp"
#define private public
#define protected public
#define main reference_node_main
#include "rplidar_node.cpp"  // -I /root/reference/src -I /root/reference/include
#undef main
#undef private
#undef protected

#include "cuda_scan_pipeline.hpp"
#include "publish_scan_b200.hpp"

namespace {

// ---- INTEGRATION.md section 3 ------------------------------------------------------------------------------------
bool publish_scan_b200(RPlidarNode& self, rplidar_b200::CudaScanPipeline& pipeline,
                       const std::vector<sl_lidar_response_measurement_node_hq_t>& nodes, rclcpp::Time start_time,
                       double scan_duration) {
  if (nodes.empty()) return true;
  bool is_new_protocol = false;
  auto real_drv = dynamic_cast<RealLidarDriver*>(self.driver_.get());
  if (real_drv && real_drv->is_new_type()) is_new_protocol = true;

  rplidar_b200::LaserScanData d;
  if (!pipeline.laserscan(nodes, is_new_protocol, self.params_.scan_processing, self.params_.inverted, scan_duration,
                          self.cached_current_max_range_, d))
    return false;
  sensor_msgs::msg::LaserScan scan_msg;
  if (rplidar_b200::fill_laserscan_msg(scan_msg, std::move(d), start_time, self.params_.frame_id))
    self.scan_pub_->publish(scan_msg);
  return true;
}
// --------------------------------------------------------------------------------------------------------------------

int fail(const char* what) {
  std::printf("FAIL: %s\n", what);
  return 1;
}
bool same_bits(const std::vector<float>& a, const std::vector<float>& b) {
  return a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * sizeof(float)) == 0);
}
bool same_bits(float a, float b) { return std::memcmp(&a, &b, sizeof(float)) == 0; }

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) return fail("usage");
  const std::string mode = argv[1];
  RPlidarNode node;  // declares its parameters against the stubs
  node.scan_pub_ = std::make_shared<rclcpp_lifecycle::LifecyclePublisher<sensor_msgs::msg::LaserScan>>();
  node.params_.frame_id = "laser_frame";
  if (mode == "cpu") {
    bool threw = false;
    try {
      rplidar_b200::CudaScanPipeline p(0, 8192, 1);
    } catch (const std::exception&) {
      threw = true;
    }
    std::printf("OK cpu (node built; pipeline without a device %s)\n", threw ? "refused: no CPU fallback" : "constructed");
    return 0;
  }
  if (argc < 3) return fail("usage gpu");
  std::ifstream f(argv[2], std::ios::binary);
  const std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (raw.size() != 16u * 360u * 8u) return fail("golden_raw size");
  rplidar_b200::CudaScanPipeline pipeline(0, 8192, 1);
  std::unique_ptr<LidarDriverInterface> real(new RealLidarDriver()), dummy(new DummyLidarDriver());
  sensor_msgs::msg::LaserScan got;
  int published = 0;
  ros_stub::laserscan_sink() = [&](const sensor_msgs::msg::LaserScan& m) {
    got = m;
    ++published;
  };
  int checked = 0;
  for (int scan = 0; scan < 16; ++scan) {
    std::vector<sl_lidar_response_measurement_node_hq_t> nodes(360);
    std::memcpy(nodes.data(), raw.data() + static_cast<size_t>(scan) * 360 * 8, 360 * 8);
    if (scan % 3 == 1)
      for (int i = 0; i < 360; i += 7) nodes[i].dist_mm_q2 = 0;  // some unmeasured nodes
    for (int cfg = 0; cfg < 8; ++cfg) {
      const bool newp = cfg & 1, mode_a = cfg & 2, inv = cfg & 4;
      static_cast<RealLidarDriver*>(real.get())->profile_.protocol = newp ? ProtocolType::NEW_TYPE : ProtocolType::OLD_TYPE;
      node.driver_.release();
      node.driver_.reset(newp ? real.get() : dummy.get());
      node.params_.scan_processing = mode_a;
      node.params_.inverted = inv;
      node.cached_current_max_range_ = 12.0f + static_cast<float>(scan);
      const double duration = 0.1 + 0.001 * scan;
      published = 0;
      node.publish_scan(nodes, rclcpp::Time(123456789), duration);  // the reference
      if (published != 1) return fail("reference did not publish");
      const sensor_msgs::msg::LaserScan ref = got;
      published = 0;
      if (!publish_scan_b200(node, pipeline, nodes, rclcpp::Time(123456789), duration)) return fail(pipeline.last_error());
      if (published != 1) return fail("drop-in did not publish");
      if (!same_bits(got.ranges, ref.ranges) || !same_bits(got.intensities, ref.intensities)) return fail("arrays differ");
      if (!same_bits(got.angle_min, ref.angle_min) || !same_bits(got.angle_max, ref.angle_max) ||
          !same_bits(got.angle_increment, ref.angle_increment) || !same_bits(got.time_increment, ref.time_increment) ||
          !same_bits(got.scan_time, ref.scan_time) || !same_bits(got.range_min, ref.range_min) ||
          !same_bits(got.range_max, ref.range_max))
        return fail("header scalars differ");
      if (got.header.frame_id != ref.header.frame_id || got.header.stamp.sec != ref.header.stamp.sec ||
          got.header.stamp.nanosec != ref.header.stamp.nanosec)
        return fail("header differs");
      ++checked;
    }
  }
  node.driver_.release();
  ros_stub::laserscan_sink() = nullptr;
  std::printf("OK gpu: %d messages identical to the reference node's\n", checked);
  return 0;
}
