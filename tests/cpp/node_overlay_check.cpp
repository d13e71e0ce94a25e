// node_overlay_check.cpp -- the PATCHED reference node (ros2_overlay/patches applied to copies of the reference's
// own sources by ros2_overlay/apply.sh; nothing of the reference lives in this repository) compiled against the
// ROS 2 API stubs of oracle/ros_stubs/ and driven like scan_loop drives it.
//   node_overlay_check cpu                       the patched node builds and constructs; without a B200 the first
//                                                publish_scan must refuse (std::runtime_error: no CPU fallback)
//   node_overlay_check gpu <raw.bin> <out.bin>   16 captured dummy scans x 8 (protocol, mode, inverted): the
//                                                LaserScan AND PointCloud2 messages the patched node publishes are
//                                                written to out.bin for the Python side to compare with the
//                                                reference's own publish_scan and with the cloud definition
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>

#include "ros_stub_core.hpp"
#define private public
#define protected public
#define main patched_node_main
#include "rplidar_node.cpp"  // the PATCHED copy: -I <overlay build dir>/src comes first
#undef main
#undef private
#undef protected

namespace {
int fail(const char* what) {
  std::printf("FAIL: %s\n", what);
  return 1;
}
template <class T>
void put(std::ofstream& f, const T& v) {
  f.write(reinterpret_cast<const char*>(&v), sizeof(T));
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) return fail("usage");
  const std::string mode = argv[1];
  RPlidarNode node;
  node.scan_pub_ = std::make_shared<rclcpp_lifecycle::LifecyclePublisher<sensor_msgs::msg::LaserScan>>();
  node.params_.frame_id = "laser_frame";
  node.driver_.reset(new DummyLidarDriver());
  std::vector<sl_lidar_response_measurement_node_hq_t> one(360);
  for (int i = 0; i < 360; ++i) {
    one[i].angle_z_q14 = static_cast<uint16_t>(i * 182);
    one[i].dist_mm_q2 = 4000;
  }
  if (mode == "cpu") {
    bool threw = false;
    try {
      node.publish_scan(one, rclcpp::Time(1), 0.1);
    } catch (const std::exception& e) {
      threw = true;
      std::printf("publish_scan without a device: %s\n", e.what());
    }
    std::printf("OK cpu (patched node built; publish_scan %s)\n", threw ? "refused: no CPU fallback" : "ran");
    return 0;
  }
  if (argc < 4) return fail("usage gpu");
  std::ifstream f(argv[2], std::ios::binary);
  const std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (raw.size() != 16u * 360u * 8u) return fail("golden_raw size");
  std::ofstream out(argv[3], std::ios::binary);
  // the additive parameters, set the way a launch file would before the node reads them
  node.stub_override_parameter("publish_pointcloud", true);
  node.stub_override_parameter("cloud_voxel_size", 0.05);
  node.stub_override_parameter("cloud_sor_k", 8);
  std::unique_ptr<LidarDriverInterface> real(new RealLidarDriver()), dummy(new DummyLidarDriver());
  sensor_msgs::msg::LaserScan scan;
  sensor_msgs::msg::PointCloud2 cloud;
  int n_scan = 0, n_cloud = 0;
  ros_stub::laserscan_sink() = [&](const sensor_msgs::msg::LaserScan& m) {
    scan = m;
    ++n_scan;
  };
  ros_stub::pointcloud2_sink() = [&](const sensor_msgs::msg::PointCloud2& m) {
    cloud = m;
    ++n_cloud;
  };
  node.driver_.release();
  int written = 0;
  for (int s = 0; s < 16; ++s) {
    std::vector<sl_lidar_response_measurement_node_hq_t> nodes(360);
    std::memcpy(nodes.data(), raw.data() + static_cast<size_t>(s) * 360 * 8, 360 * 8);
    if (s % 3 == 1)
      for (int i = 0; i < 360; i += 7) nodes[i].dist_mm_q2 = 0;
    for (int cfg = 0; cfg < 8; ++cfg) {
      const bool newp = cfg & 1, mode_a = cfg & 2, inv = cfg & 4;
      static_cast<RealLidarDriver*>(real.get())->profile_.protocol = newp ? ProtocolType::NEW_TYPE : ProtocolType::OLD_TYPE;
      node.driver_.release();
      node.driver_.reset(newp ? real.get() : dummy.get());
      node.params_.scan_processing = mode_a;
      node.params_.inverted = inv;
      node.cached_current_max_range_ = 12.0f + static_cast<float>(s);
      n_scan = n_cloud = 0;
      node.publish_scan(nodes, rclcpp::Time(123456789 + s), 0.1 + 0.001 * s);
      if (n_scan != 1 || n_cloud != 1) return fail("patched node did not publish both messages");
      if (cloud.point_step != 16 || cloud.height != 1 || cloud.fields.size() != 4 || cloud.data.size() != cloud.width * 16u ||
          cloud.header.frame_id != "laser_frame" || cloud.header.stamp.nanosec != scan.header.stamp.nanosec)
        return fail("PointCloud2 layout");
      put(out, static_cast<int32_t>(s));
      put(out, static_cast<int32_t>(cfg));
      put(out, static_cast<uint32_t>(scan.ranges.size()));
      const float hdr[7] = {scan.angle_min, scan.angle_max, scan.angle_increment, scan.time_increment, scan.scan_time,
                            scan.range_min, scan.range_max};
      out.write(reinterpret_cast<const char*>(hdr), sizeof(hdr));
      out.write(reinterpret_cast<const char*>(scan.ranges.data()), scan.ranges.size() * 4);
      out.write(reinterpret_cast<const char*>(scan.intensities.data()), scan.intensities.size() * 4);
      put(out, static_cast<uint32_t>(cloud.width));
      out.write(reinterpret_cast<const char*>(cloud.data.data()), cloud.data.size());
      ++written;
    }
  }
  node.driver_.release();
  ros_stub::laserscan_sink() = nullptr;
  ros_stub::pointcloud2_sink() = nullptr;
  std::printf("OK gpu: %d message pairs written\n", written);
  return 0;
}
