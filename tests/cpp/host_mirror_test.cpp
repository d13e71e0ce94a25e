// host_mirror_test.cpp -- exercises the C++ host mirror (rplidar_ros2_driver_b200/host).
//   host_mirror_test cpu <golden_raw.bin>                      no GPU: interface + dummy generator
//   host_mirror_test gpu <golden_raw.bin> <golden_asc.bin> <ranges.bin> <intens.bin>
// golden_raw.bin = 16 captured dummy scans (16*360*8 bytes), golden_asc.bin = the reference's
// ascendScanData of scan 1, ranges/intens = Mode A LaserScan of scan 1 (is_new_protocol = 0).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <memory>
#include <string>
#include <vector>

#include "lidar_driver_wrapper.hpp"
#include "publish_cloud_b200.hpp"
#include "publish_scan_b200.hpp"

namespace {
std::vector<char> slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
struct Header { int stamp = 0; std::string frame_id; };
struct FakeLaserScan {  // same members as sensor_msgs::msg::LaserScan
  Header header;
  float angle_min, angle_max, angle_increment, time_increment, scan_time, range_min, range_max;
  std::vector<float> ranges, intensities;
};
struct FakePointField {  // same members as sensor_msgs::msg::PointField
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 0;
};
struct FakePointCloud2 {  // same members as sensor_msgs::msg::PointCloud2
  Header header;
  uint32_t height = 0, width = 0;
  std::vector<FakePointField> fields;
  bool is_bigendian = true;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = false;
};
int fail(const char* what) { std::printf("FAIL: %s\n", what); return 1; }
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) return fail("usage");
  const std::string mode = argv[1];
  const std::vector<char> raw = slurp(argv[2]);
  if (raw.size() != 16u * 360u * 8u) return fail("golden_raw size");

  // the dummy backend reproduces the captured reference sequence byte for byte
  std::unique_ptr<LidarDriverInterface> drv = std::make_unique<DummyLidarDriver>(0);
  if (!drv->connect("/dev/null", 115200, true) || !drv->isConnected() || drv->getHealth() != 0)
    return fail("dummy driver interface");
  std::vector<sl_lidar_response_measurement_node_hq_t> nodes;
  for (int c = 0; c < 16; ++c) {
    if (!drv->grab_scan_data(nodes) || nodes.size() != 360) return fail("grab_scan_data");
    if (std::memcmp(nodes.data(), raw.data() + static_cast<size_t>(c) * 360 * 8, 360 * 8) != 0)
      return fail("dummy scan differs from the captured reference");
  }
  if (drv->get_hw_max_distance() != 40.0f) return fail("hw max distance");
  {  // PointCloud2 message from the cloud path's [n][x, y, z, intensity] points (pure host code)
    const float pts[8] = {1.0f, 2.0f, 0.0f, 47.0f, -3.0f, 4.5f, 0.0f, 12.0f};
    FakePointCloud2 pc;
    rplidar_b200::fill_pointcloud2_msg(pc, pts, 2, 7, std::string("laser_frame"));
    if (pc.header.stamp != 7 || pc.header.frame_id != "laser_frame" || pc.height != 1 || pc.width != 2 ||
        pc.fields.size() != 4 || pc.fields[3].name != "intensity" || pc.fields[3].offset != 12 ||
        pc.fields[1].datatype != 7 || pc.fields[2].count != 1 || pc.is_bigendian || pc.point_step != 16 ||
        pc.row_step != 32 || pc.data.size() != 32 || std::memcmp(pc.data.data(), pts, 32) != 0 || !pc.is_dense)
      return fail("PointCloud2 layout");
    rplidar_b200::fill_pointcloud2_msg(pc, nullptr, 0, 8, std::string());
    if (pc.width != 0 || !pc.data.empty() || pc.row_step != 0) return fail("empty PointCloud2");
  }
  if (mode == "cpu") {
    bool threw = false;
    try {
      rplidar_b200::CudaScanPipeline p(0, 8192, 1);
    } catch (const std::exception&) {
      threw = true;
    }
    std::printf("OK cpu (pipeline without a device %s)\n", threw ? "refused: no CPU fallback" : "constructed");
    return 0;
  }

  if (argc < 6) return fail("usage gpu");
  const std::vector<char> asc = slurp(argv[3]), gr = slurp(argv[4]), gi = slurp(argv[5]);
  auto pipe = std::make_shared<rplidar_b200::CudaScanPipeline>(0, 8192, 1);
  GpuDummyLidarDriver gdrv(pipe, 0);
  gdrv.connect("/dev/null", 115200, true);
  if (!gdrv.grab_scan_data(nodes)) return fail("gpu grab");
  if (asc.size() != 360 * 8 || std::memcmp(nodes.data(), asc.data(), asc.size()) != 0)
    return fail("ascended buffer differs from the reference's ascendScanData");
  rplidar_b200::LaserScanData d;
  if (!pipe->laserscan(nodes, false, true, false, 0.1, 12.0f, d) || !d.publish) return fail("laserscan");
  if (d.ranges.size() * 4 != gr.size() || std::memcmp(d.ranges.data(), gr.data(), gr.size()) != 0)
    return fail("ranges differ");
  if (std::memcmp(d.intensities.data(), gi.data(), gi.size()) != 0) return fail("intensities differ");
  FakeLaserScan msg{};
  if (!rplidar_b200::fill_laserscan_msg(msg, std::move(d), 7, std::string("laser_frame"))) return fail("fill msg");
  if (msg.ranges.size() != 360 || msg.range_min != 0.15f || msg.range_max != 12.0f || msg.header.frame_id != "laser_frame")
    return fail("message fields");
  // all-unmeasured scan: OPERATION_FAIL, nothing to publish
  std::vector<sl_lidar_response_measurement_node_hq_t> dead(5);
  std::memset(dead.data(), 0, dead.size() * 8);
  if (pipe->ascend(dead.data(), dead.size()) != 0x80008001u) return fail("ascend of an all-unmeasured scan");
  rplidar_b200::LaserScanData e;
  if (!pipe->laserscan(dead, false, false, false, 0.1, 12.0f, e) || e.publish) return fail("empty publish");
  std::printf("OK gpu\n");
  return 0;
}
