// oracle/ref_shim_node.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the reference's REAL RPlidarNode::publish_scan (src/rplidar_node.cpp:556-680).  ROS 2 is absent in
// this image, so the node's translation unit is compiled in place (an #include of the file where it lies
// under /root/reference; nothing is copied) against the API stubs in oracle/ros_stubs/ -- types and
// do-nothing bodies only -- and the one call that matters, scan_pub_->publish(scan_msg), hands the
// message to the capture hook below.  Everything publish_scan computes (filter, unpack, std::sort, Mode A
// / Mode B fill, header scalars) is the reference's own compiled code.  `private` is opened for this
// translation unit only, to reach publish_scan and the three members it reads.
#include <sstream>  // before the keyword games: the standard headers must see the real `private`

#include "ros_stub_core.hpp"
#define private public
#define protected public
#include "rplidar_node.cpp"  // -I /root/reference/src -I /root/reference/include
#undef private
#undef protected

#include <malloc.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "sl_lidar_driver.h"

namespace {
// One worker's private copy of everything the path touches: a node (constructed against the stubs), the two
// driver flavours publish_scan distinguishes, and an SDK driver object for ascendScanData.
struct Worker {
  RPlidarNode node;
  std::unique_ptr<LidarDriverInterface> real{new RealLidarDriver()};
  std::unique_ptr<LidarDriverInterface> dummy{new DummyLidarDriver()};
  std::vector<sl_lidar_response_measurement_node_hq_t> raw, nodes;
  // the SDK driver object the wrapper itself holds and calls ascendScanData on (lidar_driver_wrapper.cpp:78,329)
  sl::ILidarDriver* sdk() { return static_cast<RealLidarDriver*>(real.get())->drv_; }
  void configure(int is_new_protocol, int scan_processing, int inverted, float max_range) {
    static_cast<RealLidarDriver*>(real.get())->profile_.protocol =
        is_new_protocol ? ProtocolType::NEW_TYPE : ProtocolType::OLD_TYPE;
    node.driver_.release();
    node.driver_.reset(is_new_protocol ? real.get() : dummy.get());
    node.params_.scan_processing = scan_processing != 0;
    node.params_.inverted = inverted != 0;
    node.params_.frame_id = "laser_frame";
    node.cached_current_max_range_ = max_range;
    node.scan_pub_ = std::make_shared<rclcpp_lifecycle::LifecyclePublisher<sensor_msgs::msg::LaserScan>>();
  }
  ~Worker() { node.driver_.release(); }
};
}  // namespace

extern "C" {

// The reference's per-scan path on a batch, `threads` workers, one scan per task:
//   RealLidarDriver::grab_scan_data's glue (src/lidar_driver_wrapper.cpp:324-336): the SDK's buffer, ascendScanData
//   in place if asked, nodes.assign(...);  then RPlidarNode::publish_scan (src/rplidar_node.cpp:556-680).
// nodes: [n_scans][stride] (left untouched: each scan is first copied into the worker's buffer, as the SDK's
// grabScanDataHq would deliver it).  ranges / intensities / beams (nullable together): outputs for checking;
// when null the published message is dropped (timing).  Returns the wall time of the parallel section.
double ref_pipeline_batch(const void* nodes_v, const uint32_t* counts, uint32_t n_scans, uint32_t stride,
                          int is_new_protocol, int scan_processing, int inverted, int apply_ascend, float max_range,
                          double scan_duration, float* ranges, float* intensities, uint32_t* beams, int threads) {
  using Node = sl_lidar_response_measurement_node_hq_t;
  const Node* all = static_cast<const Node*>(nodes_v);
  if (threads < 1) threads = 1;
  // Harness tuning only: the path allocates and frees several 100 KB vectors per scan; glibc would mmap/munmap
  // each of them and the worker threads would queue on the process's address-space lock.  Keeping the blocks
  // on the heap lets the reference's loop scale across the host's cores as separate node processes would.
  mallopt(M_MMAP_THRESHOLD, 1 << 30);
  mallopt(M_TRIM_THRESHOLD, 1 << 30);
  static std::vector<std::unique_ptr<Worker>> workers;  // built once, reused by later calls
  while (static_cast<int>(workers.size()) < threads) workers.emplace_back(new Worker());
  for (int t = 0; t < threads; ++t) workers[t]->configure(is_new_protocol, scan_processing, inverted, max_range);
  std::atomic<uint32_t> next{0};
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t]() {
      Worker& w = *workers[t];
      uint32_t cur = 0;
      if (ranges) {
        ros_stub::laserscan_sink() = [&](const sensor_msgs::msg::LaserScan& m) {
          beams[cur] = static_cast<uint32_t>(m.ranges.size());
          std::memcpy(ranges + static_cast<size_t>(cur) * stride, m.ranges.data(), m.ranges.size() * sizeof(float));
          std::memcpy(intensities + static_cast<size_t>(cur) * stride, m.intensities.data(),
                      m.intensities.size() * sizeof(float));
        };
      }
      for (;;) {
        const uint32_t s = next.fetch_add(1);
        if (s >= n_scans) break;
        cur = s;
        const uint32_t n = counts[s];
        if (beams) beams[s] = 0;
        w.raw.assign(all + static_cast<size_t>(s) * stride, all + static_cast<size_t>(s) * stride + n);
        if (apply_ascend) w.sdk()->ascendScanData(w.raw.data(), n);
        w.nodes.assign(w.raw.begin(), w.raw.begin() + n);
        w.node.publish_scan(w.nodes, rclcpp::Time(0), scan_duration);
      }
      ros_stub::laserscan_sink() = nullptr;
    });
  }
  for (auto& th : pool) th.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// One call of RPlidarNode::publish_scan.  header7 = {angle_min, angle_max, angle_increment, time_increment,
// scan_time, range_min, range_max}.  Returns 1 if a message was published (beam_count / arrays filled), 0 if
// publish_scan returned without publishing, -1 if the output capacity is too small.
int ref_publish_scan(const void* nodes_v, size_t n, int is_new_protocol, int scan_processing, int inverted,
                     float max_range, double scan_duration, float* ranges, float* intensities, size_t capacity,
                     float* header7, uint32_t* beam_count) {
  using Node = sl_lidar_response_measurement_node_hq_t;
  static RPlidarNode* node = new RPlidarNode();  // constructed once: declares its parameters against the stubs
  static std::unique_ptr<LidarDriverInterface> real(new RealLidarDriver());
  static std::unique_ptr<LidarDriverInterface> dummy(new DummyLidarDriver());
  // is_new_protocol comes from dynamic_cast<RealLidarDriver*>(driver_) && is_new_type() (:577-581)
  RealLidarDriver* rd = static_cast<RealLidarDriver*>(real.get());
  rd->profile_.protocol = is_new_protocol ? ProtocolType::NEW_TYPE : ProtocolType::OLD_TYPE;
  node->driver_.release();
  node->driver_.reset(is_new_protocol ? real.get() : dummy.get());
  node->params_.scan_processing = scan_processing != 0;
  node->params_.inverted = inverted != 0;
  node->params_.frame_id = "laser_frame";
  node->cached_current_max_range_ = max_range;
  node->scan_pub_ = std::make_shared<rclcpp_lifecycle::LifecyclePublisher<sensor_msgs::msg::LaserScan>>();
  sensor_msgs::msg::LaserScan got;
  bool published = false;
  ros_stub::laserscan_sink() = [&](const sensor_msgs::msg::LaserScan& m) {
    got = m;
    published = true;
  };
  const Node* p = static_cast<const Node*>(nodes_v);
  std::vector<Node> v(p, p + n);
  node->publish_scan(v, rclcpp::Time(0), scan_duration);
  ros_stub::laserscan_sink() = nullptr;
  node->driver_.release();  // the two drivers are owned by the statics above
  *beam_count = 0;
  if (!published) return 0;
  if (got.ranges.size() > capacity) return -1;
  *beam_count = static_cast<uint32_t>(got.ranges.size());
  std::memcpy(ranges, got.ranges.data(), got.ranges.size() * sizeof(float));
  std::memcpy(intensities, got.intensities.data(), got.intensities.size() * sizeof(float));
  const float h[7] = {got.angle_min, got.angle_max, got.angle_increment, got.time_increment,
                      got.scan_time,  got.range_min, got.range_max};
  std::memcpy(header7, h, sizeof(h));
  return 1;
}

}  // extern "C"
