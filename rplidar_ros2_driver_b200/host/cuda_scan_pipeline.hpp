// cuda_scan_pipeline.hpp -- C++17 owner of an rpl_ctx: what RealLidarDriver::grab_scan_data and
// RPlidarNode::publish_scan call instead of the SDK's ascendScanData and the CPU loops.
//
// Replaces (reference):
//   drv_->ascendScanData(buf, count)             src/lidar_driver_wrapper.cpp:328-329
//   the compute body of publish_scan             src/rplidar_node.cpp:581-677
// Error conventions follow the reference: bool at the wrapper level, sl_result below, never
// throws after construction (construction throws std::runtime_error without a CUDA device:
// there is no CPU fallback).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "sdk_types.hpp"

namespace rplidar_b200 {

// The arrays and scalars publish_scan derives from one scan (everything except stamp/frame_id).
struct LaserScanData {
  float angle_min = 0.0f;        // reference rplidar_node.cpp:621
  float angle_max = 0.0f;        // :622  (float)(2*pi)
  float angle_increment = 0.0f;  // :633 / :664
  float time_increment = 0.0f;   // :635 / :667
  float scan_time = 0.0f;        // :625
  float range_min = 0.15f;       // :623
  float range_max = 0.0f;        // :624  cached_current_max_range_
  std::vector<float> ranges;
  std::vector<float> intensities;
  bool publish = false;          // false when the reference returns early (:558, :609)
};

class CudaScanPipeline {
 public:
  explicit CudaScanPipeline(int device = 0, uint32_t max_nodes = 8192, uint32_t max_scans = 1) {
    if (rpl_ctx_create(device, max_nodes, max_scans, &ctx_) != RPL_RESULT_OK || !ctx_)
      throw std::runtime_error("rpl_ctx_create failed: librplidar_b200 needs a B200 (no CPU fallback)");
  }
  ~CudaScanPipeline() { rpl_ctx_destroy(ctx_); }
  CudaScanPipeline(const CudaScanPipeline&) = delete;
  CudaScanPipeline& operator=(const CudaScanPipeline&) = delete;

  // ILidarDriver::ascendScanData semantics (in place; SL_RESULT_OPERATION_FAIL = 0x80008001
  // when no node is measured, buffer untouched).
  sl_result ascend(sl_lidar_response_measurement_node_hq_t* nodes, size_t count) {
    return rpl_ascend_scan(ctx_, reinterpret_cast<rpl_node_hq*>(nodes), count);
  }

  // publish_scan's compute body.  `scan_duration` and `range_max` only feed header scalars,
  // exactly as in the reference (range_min/range_max are never applied to ranges[]).
  bool laserscan(const std::vector<sl_lidar_response_measurement_node_hq_t>& nodes, bool is_new_protocol,
                 bool scan_processing, bool inverted, double scan_duration, float range_max,
                 LaserScanData& out) {
    out = LaserScanData{};
    if (nodes.empty()) return true;  // reference :558
    rpl_scan_params p{};
    p.is_new_protocol = is_new_protocol;
    p.scan_processing = scan_processing;
    p.inverted = inverted;
    out.ranges.resize(nodes.size());
    out.intensities.resize(nodes.size());
    uint32_t beams = 0;
    float inc = 0.0f;
    if (rpl_laserscan(ctx_, reinterpret_cast<const rpl_node_hq*>(nodes.data()), nodes.size(), &p,
                      out.ranges.data(), out.intensities.data(), &beams, &inc) != RPL_RESULT_OK)
      return false;
    out.ranges.resize(beams);
    out.intensities.resize(beams);
    if (beams == 0) return true;  // reference :609-611
    const double two_pi = 2.0 * 3.14159265358979323846;
    const double denom = scan_processing ? static_cast<double>(beams)
                                         : static_cast<double>(beams > 1 ? beams - 1 : 1);
    out.angle_min = 0.0f;
    out.angle_max = static_cast<float>(two_pi);
    out.range_min = 0.15f;
    out.range_max = range_max;
    out.scan_time = static_cast<float>(scan_duration);
    out.angle_increment = inc;
    out.time_increment = static_cast<float>(scan_duration / denom);
    out.publish = true;
    return true;
  }

  // grab glue + publish in one device round trip: ascends `nodes` in place when asked.
  bool scan(std::vector<sl_lidar_response_measurement_node_hq_t>& nodes, bool apply_ascend,
            bool is_new_protocol, bool scan_processing, bool inverted, double scan_duration,
            float range_max, LaserScanData& out, sl_result* ascend_status = nullptr) {
    out = LaserScanData{};
    if (nodes.empty()) return true;
    rpl_scan_params p{};
    p.is_new_protocol = is_new_protocol;
    p.scan_processing = scan_processing;
    p.inverted = inverted;
    p.apply_ascend = apply_ascend;
    out.ranges.resize(nodes.size());
    out.intensities.resize(nodes.size());
    uint32_t beams = 0;
    float inc = 0.0f;
    rpl_result st = 0;
    if (rpl_scan(ctx_, reinterpret_cast<rpl_node_hq*>(nodes.data()), nodes.size(), &p, out.ranges.data(),
                 out.intensities.data(), &beams, &inc, &st) != RPL_RESULT_OK)
      return false;
    if (ascend_status) *ascend_status = st;
    out.ranges.resize(beams);
    out.intensities.resize(beams);
    if (beams == 0) return true;
    const double two_pi = 2.0 * 3.14159265358979323846;
    const double denom = scan_processing ? static_cast<double>(beams)
                                         : static_cast<double>(beams > 1 ? beams - 1 : 1);
    out.angle_max = static_cast<float>(two_pi);
    out.range_max = range_max;
    out.scan_time = static_cast<float>(scan_duration);
    out.angle_increment = inc;
    out.time_increment = static_cast<float>(scan_duration / denom);
    out.publish = true;
    return true;
  }

  // PointCloud2 extension (north star; no reference line): one revolution -> [n][x, y, z, intensity] through
  // rpl_cloud_batch (window, polar->xyz, optional SOR and voxel grid; oracle/cloud_oracle.cpp is the definition).
  bool cloud(const std::vector<sl_lidar_response_measurement_node_hq_t>& nodes, const rpl_cloud_params& params,
             std::vector<float>& xyzi, uint32_t& n_points) {
    n_points = 0;
    xyzi.clear();
    if (nodes.empty()) return true;
    // rpl_cloud_batch moves [n_scans][stride] blocks; an even stride keeps the scan on the bulk-copy path
    const uint32_t n = static_cast<uint32_t>(nodes.size()), stride = (n + 1u) & ~1u;
    staging_.assign(stride, rpl_node_hq{});
    std::memcpy(staging_.data(), nodes.data(), nodes.size() * sizeof(rpl_node_hq));
    xyzi.resize(static_cast<size_t>(stride) * 4);
    uint32_t count = n, pts = 0;
    if (rpl_cloud_batch(ctx_, staging_.data(), &count, 1, stride, &params, xyzi.data(), &pts) != RPL_RESULT_OK) return false;
    n_points = pts;
    xyzi.resize(static_cast<size_t>(pts) * 4);
    return true;
  }

  const char* last_error() const { return rpl_last_error(ctx_); }
  rpl_ctx* raw() { return ctx_; }

 private:
  rpl_ctx* ctx_ = nullptr;
  std::vector<rpl_node_hq> staging_;
};

}  // namespace rplidar_b200
