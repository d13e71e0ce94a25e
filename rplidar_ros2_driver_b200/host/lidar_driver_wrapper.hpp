// lidar_driver_wrapper.hpp -- host mirror of the reference's wrapper layer.
//
// Same seam as reference include/lidar_driver_wrapper.hpp: LidarDriverInterface (12 pure
// virtuals, :139-267), DriverProfile / ProtocolType (:77-118), DummyLidarDriver (:386-440).
// RealLidarDriver itself is device management on top of the Slamtec SDK and stays the
// reference's own code; the only change it needs is in grab_scan_data, where
// `drv_->ascendScanData(buf, count)` becomes `pipeline_.ascend(buf, count)`.
// GpuDummyLidarDriver shows that seam end to end without hardware: the reference's dummy
// generator followed by the wrapper's "ascend iff the profile asks" glue on the GPU.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "cuda_scan_pipeline.hpp"
#include "sdk_types.hpp"

enum class ProtocolType { OLD_TYPE, NEW_TYPE };

struct DriverProfile {
  std::string model_name = "Unknown";
  ProtocolType protocol = ProtocolType::OLD_TYPE;
  float hw_max_distance = 12.0f;
  std::string active_mode = "Standard";
  uint16_t active_rpm = 0;
  bool apply_geometric_correction = true;
};

class LidarDriverInterface {
 public:
  virtual ~LidarDriverInterface() = default;
  virtual bool connect(const std::string& port, sl_u32 baudrate, bool use_geometric_compensation = true) = 0;
  virtual void disconnect() = 0;
  virtual bool isConnected() = 0;
  virtual bool start_motor(std::string user_mode_pref = "", uint16_t user_rpm_pref = 0) = 0;
  virtual void stop_motor() = 0;
  virtual int getHealth() = 0;
  virtual void reset() = 0;
  virtual bool grab_scan_data(std::vector<sl_lidar_response_measurement_node_hq_t>& nodes) = 0;
  virtual void detect_and_init_strategy() = 0;
  virtual void print_summary() = 0;
  virtual float get_hw_max_distance() const = 0;
  virtual bool set_motor_speed(uint16_t rpm) = 0;
};

// The reference's fake backend (src/lidar_driver_wrapper.cpp:417-471): a 360-node ring,
// phase advancing 0.1 rad per call.  `sleep_ms` is 100 in the reference; tests pass 0.
class DummyLidarDriver : public LidarDriverInterface {
 public:
  explicit DummyLidarDriver(int sleep_ms = 100) : sleep_ms_(sleep_ms) {}
  bool connect(const std::string&, sl_u32, bool = true) override { return true; }
  void disconnect() override {}
  bool isConnected() override { return true; }
  bool start_motor(std::string = "", uint16_t = 0) override { return true; }
  void stop_motor() override {}
  int getHealth() override { return 0; }
  void reset() override {}
  bool grab_scan_data(std::vector<sl_lidar_response_measurement_node_hq_t>& nodes) override;
  void detect_and_init_strategy() override {}
  void print_summary() override;
  float get_hw_max_distance() const override { return 40.0f; }
  bool set_motor_speed(uint16_t) override { return true; }

 private:
  float phase_ = 0.0f;  // per instance (the reference keeps a function-static: not re-entrant)
  int sleep_ms_;
};

// Dummy generator + the wrapper's grab glue on the GPU (reference
// src/lidar_driver_wrapper.cpp:323-337): ascend in place iff the profile asks.
class GpuDummyLidarDriver : public DummyLidarDriver {
 public:
  explicit GpuDummyLidarDriver(std::shared_ptr<rplidar_b200::CudaScanPipeline> pipeline, int sleep_ms = 100)
      : DummyLidarDriver(sleep_ms), pipeline_(std::move(pipeline)) {}
  bool connect(const std::string&, sl_u32, bool use_geometric_compensation = true) override {
    profile_.apply_geometric_correction = use_geometric_compensation;  // reference :107
    return true;
  }
  bool grab_scan_data(std::vector<sl_lidar_response_measurement_node_hq_t>& nodes) override {
    if (!DummyLidarDriver::grab_scan_data(nodes)) return false;
    if (profile_.apply_geometric_correction) pipeline_->ascend(nodes.data(), nodes.size());  // result ignored, :329
    return true;
  }
  const DriverProfile& profile() const { return profile_; }

 private:
  std::shared_ptr<rplidar_b200::CudaScanPipeline> pipeline_;
  DriverProfile profile_;
};
