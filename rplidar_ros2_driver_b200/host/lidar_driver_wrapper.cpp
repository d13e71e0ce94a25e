// lidar_driver_wrapper.cpp -- DummyLidarDriver (fixture generator, kept bit-identical to the
// reference's: src/lidar_driver_wrapper.cpp:441-471).
#include "lidar_driver_wrapper.hpp"

#include <chrono>
#include <cmath>
#include <iostream>
#include <thread>

bool DummyLidarDriver::grab_scan_data(std::vector<sl_lidar_response_measurement_node_hq_t>& nodes) {
  constexpr int kCount = 360;
  nodes.clear();
  nodes.reserve(kCount);
  phase_ += 0.1f;
  for (int i = 0; i < kCount; ++i) {
    sl_lidar_response_measurement_node_hq_t n{};
    const float fi = static_cast<float>(i);
    n.angle_z_q14 = static_cast<sl_u16>(fi * 16384.0f / 90.0f);
    const float metres = 2.0f + 0.5f * std::sin(fi * 3.141592f / 180.0f + phase_);
    n.dist_mm_q2 = static_cast<sl_u32>(metres * 1000.0f * 4.0f);
    n.quality = 200;
    nodes.push_back(n);
  }
  if (sleep_ms_ > 0) std::this_thread::sleep_for(std::chrono::milliseconds(sleep_ms_));
  return true;
}

void DummyLidarDriver::print_summary() { std::cout << "[Dummy] Virtual RPLIDAR device ready." << std::endl; }
