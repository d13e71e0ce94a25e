// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" window onto the UNMODIFIED reference, compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/libref_rplidar.so:
//
//   ref_ascend_scan()  -> sl::ILidarDriver::ascendScanData
//                         (reference src/sdk/include/sl_lidar_driver.h:477,
//                          body src/sdk/src/sl_lidar_driver.cpp:128-184, entry :957-960)
//   ref_dummy_grab()   -> DummyLidarDriver::grab_scan_data
//                         (reference src/lidar_driver_wrapper.cpp:441-471)
//
// No reference source is copied here; this file only calls the reference's public
// symbols.  rclcpp is absent in this image, so RPlidarNode::publish_scan cannot be
// compiled -- that body is restated in oracle/scan_oracle.cpp instead.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "lidar_driver_wrapper.hpp"
#include "sl_lidar.h"
#include "sl_lidar_driver.h"

namespace {
sl::ILidarDriver* sdk_driver() {
  // createLidarDriver() needs no device: it only constructs SlamtecLidarDriver.
  static sl::ILidarDriver* drv = *sl::createLidarDriver();
  return drv;
}
}  // namespace

extern "C" {

// Returns the reference's sl_result (0 = SL_RESULT_OK, 0x80008001 = OPERATION_FAIL).
uint32_t ref_ascend_scan(void* nodes, size_t count) {
  return static_cast<uint32_t>(sdk_driver()->ascendScanData(
      static_cast<sl_lidar_response_measurement_node_hq_t*>(nodes), count));
}

// One call of the reference's dummy generator (it advances a function-static phase
// by 0.1 rad and sleeps 100 ms per call).  Returns the node count (360) or -1.
int ref_dummy_grab(void* out_nodes, size_t capacity_nodes) {
  static DummyLidarDriver dummy;
  std::vector<sl_lidar_response_measurement_node_hq_t> v;
  if (!dummy.grab_scan_data(v)) return -1;
  if (v.size() > capacity_nodes) return -1;
  std::memcpy(out_nodes, v.data(), v.size() * sizeof(v[0]));
  return static_cast<int>(v.size());
}

size_t ref_sizeof_node(void) { return sizeof(sl_lidar_response_measurement_node_hq_t); }

}  // extern "C"
