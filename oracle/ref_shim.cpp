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
// the SDK's own sample-data unpacker (reference src/sdk/src/dataunpacker/dataunpacker.h)
#include "dataunpacker/dataunnpacker_commondef.h"
#include "dataunpacker/dataunpacker.h"

#ifdef REF_FAKE_CLOCK
// libref_clock.so: the SDK compiled WITHOUT its arch/linux/timer.cpp; the two clock functions that
// file defines are supplied here and return a value the test sets, so that the timestamps the
// unpackers attach to every node (rx time - _getSampleDelayOffsetIn*Mode) become reproducible.
#include "arch/linux/arch_linux.h"
static _u64 g_fake_now_us = 0;
namespace rp { namespace arch {
_u64 rp_getus() { return g_fake_now_us; }
_u64 rp_getms() { return g_fake_now_us / 1000; }
}}
#endif

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

// Feeds `n_bytes` of a dense-capsule answer stream (answer type 0x85, 84-byte capsules) to the
// SDK's LIDARSampleDataUnpacker exactly as SlamtecLidarDriver::onProtocolMessageDecoded does
// (reference src/sdk/src/sl_lidar_driver.cpp:1655-1662), `chunk` bytes per onSampleData call,
// and records what the listener sees: decoded HQ nodes (handler_capsules.cpp:736-791), scan
// reset requests and decoding errors, each event tagged with the number of nodes decoded
// before it.  events: [n_events][3] = {kind (1 = scan reset, 2 = error), node_count, error code}.
// NOTE the reference keeps `lastNodeSyncBit` in a function-static: it survives across calls.
// ref_unpack: the same for any measurement answer type (0x81..0x86).
int ref_unpack(uint32_t ans_type, const uint8_t* bytes, size_t n_bytes, size_t chunk, uint32_t sample_duration_us,
               void* nodes_out, size_t cap_nodes, uint32_t* n_nodes, uint32_t* events, size_t cap_events,
               uint32_t* n_events) {
  using namespace sl::internal;
  struct Capture : public LIDARSampleDataListener {
    sl_lidar_response_measurement_node_hq_t* out;
    size_t cap, n = 0;
    uint32_t* ev;
    size_t ecap, ne = 0;
    bool overflow = false;
    void push(uint32_t kind, uint32_t code) {
      if (ne < ecap) {
        ev[3 * ne] = kind;
        ev[3 * ne + 1] = static_cast<uint32_t>(n);
        ev[3 * ne + 2] = code;
        ++ne;
      } else {
        overflow = true;
      }
    }
    void onHQNodeScanResetReq() override { push(1, 0); }
    void onHQNodeDecoded(_u64, const rplidar_response_measurement_node_hq_t* node) override {
      if (n < cap) out[n++] = *node;
      else overflow = true;
    }
    void onDecodingError(int err, _u8, const void*, size_t) override { push(2, static_cast<uint32_t>(err)); }
  } cap;
  cap.out = static_cast<sl_lidar_response_measurement_node_hq_t*>(nodes_out);
  cap.cap = cap_nodes;
  cap.ev = events;
  cap.ecap = cap_events;
  LIDARSampleDataUnpacker* up = LIDARSampleDataUnpacker::CreateInstance(cap);
  if (!up) return -1;
  sl::SlamtecLidarTimingDesc timing{};
  timing.sample_duration_uS = sample_duration_us;
  timing.native_baudrate = 1000000;
  timing.linkage_delay_uS = 0;
  up->updateUnpackerContext(LIDARSampleDataUnpacker::UNPACKER_CONTEXT_TYPE_LIDAR_TIMING, &timing, sizeof(timing));
  up->enable();
  if (chunk == 0) chunk = n_bytes ? n_bytes : 1;
  for (size_t off = 0; off < n_bytes; off += chunk) {
    const size_t len = (n_bytes - off < chunk) ? (n_bytes - off) : chunk;
    up->onSampleData(static_cast<_u8>(ans_type), bytes + off, len);
  }
  LIDARSampleDataUnpacker::ReleaseInstance(up);
  *n_nodes = static_cast<uint32_t>(cap.n);
  *n_events = static_cast<uint32_t>(cap.ne);
  return cap.overflow ? 1 : 0;
}

int ref_dense_decode(const uint8_t* bytes, size_t n_bytes, size_t chunk, uint32_t sample_duration_us,
                     void* nodes_out, size_t cap_nodes, uint32_t* n_nodes, uint32_t* events, size_t cap_events,
                     uint32_t* n_events) {
  return ref_unpack(SL_LIDAR_ANS_TYPE_MEASUREMENT_DENSE_CAPSULED, bytes, n_bytes, chunk, sample_duration_us,
                    nodes_out, cap_nodes, n_nodes, events, cap_events, n_events);
}

#ifdef REF_FAKE_CLOCK
// Like ref_unpack, but the stream is fed one `chunk`-byte piece at a time with the fake clock set to
// rx_us[k] before piece k, and the timestamp of every decoded node is recorded.
// timing: {sample_duration_uS, native_baudrate, linkage_delay_uS, native_interface_type}.
int ref_unpack_ts(uint32_t ans_type, const uint8_t* bytes, size_t n_bytes, size_t chunk, const uint64_t* rx_us,
                  const uint32_t* timing4, void* nodes_out, uint64_t* ts_out, size_t cap_nodes, uint32_t* n_nodes) {
  using namespace sl::internal;
  struct Capture : public LIDARSampleDataListener {
    sl_lidar_response_measurement_node_hq_t* out;
    uint64_t* ts;
    size_t cap, n = 0;
    bool overflow = false;
    void onHQNodeScanResetReq() override {}
    void onHQNodeDecoded(_u64 t, const rplidar_response_measurement_node_hq_t* node) override {
      if (n < cap) {
        out[n] = *node;
        ts[n++] = t;
      } else {
        overflow = true;
      }
    }
    void onDecodingError(int, _u8, const void*, size_t) override {}
  } cap;
  cap.out = static_cast<sl_lidar_response_measurement_node_hq_t*>(nodes_out);
  cap.ts = ts_out;
  cap.cap = cap_nodes;
  LIDARSampleDataUnpacker* up = LIDARSampleDataUnpacker::CreateInstance(cap);
  if (!up) return -1;
  sl::SlamtecLidarTimingDesc timing{};
  timing.sample_duration_uS = timing4[0];
  timing.native_baudrate = timing4[1];
  timing.linkage_delay_uS = timing4[2];
  timing.native_interface_type = static_cast<sl::LIDARInterfaceType>(timing4[3]);
  up->updateUnpackerContext(LIDARSampleDataUnpacker::UNPACKER_CONTEXT_TYPE_LIDAR_TIMING, &timing, sizeof(timing));
  up->enable();
  size_t k = 0;
  for (size_t off = 0; off < n_bytes; off += chunk, ++k) {
    const size_t len = (n_bytes - off < chunk) ? (n_bytes - off) : chunk;
    g_fake_now_us = rx_us[k];
    up->onSampleData(static_cast<_u8>(ans_type), bytes + off, len);
  }
  LIDARSampleDataUnpacker::ReleaseInstance(up);
  *n_nodes = static_cast<uint32_t>(cap.n);
  return cap.overflow ? 1 : 0;
}
#endif

size_t ref_sizeof_node(void) { return sizeof(sl_lidar_response_measurement_node_hq_t); }

}  // extern "C"
