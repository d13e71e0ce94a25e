// unpacker_mirror_test.cpp -- the C++ mirror of the SDK's sample-data unpacker seam
// (rplidar_ros2_driver_b200/host/sample_data_unpacker_b200.hpp) against vectors captured from the real SDK.
//   unpacker_mirror_test cpu                                   no GPU: the constructor must refuse
//   unpacker_mirror_test gpu <ans> <batch> <wire.bin> <rx.bin> <nodes.bin> <ts.bin> <events.bin> <timing.bin>
// wire = whole capsules, rx = u64 per capsule, nodes = 8 B each, ts = u64 per node, events = u32[n][3]
// (kind 1 = scan reset / 2 = error, nodes decoded so far, error code), timing = u32[4].
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "sample_data_unpacker_b200.hpp"

namespace {
std::vector<char> slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
int fail(const char* what) {
  std::printf("FAIL: %s\n", what);
  return 1;
}
struct Capture : rplidar_b200::SampleDataListener {
  std::vector<rpl_node_hq> nodes;
  std::vector<uint64_t> ts;
  std::vector<uint32_t> events;  // [n][3]
  void onHQNodeScanResetReq() override { push(1, 0); }
  void onHQNodeDecoded(uint64_t t, const sl_lidar_response_measurement_node_hq_t* n) override {
    nodes.push_back(*n);
    ts.push_back(t);
  }
  void onDecodingError(int err, uint8_t, const void*, size_t) override { push(2, static_cast<uint32_t>(err)); }
  void push(uint32_t kind, uint32_t code) {
    events.push_back(kind);
    events.push_back(static_cast<uint32_t>(nodes.size()));
    events.push_back(code);
  }
};
}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) return fail("usage");
  const std::string mode = argv[1];
  Capture cap;
  if (mode == "cpu") {
    bool threw = false;
    try {
      rplidar_b200::GpuSampleDataUnpacker u(cap);
    } catch (const std::exception&) {
      threw = true;
    }
    std::printf("OK cpu (unpacker without a device %s)\n", threw ? "refused: no CPU decoder" : "constructed");
    return 0;
  }
  if (argc < 10) return fail("usage gpu");
  const uint8_t ans = static_cast<uint8_t>(std::strtoul(argv[2], nullptr, 0));
  const uint32_t batch = static_cast<uint32_t>(std::strtoul(argv[3], nullptr, 0));
  const std::vector<char> wire = slurp(argv[4]), rx = slurp(argv[5]), nodes = slurp(argv[6]), ts = slurp(argv[7]),
                          events = slurp(argv[8]), timing = slurp(argv[9]);
  const uint32_t cb = rpl_capsule_bytes(ans);
  if (cb == 0 || wire.size() % cb != 0 || rx.size() != wire.size() / cb * 8 || timing.size() != 16) return fail("inputs");
  rplidar_b200::GpuSampleDataUnpacker u(cap, 0, batch);
  rpl_timing t;
  std::memcpy(&t, timing.data(), 16);
  u.enable();
  u.updateTiming(t);
  const uint64_t* rxp = reinterpret_cast<const uint64_t*>(rx.data());
  for (size_t j = 0; j < wire.size() / cb; ++j)  // one capsule per call, as the protocol codec delivers them
    if (!u.onSampleData(ans, wire.data() + j * cb, cb, rxp[j])) return fail("onSampleData");
  if (!u.flush()) return fail(u.last_error().c_str());
  if (cap.nodes.size() * 8 != nodes.size() || std::memcmp(cap.nodes.data(), nodes.data(), nodes.size()) != 0)
    return fail("nodes differ from the SDK's");
  if (cap.ts.size() * 8 != ts.size() || std::memcmp(cap.ts.data(), ts.data(), ts.size()) != 0)
    return fail("timestamps differ from the SDK's");
  if (cap.events.size() * 4 != events.size() || std::memcmp(cap.events.data(), events.data(), events.size()) != 0)
    return fail("event sequence differs from the SDK's");
  std::printf("OK gpu ans 0x%02x batch %u: %zu nodes, %zu events\n", ans, batch, cap.nodes.size(), cap.events.size() / 3);
  return 0;
}
