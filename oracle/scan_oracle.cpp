// oracle/scan_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
//
// CPU restatement of the reference's per-scan hot path, written from the behaviour of
//   * ascendScanData_<sl_lidar_response_measurement_node_hq_t>
//       reference src/sdk/src/sl_lidar_driver.cpp:128-184  (accessors :102-126)
//   * RPlidarNode::publish_scan (compute body, ROS types replaced by plain arrays)
//       reference src/rplidar_node.cpp:556-680
//   * RealLidarDriver::grab_scan_data glue (ascend only when the profile asks)
//       reference src/lidar_driver_wrapper.cpp:307-342
//   * DummyLidarDriver::grab_scan_data generator
//       reference src/lidar_driver_wrapper.cpp:441-471
//
// PARITY PINNED: tests/test_oracle_vs_ref.py checks orc_ascend_scan against the
// reference's own compiled ascendScanData (oracle/_ref/libref_rplidar.so) node-for-node and
// orc_publish_scan against the reference's REAL RPlidarNode::publish_scan, bit for bit
// (oracle/_ref/libref_node.so: src/rplidar_node.cpp compiled in place against the ROS API stubs
// in oracle/ros_stubs/, since rclcpp is not installed here); tests/test_oracle_golden.py checks
// everything against tests/golden/ fixtures, which the same test file confirms to be what the
// compiled reference produces.
//
// Build: g++ -std=c++17 -O2 -ffp-contract=off (x86-64 baseline: mul and add are rounded
// separately, as in the reference build).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "oracle.h"

namespace {

// ---- fixed-point accessors (reference sl_lidar_driver.cpp:102-120) -----------------
inline float angle_deg_of(const orc_node_hq& n) { return n.angle_z_q14 * 90.f / 16384.f; }
inline void put_angle_deg(orc_node_hq& n, float deg) {
  // float -> u32 -> u16 (the struct field truncates modulo 2^16; 360.0 -> 0)
  n.angle_z_q14 = static_cast<uint16_t>(static_cast<uint32_t>(deg * 16384.f / 90.f));
}
inline bool node_before(const orc_node_hq& a, const orc_node_hq& b) {
  return angle_deg_of(a) < angle_deg_of(b);
}

uint32_t ascend_impl(orc_node_hq* buf, size_t count, bool stable) {
  const float step = 360.f / count;  // :130 (size_t -> float)
  size_t first = 0;
  // head (:133-147): first measured node, then walk back re-quantising every step
  while (first < count && buf[first].dist_mm_q2 == 0) ++first;
  if (first == count) return ORC_RESULT_OPERATION_FAIL;  // :150, buffer untouched
  for (size_t j = first; j > 0; --j) {
    float a = angle_deg_of(buf[j]) - step;
    if (a < 0.0f) a = 0.0f;
    put_angle_deg(buf[j - 1], a);
  }
  // tail (:153-168): last measured node, walk forward
  size_t last = count - 1;
  while (buf[last].dist_mm_q2 == 0) --last;  // terminates: `first` is measured
  for (size_t j = last; j + 1 < count; ++j) {
    float a = angle_deg_of(buf[j]) + step;
    if (a > 360.0f) a -= 360.0f;
    put_angle_deg(buf[j + 1], a);
  }
  // fill (:170-178): every unmeasured node i>=1 from node 0's angle
  const float front = angle_deg_of(buf[0]);
  for (size_t i = 1; i < count; ++i) {
    if (buf[i].dist_mm_q2 != 0) continue;
    float a = front + i * step;  // mul, then add: two roundings (-ffp-contract=off)
    if (a > 360.0f) a -= 360.0f;
    put_angle_deg(buf[i], a);
  }
  // :181
  if (stable)
    std::stable_sort(buf, buf + count, node_before);
  else
    std::sort(buf, buf + count, node_before);
  return ORC_RESULT_OK;
}

struct PolarPoint {  // reference rplidar_node.cpp:566-570
  float angle_rad;
  float dist_m;
  float intensity;
};

uint32_t publish_impl(const orc_node_hq* nodes, size_t count, const orc_scan_params& prm,
                      bool stable, float* ranges, float* intensities, orc_scan_header* hdr) {
  orc_scan_header h;
  std::memset(&h, 0, sizeof(h));
  if (hdr) *hdr = h;
  if (count == 0) return 0;  // :559

  std::vector<PolarPoint> pts;
  pts.reserve(count);
  const bool new_proto = prm.is_new_protocol != 0;
  for (size_t i = 0; i < count; ++i) {  // :581-600
    const orc_node_hq& n = nodes[i];
    if (n.dist_mm_q2 == 0) continue;
    float deg = n.angle_z_q14 * 90.0f / 16384.0f;
    float rad = deg * (M_PI / 180.0f);  // double product, rounded to float
    float dist = n.dist_mm_q2 / 4000.0f;
    float inten = new_proto ? static_cast<float>(n.quality) : static_cast<float>(n.quality >> 2);
    if (rad < 0.0f) rad += 2.0f * M_PI;
    if (rad >= 2.0f * M_PI) rad -= 2.0f * M_PI;
    pts.push_back({rad, dist, inten});
  }
  auto by_angle = [](const PolarPoint& a, const PolarPoint& b) { return a.angle_rad < b.angle_rad; };
  if (stable)
    std::stable_sort(pts.begin(), pts.end(), by_angle);
  else
    std::sort(pts.begin(), pts.end(), by_angle);  // :605-607
  if (pts.empty()) return 0;  // :609-611

  h.angle_min = 0.0f;  // :621-625
  h.angle_max = 2.0f * M_PI;
  h.range_min = 0.15f;
  h.range_max = prm.range_max;
  h.scan_time = prm.scan_duration;
  h.published = 1;

  const size_t m = pts.size();
  if (prm.scan_processing) {  // Mode A :630-660
    h.beam_count = static_cast<uint32_t>(m);
    h.angle_increment = static_cast<float>((2.0 * M_PI) / static_cast<double>(m));
    h.time_increment = static_cast<float>(prm.scan_duration / static_cast<double>(m));
    for (size_t b = 0; b < m; ++b) {
      ranges[b] = std::numeric_limits<float>::infinity();
      intensities[b] = 0.0f;
    }
    for (const PolarPoint& p : pts) {
      float a = p.angle_rad;
      if (prm.inverted) {
        a = (2.0f * M_PI) - a;
        if (a >= 2.0f * M_PI) a -= 2.0f * M_PI;
      }
      int idx = static_cast<int>((a - h.angle_min) / h.angle_increment);
      if (idx >= 0 && idx < static_cast<int>(m)) {
        if (p.dist_m < ranges[idx]) {
          ranges[idx] = p.dist_m;
          intensities[idx] = p.intensity;
        }
      }
    }
  } else {  // Mode B :661-677
    const double denom = static_cast<double>(m > 1 ? m - 1 : 1);
    h.beam_count = static_cast<uint32_t>(m);
    h.angle_increment = static_cast<float>((2.0 * M_PI) / denom);
    h.time_increment = static_cast<float>(prm.scan_duration / denom);
    for (size_t i = 0; i < m; ++i) {
      size_t o = prm.inverted ? (m - 1 - i) : i;
      ranges[o] = pts[i].dist_m;
      intensities[o] = pts[i].intensity;
    }
  }
  if (hdr) *hdr = h;
  return 1;
}

// ---- counter-based splitmix64 (shared definition with the CUDA
// generator in rplidar_ros2_driver_b200/csrc/synth.cuh) ---------------------------------
inline uint64_t mix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t scan_seed(uint64_t scan_id) { return mix64(0x5EED0000ull + scan_id); }
inline uint64_t draw(uint64_t seed, uint64_t i, uint64_t field) { return mix64(seed + i * 8 + field); }

// pseudo-random permutation of [0,n): 4-round Feistel on ceil(log2 n) bits + cycle walking
inline uint32_t feistel_perm(uint64_t seed, uint32_t x, uint32_t n) {
  uint32_t bits = 2;
  while ((1ull << bits) < n) ++bits;
  if (bits & 1) ++bits;
  const uint32_t half = bits / 2, mask = (1u << half) - 1;
  uint32_t v = x;
  do {
    uint32_t l = v >> half, r = v & mask;
    for (uint32_t round = 0; round < 4; ++round) {
      uint32_t f = static_cast<uint32_t>(mix64(seed + 0x1000000ull * (round + 1) + r)) & mask;
      uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    v = (l << half) | r;
  } while (v >= n);
  return v;
}

}  // namespace

extern "C" {

uint32_t orc_ascend_scan(orc_node_hq* nodes, size_t count, int stable) {
  return ascend_impl(nodes, count, stable != 0);
}

uint32_t orc_publish_scan(const orc_node_hq* nodes, size_t count, const orc_scan_params* p,
                          int stable, float* ranges, float* intensities, orc_scan_header* hdr) {
  return publish_impl(nodes, count, *p, stable != 0, ranges, intensities, hdr);
}

double orc_pipeline_batch(orc_node_hq* nodes, const uint32_t* counts, uint32_t n_scans,
                          uint32_t stride, const orc_scan_params* p, int stable, float* ranges,
                          float* intensities, uint32_t* beam_counts, float* angle_increment,
                          uint32_t* status, int threads) {
  if (threads < 1) threads = 1;
  std::atomic<uint32_t> next{0};
  auto work = [&]() {
    for (;;) {
      uint32_t s = next.fetch_add(1);
      if (s >= n_scans) break;
      orc_node_hq* sn = nodes + static_cast<size_t>(s) * stride;
      const size_t cnt = counts[s];
      uint32_t st = ORC_RESULT_OK;
      // reference lidar_driver_wrapper.cpp:328-329 (return value ignored there)
      if (p->apply_ascend) st = ascend_impl(sn, cnt, stable != 0);
      orc_scan_header h;
      publish_impl(sn, cnt, *p, stable != 0, ranges + static_cast<size_t>(s) * stride,
                   intensities + static_cast<size_t>(s) * stride, &h);
      if (beam_counts) beam_counts[s] = h.beam_count;
      if (angle_increment) angle_increment[s] = h.angle_increment;
      if (status) status[s] = st;
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  if (threads == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

void orc_dummy_scan(uint32_t call_index, orc_node_hq* out360) {
  // the reference keeps a function-static float phase, += 0.1f before every scan
  float phase = 0.0f;
  for (uint32_t c = 0; c < call_index; ++c) phase += 0.1f;
  for (int i = 0; i < 360; ++i) {
    orc_node_hq n;
    std::memset(&n, 0, sizeof(n));
    n.angle_z_q14 = static_cast<uint16_t>(static_cast<float>(i) * 16384.0f / 90.0f);
    float metres = 2.0f + 0.5f * std::sin(static_cast<float>(i) * 3.141592f / 180.0f + phase);
    n.dist_mm_q2 = static_cast<uint32_t>(metres * 1000.0f * 4.0f);
    n.quality = 200;
    out360[i] = n;
  }
}

void orc_synth_scan(uint64_t scan_id, uint32_t n, int variant, orc_node_hq* out) {
  if (n == 0) return;
  const uint64_t seed = scan_seed(scan_id);
  const uint32_t rot = static_cast<uint32_t>(draw(seed, 0xFFFFFFFFull, 7) % n);
  for (uint32_t p = 0; p < n; ++p) {
    // position p of the buffer holds revolution sample i
    uint32_t i;
    if (variant == 3)
      i = feistel_perm(seed, p, n);
    else
      i = (p + rot) % n;
    uint32_t key;
    if (variant == 2) {
      key = static_cast<uint32_t>(draw(seed, i, 0) & 0xFFFF);
    } else {
      const uint32_t base = static_cast<uint32_t>((static_cast<uint64_t>(i) << 16) / n);
      const uint32_t nxt = static_cast<uint32_t>((static_cast<uint64_t>(i + 1) << 16) / n);
      const uint32_t span = nxt - base;
      key = base + (span > 1 ? static_cast<uint32_t>(draw(seed, i, 0) % span) : 0);
    }
    const bool invalid = (draw(seed, i, 1) % 100) < 5;
    orc_node_hq nd;
    nd.angle_z_q14 = static_cast<uint16_t>(key);
    uint32_t dist = 600u + static_cast<uint32_t>(draw(seed, i, 2) % 159401ull);
    if (variant == 4) {
      // a "room": 16 angular segments of constant range (2..10 m) + 2 cm of noise
      const uint32_t seg = static_cast<uint32_t>((static_cast<uint64_t>(i) * 16) / n);
      dist = 8000u + static_cast<uint32_t>(draw(seed, seg, 4) % 32001ull) +
             static_cast<uint32_t>(draw(seed, i, 2) % 161ull) - 80u;
    }
    nd.dist_mm_q2 = invalid ? 0u : dist;
    uint8_t q = (variant == 1) ? static_cast<uint8_t>(draw(seed, i, 3) & 0xFF) : 188;
    nd.quality = invalid ? 0 : q;
    nd.flag = (p == 0) ? 1 : 2;
    out[p] = nd;
  }
}

void orc_synth_batch(uint64_t first_scan_id, uint32_t n_scans, uint32_t n, uint32_t stride,
                     int variant, orc_node_hq* out, int threads) {
  if (threads < 1) threads = 1;
  std::atomic<uint32_t> next{0};
  auto work = [&]() {
    for (;;) {
      uint32_t s = next.fetch_add(1);
      if (s >= n_scans) break;
      orc_synth_scan(first_scan_id + s, n, variant, out + static_cast<size_t>(s) * stride);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) pool.emplace_back(work);
  for (auto& t : pool) t.join();
}

}  // extern "C"
