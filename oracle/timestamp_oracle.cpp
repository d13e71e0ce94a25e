// oracle/timestamp_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
//
// the per-sample timestamps the SDK's unpackers attach to every decoded node
// (receive time of a capsule minus a per-format delay model) and the scan-begin timestamp the scan
// holder keeps.  Restated from (reference src/sdk/src/dataunpacker/unpacker/):
//   _getSampleDelayOffsetInLegacyMode      handler_normalnode.cpp:49-68     (0x81)
//   _getSampleDelayOffsetInExpressMode     handler_capsules.cpp:55-76       (0x82, base = previous capsule's rx time :261,:264)
//   _getSampleDelayOffsetInHQMode          handler_hqnode.cpp:53-72         (0x83)
//   _getSampleDelayOffsetInUltraBoostMode  handler_capsules.cpp:272-293     (0x84, base = previous capsule's rx time)
//   _getSampleDelayOffsetInDenseMode       handler_capsules.cpp:586-607     (0x85, base = releasing capsule's rx time :739,:786)
//   _getSampleDelayOffsetInUltraDenseMode  handler_capsules.cpp:795-816     (0x86, base = releasing capsule's rx time)
//   ScanDataHolder::_scan_begin_timestamp_uS   src/sdk/src/sl_lidar_driver.cpp:293,326-328
// PARITY PINNED: tests/test_timestamps_vs_ref.py runs the SDK's own unpackers with a settable clock
// (oracle/_ref/libref_clock.so: the SDK minus its timer.cpp) and the real ScanDataHolder.
#include "oracle.h"

extern "C" uint64_t orc_sample_delay_us(uint32_t ans_type, const uint32_t* timing4, uint32_t sample_idx) {
  const uint64_t sd = timing4[0], baud_in = timing4[1], linkage = timing4[2];
  const bool ethernet = timing4[3] == 1u;  // LIDARInterfaceType::LIDAR_INTERFACE_ETHERNET
  uint64_t def_baud = 115200, size = 5;
  int64_t group = -1;
  switch (ans_type) {
    case ORC_ANS_NORMAL: def_baud = 115200; size = 5; break;
    case ORC_ANS_EXPRESS: def_baud = 115200; size = 84; group = 31; break;
    case ORC_ANS_HQ: def_baud = 1000000; size = 8; break;
    case ORC_ANS_ULTRA: def_baud = 256000; size = 132; group = 95; break;
    case ORC_ANS_DENSE: def_baud = 256000; size = 84; group = 39; break;
    case ORC_ANS_ULTRA_DENSE: def_baud = 1000000; size = 170; group = 63; break;
    default: return 0;
  }
  const uint64_t baud = baud_in ? baud_in : def_baud;
  uint64_t tx = 1000000ull * size * 10 / baud;
  if (ethernet) tx = 100;
  uint64_t d = sd + (sd >> 1) + tx + linkage;
  if (group >= 0) d += static_cast<uint64_t>(group - static_cast<int64_t>(sample_idx)) * sd;
  return d;
}

// Capsule formats: ts_out[capsule_node_offset[j] + pos] for every capsule j that released nodes.
extern "C" void orc_node_timestamps(uint32_t ans_type, const uint32_t* timing4, const uint64_t* capsule_rx_us,
                                    const uint32_t* capsule_status, const uint32_t* capsule_node_offset,
                                    uint32_t n_capsules, uint64_t* ts_out) {
  const uint32_t per = orc_capsule_nodes(ans_type);
  const bool prev_base = ans_type == ORC_ANS_EXPRESS || ans_type == ORC_ANS_ULTRA;
  for (uint32_t j = 0; j < n_capsules; ++j) {
    if (!(capsule_status[j] & ORC_CAPSULE_EMIT)) continue;
    const uint64_t base = prev_base ? capsule_rx_us[j - 1] : capsule_rx_us[j];
    for (uint32_t pos = 0; pos < per; ++pos)
      ts_out[capsule_node_offset[j] + pos] = base - orc_sample_delay_us(ans_type, timing4, pos);
  }
}

// Standard nodes: the record completing at byte node_end[i] is stamped with the rx time of the piece
// (chunk_bytes each) that byte arrived in.
extern "C" void orc_normal_timestamps(const uint32_t* timing4, const uint32_t* node_end, uint32_t n_nodes,
                                      uint32_t chunk_bytes, const uint64_t* chunk_rx_us, uint64_t* ts_out) {
  const uint64_t d = orc_sample_delay_us(ORC_ANS_NORMAL, timing4, 0);
  for (uint32_t i = 0; i < n_nodes; ++i) ts_out[i] = chunk_rx_us[node_end[i] / chunk_bytes] - d;
}
