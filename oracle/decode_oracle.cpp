// oracle/decode_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
//
// the step immediately BEFORE the hot path.  CPU restatement of the
// reference's dense-capsule unpacker (S2/S3 DenseBoost, answer type 0x85), written from the
// behaviour of
//   UnpackerHandler_DenseCapsuleNode::onData                  handler_capsules.cpp:639-734
//   UnpackerHandler_DenseCapsuleNode::_onScanNodeDenseCapsuleData      :736-791
//   (reference src/sdk/src/dataunpacker/unpacker/handler_capsules.cpp; capsule layout
//    src/sdk/include/sl_lidar_cmd.h:223-234: 84 bytes = 2 sync/checksum bytes, u16
//    start_angle_sync_q6, 40 x u16 distance)
// for FRAMED input: an array of 84-byte capsules whose first two bytes carry the sync nibbles
// 0xA / 0x5 (what the protocol codec hands over, one capsule per message).  A capsule whose sync
// nibbles are wrong makes the reference's byte-level state machine hunt for the next 0xA/0x5
// pair inside the payload; that resynchronisation is not modelled: such a capsule is reported as
// ORC_CAPSULE_BAD_FRAME and equivalence with the reference is only claimed for streams without
// one.  Checksum errors, scan-start capsules, the angular-jump discard rule and the
// function-static `lastNodeSyncBit` (made an explicit in/out state here) are all reproduced.
//
// PARITY PINNED: tests/test_decode_oracle_vs_ref.py compares node for node, event for event,
// with the SDK's own LIDARSampleDataUnpacker compiled in place (oracle/_ref).
#include <cstring>

#include "oracle.h"

extern "C" uint32_t orc_dense_decode(const uint8_t* capsules, uint32_t n_capsules, uint32_t sample_duration_us,
                                     uint32_t* sync_state, orc_node_hq* nodes_out, uint32_t* capsule_status,
                                     uint32_t* capsule_node_offset) {
  bool prev_ready = false;
  const uint8_t* prev = nullptr;
  int last_sync = static_cast<int>(*sync_state & 1u);
  uint32_t n_out = 0;
  for (uint32_t j = 0; j < n_capsules; ++j) {
    const uint8_t* c = capsules + static_cast<size_t>(j) * ORC_DENSE_CAPSULE_BYTES;
    uint32_t st = 0;
    if (capsule_node_offset) capsule_node_offset[j] = n_out;
    if ((c[0] >> 4) != 0xA || (c[1] >> 4) != 0x5) {  // :651-674 (framed view)
      prev_ready = false;
      if (capsule_status) capsule_status[j] = ORC_CAPSULE_BAD_FRAME;
      continue;
    }
    uint8_t sum = 0;
    for (int b = 2; b < ORC_DENSE_CAPSULE_BYTES; ++b) sum ^= c[b];  // :686-691
    const uint8_t recv = static_cast<uint8_t>((c[0] & 0xF) | (c[1] << 4));
    if (recv != sum) {  // :721-727
      prev_ready = false;
      if (capsule_status) capsule_status[j] = ORC_CAPSULE_CHECKSUM_ERR;
      continue;
    }
    st |= ORC_CAPSULE_OK;
    const uint32_t start = static_cast<uint32_t>(c[2]) | (static_cast<uint32_t>(c[3]) << 8);
    if (start & 0x8000u) {  // :706-717 first capsule of a revolution
      st |= ORC_CAPSULE_SYNC;
      if (prev_ready) st |= ORC_CAPSULE_ENCODER_RESET_ERR;
      prev_ready = false;
    }
    if (prev_ready) {  // :741-787
      const uint32_t pstart = static_cast<uint32_t>(prev[2]) | (static_cast<uint32_t>(prev[3]) << 8);
      const int cur_q8 = static_cast<int>((start & 0x7FFFu) << 2);
      const int prev_q8 = static_cast<int>((pstart & 0x7FFFu) << 2);
      int diff_q8 = cur_q8 - prev_q8;
      if (prev_q8 > cur_q8) diff_q8 += (360 << 8);
      const int thr_q8 = (360 * 100 * 40 / static_cast<int>(1000000 / sample_duration_us)) << 8;
      if (diff_q8 > thr_q8) {
        st |= ORC_CAPSULE_DISCARD;
      } else {
        const int inc_q16 = (diff_q8 << 8) / 40;
        int cur_q16 = prev_q8 << 8;
        for (int pos = 0; pos < 40; ++pos) {
          const int dist = static_cast<int>(prev[4 + 2 * pos]) | (static_cast<int>(prev[5 + 2 * pos]) << 8);
          const int dist_q2 = dist << 2;
          int angle_q6 = cur_q16 >> 10;
          int sync = (((cur_q16 + inc_q16) % (360 << 16)) < (inc_q16 << 1)) ? 1 : 0;
          sync = (sync ^ last_sync) & sync;
          cur_q16 += inc_q16;
          if (angle_q6 < 0) angle_q6 += (360 << 6);
          if (angle_q6 >= (360 << 6)) angle_q6 -= (360 << 6);
          orc_node_hq nd;
          nd.flag = static_cast<uint8_t>(sync | ((!sync) << 1));
          nd.quality = dist_q2 ? (0x2F << 2) : 0;
          nd.angle_z_q14 = static_cast<uint16_t>((angle_q6 << 8) / 90);
          nd.dist_mm_q2 = static_cast<uint32_t>(dist_q2);
          nodes_out[n_out++] = nd;
          last_sync = sync;
        }
        st |= ORC_CAPSULE_EMIT;
      }
    }
    prev = c;
    prev_ready = true;
    if (capsule_status) capsule_status[j] = st;
  }
  *sync_state = static_cast<uint32_t>(last_sync);
  return n_out;
}

// ---- scan assembly ---------------------------------------------------------
// Restatement of ScanDataHolder::pushScanNodeData + rewindCurrentScanData (reference
// src/sdk/src/sl_lidar_driver.cpp:272-315) as a pure function of the node stream and the positions
// of the scan-reset requests: a node with flag bit 0 opens a scan (publishing the one in progress,
// if it holds anything), other nodes are appended only to an open scan, a reset empties the scan in
// progress, and a scan that reaches `max_nodes` keeps overwriting its last entry.
// PARITY PINNED: tests/test_decode_oracle_vs_ref.py runs the reference's real ScanDataHolder
// (oracle/ref_shim_holder.cpp compiles sl_lidar_driver.cpp in place) on the same streams.
// node_ts / scan_ts (nullable): a published scan carries the timestamp of the scan-start node that
// opened it (:293, reported by waitAndLockAvailableScan :326-328).
extern "C" uint32_t orc_assemble_scans_ts(const orc_node_hq* nodes, uint32_t n, const uint32_t* resets,
                                          uint32_t n_resets, uint32_t max_nodes, orc_node_hq* scans_out,
                                          uint32_t scan_stride, uint32_t* scan_len, uint32_t max_scans,
                                          const uint64_t* node_ts, uint64_t* scan_ts) {
  uint64_t begin_ts = 0;
  uint32_t n_scans = 0, cur = 0, ri = 0;  // cur: nodes in the scan in progress
  orc_node_hq* slot = scans_out;           // the scan in progress is built in place
  for (uint32_t i = 0; i <= n; ++i) {
    while (ri < n_resets && resets[ri] == i) {  // rewindCurrentScanData :310-313
      cur = 0;
      ++ri;
    }
    if (i == n) break;
    const orc_node_hq& nd = nodes[i];
    if (nd.flag & 1u) {  // :279-293
      if (cur) {
        if (n_scans < max_scans) {
          scan_len[n_scans] = cur;
          if (scan_ts) scan_ts[n_scans] = begin_ts;
        }
        ++n_scans;
        slot = (n_scans < max_scans) ? scans_out + static_cast<size_t>(n_scans) * scan_stride : nullptr;
        cur = 0;
      }
      begin_ts = node_ts ? node_ts[i] : 0;
    } else if (cur == 0) {
      continue;  // :295-299 no partial scans
    }
    if (cur >= max_nodes) {  // :302-308
      if (slot) slot[cur - 1] = nd;
    } else {
      if (slot && cur < scan_stride) slot[cur] = nd;
      ++cur;
    }
  }
  return n_scans;
}

extern "C" uint32_t orc_assemble_scans(const orc_node_hq* nodes, uint32_t n, const uint32_t* resets,
                                       uint32_t n_resets, uint32_t max_nodes, orc_node_hq* scans_out,
                                       uint32_t scan_stride, uint32_t* scan_len, uint32_t max_scans) {
  return orc_assemble_scans_ts(nodes, n, resets, n_resets, max_nodes, scans_out, scan_stride, scan_len, max_scans,
                               nullptr, nullptr);
}
