// oracle/capsule_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
//
// SURVEY.md 8(f) rank 1, the remaining measurement answer formats.  CPU restatement, written from
// the behaviour of the reference's sample-data unpackers (src/sdk/src/dataunpacker/unpacker/):
//   0x81 standard nodes       handler_normalnode.cpp:88-141      (byte-level state machine, exact)
//   0x82 express capsules     handler_capsules.cpp:109-266       (84 B -> 32 nodes)
//   0x83 HQ capsules          handler_hqnode.cpp:93-172, CRC src/sdk/src/sl_crc.cpp:52-99 (781 B -> 96)
//   0x84 ultra capsules       handler_capsules.cpp:324-580       (132 B -> 96 nodes, varbitscale)
//   0x86 ultra-dense capsules handler_capsules.cpp:852-1047      (170 B -> 64 nodes)
// (0x85 dense capsules live in decode_oracle.cpp.)  Capsule formats take FRAMED input: an array of
// fixed-size capsules as the protocol codec hands them over; a capsule whose sync marker is wrong
// is reported as ORC_CAPSULE_BAD_FRAME (the reference would hunt for the next marker byte by byte;
// that resynchronisation is not modelled, equivalence is claimed for streams without one).  The
// 5-byte standard nodes are decoded from a raw byte stream with the reference's own
// resynchronisation rules.
//
// PARITY PINNED: tests/test_capsule_oracle_vs_ref.py feeds the same bytes to the SDK's own
// LIDARSampleDataUnpacker compiled in place (oracle/_ref) and compares node for node, event for event.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

inline uint32_t rd16(const uint8_t* p) { return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8); }
inline uint32_t rd32(const uint8_t* p) { return rd16(p) | (rd16(p + 2) << 16); }

inline orc_node_hq make_node(int angle_q6, int dist_q2, int sync, int quality) {
  if (angle_q6 < 0) angle_q6 += (360 << 6);
  if (angle_q6 >= (360 << 6)) angle_q6 -= (360 << 6);
  orc_node_hq nd;
  nd.flag = static_cast<uint8_t>(sync | ((!sync) << 1));
  nd.quality = static_cast<uint8_t>(quality);
  nd.angle_z_q14 = static_cast<uint16_t>((angle_q6 << 8) / 90);
  nd.dist_mm_q2 = static_cast<uint32_t>(dist_q2);
  return nd;
}

// handler_capsules.cpp:422-458
uint32_t varbitscale_decode(uint32_t scaled, uint32_t& level) {
  static const uint32_t kScaledBase[] = {3328, 1792, 1280, 512, 0};
  static const uint32_t kLevel[] = {4, 3, 2, 1, 0};
  static const uint32_t kTargetBase[] = {1u << 14, 1u << 12, 1u << 11, 1u << 9, 0};
  for (int i = 0; i < 5; ++i) {
    const int remain = static_cast<int>(scaled) - static_cast<int>(kScaledBase[i]);
    if (remain >= 0) {
      level = kLevel[i];
      return kTargetBase[i] + (static_cast<uint32_t>(remain) << level);
    }
  }
  return 0;
}

// express: handler_capsules.cpp:206-266
uint32_t emit_express(const uint8_t* prev, const uint8_t* cur, orc_node_hq* out) {
  const int cur_q8 = static_cast<int>((rd16(cur + 2) & 0x7FFFu) << 2);
  const int prev_q8 = static_cast<int>((rd16(prev + 2) & 0x7FFFu) << 2);
  int diff_q8 = cur_q8 - prev_q8;
  if (prev_q8 > cur_q8) diff_q8 += (360 << 8);
  const int inc_q16 = diff_q8 << 3;
  int a_q16 = prev_q8 << 8;
  uint32_t n = 0;
  for (int pos = 0; pos < 16; ++pos) {
    const uint8_t* cab = prev + 4 + 5 * pos;
    const uint32_t da[2] = {rd16(cab), rd16(cab + 2)};
    const int off_q3[2] = {static_cast<int>((cab[4] & 0xF) | ((da[0] & 0x3) << 4)),
                           static_cast<int>((cab[4] >> 4) | ((da[1] & 0x3) << 4))};
    for (int c = 0; c < 2; ++c) {
      const int dist_q2 = static_cast<int>(da[c] & 0xFFFCu);
      const int angle_q6 = (a_q16 - (off_q3[c] << 13)) >> 10;
      const int sync = (((a_q16 + inc_q16) % (360 << 16)) < inc_q16) ? 1 : 0;
      a_q16 += inc_q16;
      out[n++] = make_node(angle_q6, dist_q2, sync, dist_q2 ? (0x2F << 2) : 0);
    }
  }
  return n;
}

// ultra: handler_capsules.cpp:460-580
uint32_t emit_ultra(const uint8_t* prev, const uint8_t* cur, orc_node_hq* out) {
  const int cur_q8 = static_cast<int>((rd16(cur + 2) & 0x7FFFu) << 2);
  const int prev_q8 = static_cast<int>((rd16(prev + 2) & 0x7FFFu) << 2);
  int diff_q8 = cur_q8 - prev_q8;
  if (prev_q8 > cur_q8) diff_q8 += (360 << 8);
  const int inc_q16 = (diff_q8 << 3) / 3;
  int a_q16 = prev_q8 << 8;
  uint32_t n = 0;
  for (int pos = 0; pos < 32; ++pos) {
    const uint32_t x3 = rd32(prev + 4 + 4 * pos);
    int dist_major = static_cast<int>(x3 & 0xFFF);
    int predict1 = static_cast<int>(x3 << 10) >> 22;  // signed 10-bit fields
    int predict2 = static_cast<int>(x3) >> 22;
    int dist_major2 = static_cast<int>((pos == 31 ? rd32(cur + 4) : rd32(prev + 4 + 4 * (pos + 1))) & 0xFFF);
    uint32_t lvl1 = 0, lvl2 = 0;
    dist_major = static_cast<int>(varbitscale_decode(static_cast<uint32_t>(dist_major), lvl1));
    dist_major2 = static_cast<int>(varbitscale_decode(static_cast<uint32_t>(dist_major2), lvl2));
    int base1 = dist_major;
    const int base2 = dist_major2;
    if (!dist_major && dist_major2) {
      base1 = dist_major2;
      lvl1 = lvl2;
    }
    int dist_q2[3];
    dist_q2[0] = dist_major << 2;
    if (static_cast<uint32_t>(predict1) == 0xFFFFFE00u || static_cast<uint32_t>(predict1) == 0x1FFu) {
      dist_q2[1] = 0;
    } else {
      predict1 = static_cast<int>(static_cast<uint32_t>(predict1) << lvl1);
      dist_q2[1] = static_cast<int>(static_cast<uint32_t>(predict1 + base1) << 2);
    }
    if (static_cast<uint32_t>(predict2) == 0xFFFFFE00u || static_cast<uint32_t>(predict2) == 0x1FFu) {
      dist_q2[2] = 0;
    } else {
      predict2 = static_cast<int>(static_cast<uint32_t>(predict2) << lvl2);
      dist_q2[2] = static_cast<int>(static_cast<uint32_t>(predict2 + base2) << 2);
    }
    for (int c = 0; c < 3; ++c) {
      const int sync = (((a_q16 + inc_q16) % (360 << 16)) < inc_q16) ? 1 : 0;
      int off_q16 = static_cast<int>(7.5 * 3.1415926535 * (1 << 16) / 180.0);
      if (dist_q2[c] >= (50 * 4)) {
        const int k1 = 98361;
        const int k2 = k1 / dist_q2[c];
        off_q16 = static_cast<int>(8 * 3.1415926535 * (1 << 16) / 180) - (k2 << 6) - (k2 * k2 * k2) / 98304;
      }
      const int angle_q6 = (a_q16 - static_cast<int>(off_q16 * 180 / 3.14159265)) >> 10;
      a_q16 += inc_q16;
      out[n++] = make_node(angle_q6, dist_q2[c], sync, dist_q2[c] ? (0x2F << 2) : 0);
    }
  }
  return n;
}

// ultra-dense: handler_capsules.cpp:951-1047; state[0] = _last_node_sync_bit, state[1] = _last_dist_q2
// returns 0 and sets *discard when the angular jump is above the 100 Hz bound
uint32_t emit_ultra_dense(const uint8_t* prev, const uint8_t* cur, uint32_t sample_duration_us, uint32_t* state,
                          orc_node_hq* out, bool* discard) {
  const int cur_q8 = static_cast<int>((rd16(cur + 8) & 0x7FFFu) << 2);
  const int prev_q8 = static_cast<int>((rd16(prev + 8) & 0x7FFFu) << 2);
  int diff_q8 = cur_q8 - prev_q8;
  if (prev_q8 > cur_q8) diff_q8 += (360 << 8);
  const int thr_q8 = (360 * 100 * 32 / static_cast<int>(1000000 / sample_duration_us)) << 8;
  *discard = diff_q8 > thr_q8;
  if (*discard) return 0;
  int last_sync = static_cast<int>(state[0] & 1u), last_dist = static_cast<int>(state[1]);
  const int inc_q16 = (diff_q8 << 8) / 64;
  int a_q16 = prev_q8 << 8;
  uint32_t n = 0;
  for (int pos = 0; pos < 64; ++pos) {
    const uint8_t* cab = prev + 10 + 5 * (pos >> 1);
    const uint32_t qds = (pos & 1) ? (rd16(cab + 2) | (static_cast<uint32_t>(cab[4] >> 4) << 16))
                                   : (rd16(cab) | (static_cast<uint32_t>(cab[4] & 0x0F) << 16));
    int quality = 0, dist_q2 = 0;
    switch (qds & 3u) {
      case 0:
        quality = static_cast<uint8_t>(qds >> 12);
        dist_q2 = static_cast<int>(qds & 0xFFCu) * 2;
        if (last_dist && std::abs(dist_q2 - last_dist) <= 8) dist_q2 = (dist_q2 + last_dist) >> 1;
        break;
      case 1:
        quality = static_cast<uint8_t>((qds >> 13) << 1);
        dist_q2 = static_cast<int>(qds & 0x1FFCu) * 3 + (2046 << 2);
        break;
      case 2:
        quality = static_cast<uint8_t>((qds >> 14) << 2);
        dist_q2 = static_cast<int>(qds & 0x3FFCu) * 4 + (8187 << 2);
        break;
      default:
        quality = static_cast<uint8_t>((qds >> 15) << 3);
        dist_q2 = static_cast<int>(qds & 0x7FFCu) * 5 + (24567 << 2);
        break;
    }
    last_dist = dist_q2;
    const int angle_q6 = a_q16 >> 10;
    int sync = (((a_q16 + inc_q16) % (360 << 16)) < (inc_q16 << 1)) ? 1 : 0;
    sync = (sync ^ last_sync) & sync;
    a_q16 += inc_q16;
    out[n++] = make_node(angle_q6, dist_q2, sync, quality);
    last_sync = sync;
  }
  state[0] = static_cast<uint32_t>(last_sync);
  state[1] = static_cast<uint32_t>(last_dist);
  return n;
}

uint32_t crc32_padded(const uint8_t* p, uint32_t len) {  // sl_crc.cpp:52-99 (reflected 0x04C11DB7)
  static uint32_t table[256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int j = 0; j < 8; ++j) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
      table[i] = c;
    }
    ready = true;
  }
  uint32_t crc = 0xFFFFFFFFu;
  for (uint32_t i = 0; i < len; ++i) crc = (crc >> 8) ^ table[(crc ^ p[i]) & 0xFFu];
  const uint32_t pad = 4 - (len & 3u);  // 4 zero bytes when len is already a multiple of 4
  for (uint32_t i = 0; i < pad; ++i) crc = (crc >> 8) ^ table[crc & 0xFFu];
  return crc ^ 0xFFFFFFFFu;
}

}  // namespace

extern "C" uint32_t orc_capsule_bytes(uint32_t ans_type) {
  switch (ans_type) {
    case ORC_ANS_EXPRESS: return 84;
    case ORC_ANS_HQ: return 781;
    case ORC_ANS_ULTRA: return 132;
    case ORC_ANS_DENSE: return 84;
    case ORC_ANS_ULTRA_DENSE: return 170;
    default: return 0;
  }
}
extern "C" uint32_t orc_capsule_nodes(uint32_t ans_type) {
  switch (ans_type) {
    case ORC_ANS_EXPRESS: return 32;
    case ORC_ANS_HQ: return 96;
    case ORC_ANS_ULTRA: return 96;
    case ORC_ANS_DENSE: return 40;
    case ORC_ANS_ULTRA_DENSE: return 64;
    default: return 0;
  }
}
extern "C" uint32_t orc_crc32_padded(const uint8_t* p, uint32_t len) { return crc32_padded(p, len); }

// state[0]: last node sync bit (dense, ultra-dense), state[1]: last distance (ultra-dense)
extern "C" uint32_t orc_decode_capsules(uint32_t ans_type, const uint8_t* capsules, uint32_t n_capsules,
                                        uint32_t sample_duration_us, uint32_t* state, orc_node_hq* nodes_out,
                                        uint32_t* capsule_status, uint32_t* capsule_node_offset) {
  if (ans_type == ORC_ANS_DENSE)
    return orc_dense_decode(capsules, n_capsules, sample_duration_us, &state[0], nodes_out, capsule_status,
                            capsule_node_offset);
  const uint32_t cb = orc_capsule_bytes(ans_type);
  if (!cb) return 0;
  const uint32_t sum_from = (ans_type == ORC_ANS_ULTRA_DENSE) ? 2u : 2u;  // checksum starts after the two marker bytes
  const uint32_t start_off = (ans_type == ORC_ANS_ULTRA_DENSE) ? 8u : 2u;
  bool prev_ready = false;
  const uint8_t* prev = nullptr;
  uint32_t n_out = 0;
  for (uint32_t j = 0; j < n_capsules; ++j) {
    const uint8_t* c = capsules + static_cast<size_t>(j) * cb;
    uint32_t st = 0;
    if (capsule_node_offset) capsule_node_offset[j] = n_out;
    if (ans_type == ORC_ANS_HQ) {  // handler_hqnode.cpp:93-160 (no inter-capsule state)
      if (c[0] != 0xA5) {
        st = ORC_CAPSULE_BAD_FRAME;
      } else if (crc32_padded(c, cb - 4) != rd32(c + cb - 4)) {
        st = ORC_CAPSULE_CHECKSUM_ERR;
      } else {
        std::memcpy(nodes_out + n_out, c + 9, 96 * sizeof(orc_node_hq));
        n_out += 96;
        st = ORC_CAPSULE_OK | ORC_CAPSULE_EMIT;
      }
      if (capsule_status) capsule_status[j] = st;
      continue;
    }
    if ((c[0] >> 4) != 0xA || (c[1] >> 4) != 0x5) {
      prev_ready = false;
      if (capsule_status) capsule_status[j] = ORC_CAPSULE_BAD_FRAME;
      continue;
    }
    uint8_t sum = 0;
    for (uint32_t b = sum_from; b < cb; ++b) sum ^= c[b];
    const uint8_t recv = static_cast<uint8_t>((c[0] & 0xF) | (c[1] << 4));
    if (recv != sum) {
      prev_ready = false;
      if (capsule_status) capsule_status[j] = ORC_CAPSULE_CHECKSUM_ERR;
      continue;
    }
    st |= ORC_CAPSULE_OK;
    if (rd16(c + start_off) & 0x8000u) {
      st |= ORC_CAPSULE_SYNC;
      if (prev_ready) st |= ORC_CAPSULE_ENCODER_RESET_ERR;
      prev_ready = false;
    }
    if (prev_ready) {
      uint32_t k = 0;
      if (ans_type == ORC_ANS_EXPRESS) {
        k = emit_express(prev, c, nodes_out + n_out);
      } else if (ans_type == ORC_ANS_ULTRA) {
        k = emit_ultra(prev, c, nodes_out + n_out);
      } else {
        bool discard = false;
        k = emit_ultra_dense(prev, c, sample_duration_us, state, nodes_out + n_out, &discard);
        if (discard) st |= ORC_CAPSULE_DISCARD;
      }
      if (k) st |= ORC_CAPSULE_EMIT;
      n_out += k;
    }
    prev = c;
    prev_ready = true;
    if (capsule_status) capsule_status[j] = st;
  }
  return n_out;
}

// Standard measurement nodes from a raw byte stream (handler_normalnode.cpp:88-141): 5-byte records,
// byte 0 must carry a sync bit and its inverse, byte 1 a set check bit; a byte failing its test is
// dropped (and for byte 1 the record restarts).  node_end (nullable) receives, per node, the index
// of the record's last byte.  *fsm_pos: bytes buffered on entry / exit (0 on a fresh decoder; the
// buffered bytes themselves are not carried: a record split across two calls is lost, so callers
// hand over whole streams).
extern "C" uint32_t orc_decode_normal(const uint8_t* bytes, uint32_t n_bytes, orc_node_hq* nodes_out,
                                      uint32_t* node_end, uint32_t* fsm_pos) {
  uint8_t buf[5] = {0, 0, 0, 0, 0};
  uint32_t pos = 0, n_out = 0;
  for (uint32_t i = 0; i < n_bytes; ++i) {
    const uint8_t b = bytes[i];
    if (pos == 0) {
      if (!(((b >> 1) ^ b) & 1u)) continue;
    } else if (pos == 1) {
      if (!(b & 1u)) {
        pos = 0;
        continue;
      }
    } else if (pos == 4) {
      buf[4] = b;
      pos = 0;
      const uint32_t angle_chk = static_cast<uint32_t>(buf[1]) | (static_cast<uint32_t>(buf[2]) << 8);
      orc_node_hq nd;
      nd.angle_z_q14 = static_cast<uint16_t>(((angle_chk >> 1) << 8) / 90);
      nd.dist_mm_q2 = static_cast<uint32_t>(buf[3]) | (static_cast<uint32_t>(buf[4]) << 8);
      nd.flag = buf[0] & 1u;
      nd.quality = static_cast<uint8_t>((buf[0] >> 2) << 2);
      if (node_end) node_end[n_out] = i;
      nodes_out[n_out++] = nd;
      continue;
    }
    buf[pos++] = b;
  }
  if (fsm_pos) *fsm_pos = pos;
  return n_out;
}


// ---- byte-level framing (the sync-nibble hunt in front of every capsule decoder) -------------------------------
// UnpackerHandler_CapsuleNode::onData            reference handler_capsules.cpp:107-135
// UnpackerHandler_UltraCapsuleNode::onData       :324-353
// UnpackerHandler_DenseCapsuleNode::onData       :639-668
// UnpackerHandler_UltraDenseCapsuleNode::onData  :852-880
// restated as the byte machine it is (`buf_pos` = _cached_scan_node_buf_pos): bytes are consumed one at a time;
// a completed frame is appended to capsules_out; whenever the machine clears _is_previous_capsuledataRdy while
// hunting, ONE all-zero capsule is appended before the next frame (the framed decoders treat it as a bad frame,
// which clears the same flag).  Returns the number of capsules; *bytes_left = bytes buffered at the end.
extern "C" uint32_t orc_frame_capsules(uint32_t ans_type, const uint8_t* bytes, uint32_t n, uint8_t* capsules_out,
                                       uint32_t max_capsules, uint32_t* bytes_left) {
  const uint32_t cb = orc_capsule_bytes(ans_type);
  std::vector<uint8_t> buf(cb ? cb : 1);
  uint32_t buf_pos = 0, count = 0;
  bool lost = false;
  if (cb == 0) return 0;
  for (uint32_t pos = 0; pos < n; ++pos) {
    const uint8_t cur = bytes[pos];
    switch (buf_pos) {
      case 0:
        if ((cur >> 4) != 0xA) {
          lost = true;  // _is_previous_capsuledataRdy = false; continue
          continue;
        }
        break;
      case 1:
        if ((cur >> 4) != 0x5) {
          buf_pos = 0;
          lost = true;
          continue;
        }
        break;
      default:
        break;
    }
    buf[buf_pos++] = cur;
    if (buf_pos == cb) {
      buf_pos = 0;
      if (lost) {
        if (count < max_capsules) std::memset(capsules_out + (size_t)count * cb, 0, cb);
        ++count;
        lost = false;
      }
      if (count < max_capsules) std::memcpy(capsules_out + (size_t)count * cb, buf.data(), cb);
      ++count;
    }
  }
  if (bytes_left) *bytes_left = buf_pos;
  return count;
}
