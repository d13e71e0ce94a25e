// sample_data_unpacker_b200.hpp -- host mirror of the SDK's sample-data unpacker seam (SURVEY.md 8(f) rank 1
// and 4) on top of the C-ABI.
//
// The reference decodes measurement answers through
//   LIDARSampleDataUnpacker::onSampleData(ansType, buffer, size)      src/sdk/src/dataunpacker/dataunpacker.cpp:123-146
// which calls back a
//   LIDARSampleDataListener { onHQNodeScanResetReq(); onHQNodeDecoded(timestamp_uS, node);
//                             onDecodingError(errMsg, ansType, payload, size); }   dataunpacker.h:47-57
// (SlamtecLidarDriver is the listener: src/sdk/src/sl_lidar_driver.cpp:1645-1662).  GpuSampleDataUnpacker
// keeps those names and the order of the callbacks, but decodes whole batches of capsules on the
// GPU: onSampleData() only queues framed capsules (with their receive time, which the SDK reads
// from getus() at that point), flush() runs rpl_decode_capsules once and replays, capsule by capsule,
// exactly the callbacks the SDK would have made -- checksum errors, encoder-reset errors and scan-reset
// requests in their place between the nodes, every node with the SDK's timestamp.  The capsule that
// closes a batch is kept and fed again in front of the next one (a capsule's nodes are released by its
// successor), with its own events suppressed the second time.
// Contract: capsule answer types 0x82..0x86, buffers holding whole capsules (what the protocol codec hands
// over); there is no CPU decoder behind this class -- without a B200 the constructor throws.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rpl_b200.h"
#include "sdk_types.hpp"

namespace rplidar_b200 {

class SampleDataListener {  // == sl::internal::LIDARSampleDataListener
 public:
  virtual ~SampleDataListener() = default;
  virtual void onHQNodeScanResetReq() = 0;
  virtual void onHQNodeDecoded(uint64_t timestamp_uS, const sl_lidar_response_measurement_node_hq_t* node) = 0;
  virtual void onDecodingError(int /*errMsg*/, uint8_t /*ansType*/, const void* /*payload*/, size_t /*size*/) {}
};

class GpuSampleDataUnpacker {
 public:
  enum { ERR_EVENT_ON_EXP_ENCODER_RESET = 0x8001, ERR_EVENT_ON_EXP_CHECKSUM_ERR = 0x8002 };  // dataunpacker.h:62-65

  GpuSampleDataUnpacker(SampleDataListener& listener, int device = 0, uint32_t batch_capsules = 256)
      : listener_(listener), batch_(batch_capsules ? batch_capsules : 1) {
    if (rpl_ctx_create(device, 8192, 1, &ctx_) != RPL_RESULT_OK || !ctx_)
      throw std::runtime_error("GpuSampleDataUnpacker: no usable B200 (there is no CPU decoder)");
    timing_.sample_duration_us = 476;  // the SDK's value before the first timing update (legacy A1 rate)
  }
  ~GpuSampleDataUnpacker() { rpl_ctx_destroy(ctx_); }
  GpuSampleDataUnpacker(const GpuSampleDataUnpacker&) = delete;
  GpuSampleDataUnpacker& operator=(const GpuSampleDataUnpacker&) = delete;

  // updateUnpackerContext(UNPACKER_CONTEXT_TYPE_LIDAR_TIMING, ...)   dataunpacker.cpp:113-121
  void updateTiming(const rpl_timing& t) {
    flush();
    timing_ = t;
  }
  void enable() { reset(); enabled_ = true; }
  void disable() { reset(); enabled_ = false; }
  void reset() {  // handler->reset(): cached capsule and decoder state are dropped
    queue_.clear();
    rx_.clear();
    have_carry_in_queue_ = false;
    state_[0] = state_[1] = 0;
    ans_ = 0;
  }

  // Queues whole capsules received at `rx_time_us`; decodes when `batch_capsules` are waiting.
  bool onSampleData(uint8_t ansType, const void* buffer, size_t size, uint64_t rx_time_us) {
    if (!enabled_) return false;
    const uint32_t cb = rpl_capsule_bytes(ansType);
    if (cb == 0 || size % cb != 0) return false;
    if (ansType != ans_) {  // another answer type selects another handler: the cached capsule is gone
      flush();
      reset();
      ans_ = ansType;
    }
    const uint8_t* p = static_cast<const uint8_t*>(buffer);
    for (size_t off = 0; off < size; off += cb) {
      queue_.insert(queue_.end(), p + off, p + off + cb);
      rx_.push_back(rx_time_us);
      if (rx_.size() - (have_carry_in_queue_ ? 1 : 0) >= batch_) flush();
    }
    return true;
  }

  // Decodes what is queued and replays the SDK's callbacks.  Returns false on a device error.
  bool flush() {
    const uint32_t cb = rpl_capsule_bytes(ans_), per = rpl_capsule_nodes(ans_);
    const uint32_t n = static_cast<uint32_t>(rx_.size());
    const uint32_t first_new = have_carry_in_queue_ ? 1u : 0u;
    if (cb == 0 || n <= first_new) return true;
    nodes_.resize(static_cast<size_t>(n) * per);
    ts_.resize(nodes_.size());
    status_.resize(n);
    offs_.resize(n);
    uint32_t count = 0;
    // the carried capsule releases nothing when it is fed again in front (it has no predecessor there), so the
    // decoder state carried in state_ is consumed exactly once
    const rpl_result r = rpl_decode_capsules(ctx_, ans_, queue_.data(), n, timing_.sample_duration_us, state_,
                                             reinterpret_cast<rpl_node_hq*>(nodes_.data()), &count, status_.data(),
                                             offs_.data(), &timing_, rx_.data(), ts_.data());
    if (r != RPL_RESULT_OK) {
      last_error_ = rpl_last_error(ctx_);
      return false;
    }
    for (uint32_t j = first_new; j < n; ++j) {
      const uint32_t st = status_[j];
      const uint8_t* payload = queue_.data() + static_cast<size_t>(j) * cb;
      if (st & RPL_CAPSULE_CHECKSUM_ERR)
        listener_.onDecodingError(ERR_EVENT_ON_EXP_CHECKSUM_ERR, ans_, payload, cb);
      if (st & RPL_CAPSULE_SYNC) {
        if (st & RPL_CAPSULE_ENCODER_RESET_ERR)
          listener_.onDecodingError(ERR_EVENT_ON_EXP_ENCODER_RESET, ans_, payload, cb);
        listener_.onHQNodeScanResetReq();
      }
      if (st & RPL_CAPSULE_EMIT)
        for (uint32_t k = 0; k < per; ++k) listener_.onHQNodeDecoded(ts_[offs_[j] + k], &nodes_[offs_[j] + k]);
    }
    // keep the last capsule: its nodes are released by the first capsule of the next batch
    std::vector<uint8_t> last(queue_.end() - cb, queue_.end());
    const uint64_t last_rx = rx_.back();
    queue_.swap(last);
    rx_.assign(1, last_rx);
    have_carry_in_queue_ = true;
    return true;
  }

  const std::string& last_error() const { return last_error_; }

 private:
  SampleDataListener& listener_;
  rpl_ctx* ctx_ = nullptr;
  uint32_t batch_;
  bool enabled_ = false;
  bool have_carry_in_queue_ = false;  // queue_[0] is the capsule that closed the previous batch
  uint8_t ans_ = 0;
  rpl_timing timing_{};
  uint32_t state_[2] = {0, 0};
  std::vector<uint8_t> queue_;
  std::vector<uint64_t> rx_;
  std::vector<sl_lidar_response_measurement_node_hq_t> nodes_;
  std::vector<uint64_t> ts_;
  std::vector<uint32_t> status_, offs_;
  std::string last_error_;
};

}  // namespace rplidar_b200
