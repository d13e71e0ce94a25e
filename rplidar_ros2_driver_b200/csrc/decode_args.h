// decode_args.h -- argument block of the dense-capsule decode kernel (decode.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

namespace rpl {

struct DecodeArgs {
  const uint8_t* capsules;        // [n_streams][stride_capsules][84], stream bases 4-byte aligned
  const uint32_t* counts;         // [n_streams] capsules per stream
  uint32_t n_streams;
  uint32_t stride_capsules;
  uint32_t sample_duration_us;    // SlamtecLidarTimingDesc::sample_duration_uS (jump threshold)
  const uint32_t* sync_state_in;  // [n_streams] lastNodeSyncBit entering each stream (nullable: 0)
  uint2* nodes_out;               // [n_streams][stride_capsules * 40]
  uint32_t* node_counts;          // [n_streams]
  uint32_t* capsule_status;       // [n_streams][stride_capsules] nullable
  uint32_t* capsule_node_offset;  // [n_streams][stride_capsules] nullable
  uint32_t* sync_state_out;       // [n_streams] nullable
  // positions (node offsets) of the scan-start nodes of every stream, in no particular order (nullable): what
  // rpl_assemble_scan_views_dev otherwise finds by reading every node again
  uint32_t* scan_starts;          // [n_streams][starts_stride]
  uint32_t* scan_start_counts;    // [n_streams] (may exceed starts_stride: the list is then incomplete)
  uint32_t starts_stride;
};

// the other capsule formats (decode_formats.cu): 0x82 express, 0x83 HQ, 0x84 ultra, 0x86 ultra-dense
struct CapsuleDecodeArgs {
  const uint8_t* capsules;        // [n_streams][stride_capsules][capsule bytes]
  const uint32_t* counts;         // [n_streams]
  uint32_t n_streams, stride_capsules, sample_duration_us;
  const uint32_t* state_in;       // [n_streams][2] {last node sync bit, last distance} (nullable: 0)
  uint2* nodes_out;               // [n_streams][stride_capsules * nodes per capsule]
  uint32_t* node_counts;          // [n_streams]
  uint32_t* capsule_status;       // nullable
  uint32_t* capsule_node_offset;  // nullable
  uint32_t* state_out;            // [n_streams][2] nullable
};

// 5-byte standard nodes from raw byte streams
struct NormalDecodeArgs {
  const uint8_t* bytes;           // [n_streams][stride_bytes]
  const uint32_t* byte_counts;    // [n_streams]
  uint32_t n_streams, stride_bytes;
  uint2* nodes_out;               // [n_streams][stride_bytes / 5]
  uint32_t* node_counts;          // [n_streams]
  uint32_t* fsm_state_out;        // [n_streams] nullable: bytes buffered when the stream ended
  uint32_t* node_end;             // [n_streams][stride_bytes / 5] nullable: index of each record's last byte
};

// per-sample timestamps (timestamps.cu)
struct TimingDesc {  // sl::SlamtecLidarTimingDesc without the bool
  uint32_t sample_duration_us, native_baudrate, linkage_delay_us, native_interface_type;
};
struct TimestampArgs {
  const unsigned long long* capsule_rx_us;  // [n_streams][stride_capsules]
  const uint32_t* capsule_status;
  const uint32_t* capsule_node_offset;
  const uint32_t* capsule_counts;           // [n_streams]
  uint32_t n_streams, stride_capsules;
  unsigned long long* node_ts_us;           // [n_streams][stride_capsules * nodes per capsule]
};
struct NormalTimestampArgs {
  const uint32_t* node_end;                 // [n_streams][stride_nodes] (decode_normal's report)
  const uint32_t* node_counts;              // [n_streams]
  uint32_t n_streams, stride_nodes, chunk_bytes, stride_chunks;
  const unsigned long long* chunk_rx_us;    // [n_streams][stride_chunks]
  unsigned long long* node_ts_us;           // [n_streams][stride_nodes]
};
cudaError_t launch_node_timestamps(uint32_t ans_type, const TimingDesc& t, const TimestampArgs& a, cudaStream_t stream);
cudaError_t launch_normal_timestamps(const TimingDesc& t, const NormalTimestampArgs& a, cudaStream_t stream);

cudaError_t launch_decode_capsules(uint32_t ans_type, const CapsuleDecodeArgs& a, int grid, cudaStream_t stream);
cudaError_t launch_decode_normal(const NormalDecodeArgs& a, int grid, cudaStream_t stream);
cudaError_t decode_formats_configure();

struct AssembleArgs {
  const uint2* nodes;                 // [n_streams][stride_nodes] decoded node streams
  const uint32_t* node_counts;        // [n_streams]
  uint32_t n_streams, stride_nodes;
  const uint32_t* capsule_status;     // [n_streams][stride_capsules] (nullable: no scan resets)
  const uint32_t* capsule_node_offset;
  const uint32_t* capsule_counts;     // [n_streams]
  uint32_t stride_capsules;
  uint32_t max_nodes;                 // ScanDataHolder capacity (8192 in the SDK)
  uint32_t max_scans, scan_stride;
  uint2* scans_out;                   // [n_streams][max_scans][scan_stride] (copy mode)
  uint2* views_out;                   // [n_streams][max_scans] {first node in the whole buffer, count} (view mode)
  uint2* nodes_mut;                   // view mode: the node buffer, writable (capacity rule applied in place)
  uint32_t* scan_len;                 // [n_streams][max_scans]
  uint32_t* scans_per_stream;         // [n_streams] published scans (may exceed max_scans)
  const unsigned long long* node_ts_us;  // [n_streams][stride_nodes] nullable
  unsigned long long* scan_begin_ts_us;  // [n_streams][max_scans] nullable
  const uint32_t* scan_starts;        // nullable: the decoder's list of scan-start node positions per stream
  const uint32_t* scan_start_counts;
  uint32_t starts_stride;
  uint32_t* reset_prefix;             // scratch [n_streams][stride_capsules]
  uint2* desc;                        // scratch [n_streams][max_scans]
};

// byte-level framing with the SDK's resynchronisation (frame.cu)
struct FrameArgs {
  const uint8_t* bytes;           // [n_streams][stride_bytes] raw capsule streams
  const uint32_t* byte_counts;    // [n_streams]
  uint32_t n_streams, stride_bytes;
  uint32_t capsule_bytes;         // frame size of the answer type
  uint8_t* capsules_out;          // [n_streams][stride_capsules][capsule_bytes]
  uint32_t stride_capsules;
  uint32_t* capsule_counts_out;   // [n_streams]
  uint32_t* bytes_left_out;       // [n_streams] nullable: bytes of an unfinished frame at the end
};
cudaError_t launch_frame_capsules(const FrameArgs& a, int grid, cudaStream_t stream);
cudaError_t launch_assemble(const AssembleArgs& a, int grid, cudaStream_t stream);
cudaError_t launch_decode_dense(const DecodeArgs& a, int grid, cudaStream_t stream);
cudaError_t decode_configure();  // opt-in dynamic shared memory, once per device

}  // namespace rpl
