// decode_args.h -- argument block of the dense-capsule decode kernel (decode.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
};

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
  uint2* scans_out;                   // [n_streams][max_scans][scan_stride]
  uint32_t* scan_len;                 // [n_streams][max_scans]
  uint32_t* scans_per_stream;         // [n_streams] published scans (may exceed max_scans)
  uint32_t* reset_prefix;             // scratch [n_streams][stride_capsules]
  uint2* desc;                        // scratch [n_streams][max_scans]
};

cudaError_t launch_assemble(const AssembleArgs& a, int grid, cudaStream_t stream);
cudaError_t launch_decode_dense(const DecodeArgs& a, int grid, cudaStream_t stream);
cudaError_t decode_configure();  // opt-in dynamic shared memory, once per device

}  // namespace rpl
