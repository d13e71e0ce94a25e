// cdr_args.h -- argument blocks of the CDR serialisation kernels (cdr.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

namespace rpl {

struct LaserScanMeta {  // == rpl_laserscan_meta
  int32_t stamp_sec;
  uint32_t stamp_nanosec;
  float angle_min, angle_max, angle_increment, time_increment, scan_time, range_min, range_max;
};

// fixed part of a message, built on the host (rpl_capi.cu): copied, then patched per message
struct CdrTemplate {
  uint8_t prefix[512];
  uint32_t prefix_bytes;    // multiple of 4
  uint32_t patch_width;     // PointCloud2: offset of `width`
  uint32_t patch_row_step;  // PointCloud2: offset of `row_step`
};

struct LaserScanCdrArgs {
  const LaserScanMeta* meta;      // [n_scans] device
  const float* angle_increment;   // [n_scans] device, nullable: overrides meta[s].angle_increment
  const float* ranges;            // [n_scans][stride]
  const float* intensities;       // [n_scans][stride]
  const uint32_t* beam_counts;    // [n_scans]
  uint32_t n_scans, stride;
  uint8_t* cdr_out;               // [n_scans][cdr_stride], cdr_stride % 4 == 0
  uint32_t cdr_stride;
  uint32_t* cdr_sizes;            // [n_scans] nullable
};

struct PointCloudCdrArgs {
  const uint32_t* stamps;         // [n_clouds][2] {sec, nanosec}
  const float* xyzi;              // [n_clouds][stride][4]
  const uint32_t* point_counts;   // [n_clouds]
  uint32_t n_clouds, stride;
  uint8_t* cdr_out;               // [n_clouds][cdr_stride], cdr_stride % 16 == 0
  uint32_t cdr_stride;
  uint32_t* cdr_sizes;
};

cudaError_t launch_laserscan_cdr(const LaserScanCdrArgs& a, const CdrTemplate& t, cudaStream_t stream);
cudaError_t launch_pointcloud2_cdr(const PointCloudCdrArgs& a, const CdrTemplate& t, cudaStream_t stream);

}  // namespace rpl
