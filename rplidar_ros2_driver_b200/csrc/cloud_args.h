// cloud_args.h -- argument block of the PointCloud2 path (extensions; see cloud.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rpl {

struct CloudBatchArgs {
  const uint2* nodes;      // [n_scans][stride]
  const uint32_t* counts;  // [n_scans]
  uint32_t n_scans;
  uint32_t stride;
  float4* xyzi;            // [n_scans][stride] x, y, z, intensity
  uint32_t* point_counts;  // [n_scans]
  float range_min, range_max, intensity_min, voxel_size;
  uint32_t sor_k;
  float sor_alpha;
  uint8_t is_new_protocol;
};

struct CloudWorkspace {
  float2* trig = nullptr;        // [65536] (cos, sin) of angle_rad(key), rounded from double
  void* scratch = nullptr;       // per-CTA staging, see cloud.cu
  size_t scratch_per_cta = 0;
  uint32_t max_nodes = 0;
  int ctas = 0;
};

cudaError_t cloud_configure();
cudaError_t cloud_workspace_alloc(CloudWorkspace& ws, int num_sms, uint32_t max_nodes);
void cloud_workspace_free(CloudWorkspace& ws);
cudaError_t launch_cloud(const CloudBatchArgs& a, const CloudWorkspace& ws, int num_sms,
                         cudaStream_t stream, int* launched);
cudaError_t launch_cloud_fuse(const float4* xyzi, const uint32_t* point_counts, uint32_t n_scans,
                              uint32_t stride, float4* fused, uint32_t* offsets, uint32_t* total,
                              cudaStream_t stream, int* launched);

}  // namespace rpl
