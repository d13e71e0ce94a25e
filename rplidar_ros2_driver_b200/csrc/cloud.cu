// cloud.cu -- PointCloud2 path (placeholder until the scan path is verified on the GPU).
#include "cloud_args.h"

namespace rpl {
cudaError_t cloud_configure() { return cudaSuccess; }
cudaError_t cloud_workspace_alloc(CloudWorkspace&, int, uint32_t) { return cudaSuccess; }
void cloud_workspace_free(CloudWorkspace&) {}
cudaError_t launch_cloud(const CloudBatchArgs&, const CloudWorkspace&, int, cudaStream_t, int*) {
  return cudaErrorNotSupported;
}
cudaError_t launch_cloud_fuse(const float4*, const uint32_t*, uint32_t, uint32_t, float4*, uint32_t*,
                              uint32_t*, cudaStream_t, int*) {
  return cudaErrorNotSupported;
}
}  // namespace rpl
