// synth.cu -- synthetic scan streams generated directly in HBM.
//
// Counter-based splitmix64, one thread per node, same definition as
// oracle/scan_oracle.cpp::orc_synth_scan so that the CPU baseline and the parity tests see
// byte-identical buffers.  The node shape follows what the SDK hands to the hot path
// (reference src/sdk/include/sl_lidar_cmd.h:272-278; quality 188 = 0x2F<<2 is what the dense
// capsule unpacker emits, reference src/sdk/src/dataunpacker/unpacker/handler_capsules.cpp:778).
#include "rpl_device.cuh"
#include "scan_args.h"

namespace rpl {

namespace {

__device__ __forceinline__ uint64_t draw(uint64_t seed, uint64_t i, uint64_t field) {
  return mix64(seed + i * 8 + field);
}

// pseudo-random permutation of [0,n): 4-round Feistel on an even number of bits, cycle walking
__device__ uint32_t feistel_perm(uint64_t seed, uint32_t x, uint32_t n) {
  uint32_t bits = 2;
  while ((1ull << bits) < n) ++bits;
  if (bits & 1) ++bits;
  const uint32_t half = bits / 2, mask = (1u << half) - 1;
  uint32_t v = x;
  do {
    uint32_t l = v >> half, r = v & mask;
    for (uint32_t round = 0; round < 4; ++round) {
      const uint32_t f = (uint32_t)mix64(seed + 0x1000000ull * (round + 1) + r) & mask;
      const uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    v = (l << half) | r;
  } while (v >= n);
  return v;
}

__global__ void synth_kernel(uint64_t first_scan_id, uint32_t n_scans, uint32_t n, uint32_t stride,
                             int variant, uint2* nodes, uint32_t* counts) {
  const uint32_t s = blockIdx.y;
  if (s >= n_scans) return;
  const uint64_t seed = mix64(0x5EED0000ull + first_scan_id + s);
  const uint32_t rot = (uint32_t)(draw(seed, 0xFFFFFFFFull, 7) % n);
  if (counts && blockIdx.x == 0 && threadIdx.x == 0) counts[s] = n;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const uint32_t i = (variant == 3) ? feistel_perm(seed, p, n) : (p + rot) % n;
    uint32_t key;
    if (variant == 2) {
      key = (uint32_t)(draw(seed, i, 0) & 0xFFFF);
    } else {
      const uint32_t b0 = (uint32_t)(((uint64_t)i << 16) / n);
      const uint32_t b1 = (uint32_t)(((uint64_t)(i + 1) << 16) / n);
      const uint32_t span = b1 - b0;
      key = b0 + (span > 1 ? (uint32_t)(draw(seed, i, 0) % span) : 0u);
    }
    const bool invalid = (draw(seed, i, 1) % 100) < 5;
    uint32_t dist = 600u + (uint32_t)(draw(seed, i, 2) % 159401ull);
    if (variant == 4) {  // a "room": 16 angular segments of constant range (2..10 m) + 2 cm noise
      const uint32_t seg = (uint32_t)(((uint64_t)i * 16) / n);
      dist = 8000u + (uint32_t)(draw(seed, seg, 4) % 32001ull) + (uint32_t)(draw(seed, i, 2) % 161ull) - 80u;
    }
    if (invalid) dist = 0u;
    uint32_t q = (variant == 1) ? (uint32_t)(draw(seed, i, 3) & 0xFF) : 188u;
    if (invalid) q = 0;
    const uint32_t flag = (p == 0) ? 1u : 2u;
    uint2 nd;
    nd.x = (key & 0xFFFFu) | (dist << 16);
    nd.y = (dist >> 16) | (q << 16) | (flag << 24);
    nodes[(size_t)s * stride + p] = nd;
  }
}

}  // namespace

cudaError_t launch_synth(uint64_t first_scan_id, uint32_t n_scans, uint32_t n, uint32_t stride,
                         int variant, uint2* nodes, uint32_t* counts, cudaStream_t stream) {
  if (n_scans == 0 || n == 0) return cudaSuccess;
  const int threads = 256;
  uint32_t bx = (n + threads * 4 - 1) / (threads * 4);
  if (bx == 0) bx = 1;
  // gridDim.y is limited to 65535: walk the batch in slabs
  for (uint32_t s0 = 0; s0 < n_scans; s0 += 65535) {
    const uint32_t ns = min(65535u, n_scans - s0);
    synth_kernel<<<dim3(bx, ns), threads, 0, stream>>>(first_scan_id + s0, ns, n, stride, variant,
                                                        nodes + (size_t)s0 * stride,
                                                        counts ? counts + s0 : nullptr);
  }
  return cudaGetLastError();
}

}  // namespace rpl
