// assemble.cu -- scan assembly on the GPU: cuts decoded node streams into
// scans, the bridge between decode.cu and the scan kernels without a host round trip.
//
// Replaces ScanDataHolder::pushScanNodeData / rewindCurrentScanData (reference
// src/sdk/src/sl_lidar_driver.cpp:272-315) as a pure function of the node stream and the
// scan-reset requests (one per scan-start capsule, SlamtecLidarDriver::onHQNodeScanResetReq
// :1651-1653):
//   * a node with flag bit 0 opens a scan and publishes the scan in progress if it holds anything;
//     nodes before the first such node are dropped; the last, unfinished scan is not published;
//   * a reset empties the scan in progress: the scan [s, e) between two scan-start nodes is
//     published iff no reset position r satisfies s < r <= e;
//   * a scan longer than max_nodes (8192 in the SDK) keeps overwriting its last entry.
// One CTA per stream: scan starts are found with block scans over the sync flags, reset counts with
// a prefix over the capsule flags + binary search, published scans are compacted into descriptors
// and then copied coalesced.
#include "decode_args.h"
#include "rpl_device.cuh"

namespace rpl {

namespace {

constexpr int AT = 256;
constexpr uint32_t kStSync = 2;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_tot, uint32_t* total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t inc = warp_inclusive_scan(v);
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < AT / 32; ++w) {
    const uint32_t t = warp_tot[w];
    if ((uint32_t)w < warp) base += t;
    tot += t;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(AT) assemble_kernel(AssembleArgs a) {
  __shared__ uint32_t s_warp[AT / 32];
  __shared__ int s_wmax[AT / 32];
  __shared__ int s_last_sync;       // position of the latest scan-start node seen so far (-1: none)
  __shared__ uint32_t s_published;  // scans published so far
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (uint32_t s = blockIdx.x; s < a.n_streams; s += gridDim.x) {
    const uint32_t n = a.node_counts[s];
    const uint2* nodes = a.nodes + (size_t)s * a.stride_nodes;
    const bool have_resets = a.capsule_status != nullptr;
    const uint32_t ncap = have_resets ? a.capsule_counts[s] : 0u;
    const uint32_t* cst = have_resets ? a.capsule_status + (size_t)s * a.stride_capsules : nullptr;
    const uint32_t* coff = have_resets ? a.capsule_node_offset + (size_t)s * a.stride_capsules : nullptr;
    uint32_t* rs = a.reset_prefix + (size_t)s * a.stride_capsules;  // inclusive count of reset capsules
    uint2* desc = a.desc + (size_t)s * a.max_scans;                 // (start, length) of published scans
    uint2* out = a.scans_out + (size_t)s * a.max_scans * a.scan_stride;
    uint32_t* out_len = a.scan_len + (size_t)s * a.max_scans;

    // ---- resets: inclusive prefix count of scan-start capsules ----------------------------------
    {
      uint32_t carry = 0;
      for (uint32_t c0 = 0; c0 < ncap; c0 += AT) {
        const uint32_t j = c0 + tid;
        const uint32_t v = (j < ncap && (cst[j] & kStSync)) ? 1u : 0u;
        uint32_t tot = 0;
        const uint32_t ex = block_excl_scan(v, s_warp, &tot);
        if (j < ncap) rs[j] = carry + ex + v;
        carry += tot;
      }
    }
    if (tid == 0) {
      s_last_sync = -1;
      s_published = 0;
    }
    __syncthreads();
    // R(x) = number of reset positions <= x
    auto resets_upto = [&](uint32_t x) -> uint32_t {
      if (!have_resets || ncap == 0) return 0u;
      uint32_t lo = 0, hi = ncap;  // first capsule with offset > x
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (coff[mid] <= x) lo = mid + 1;
        else hi = mid;
      }
      return lo ? rs[lo - 1] : 0u;
    };

    // ---- find the scans: a scan-start node at position e closes the scan opened at the previous one --
    for (uint32_t c0 = 0; c0 < n; c0 += AT) {
      const uint32_t i = c0 + tid;
      const bool sync = i < n && ((nodes[i].y >> 24) & 1u);
      // previous scan-start position: exclusive running maximum of the sync positions
      int m = sync ? (int)i : -1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, m, o);
        if (lane >= (uint32_t)o) m = max(m, t);
      }
      if (lane == 31) s_wmax[warp] = m;
      __syncthreads();
      int prev = __shfl_up_sync(0xffffffffu, m, 1);
      if (lane == 0) prev = -1;
      int before = s_last_sync;
      for (uint32_t w = 0; w < warp; ++w) before = max(before, s_wmax[w]);
      prev = max(prev, before);
      uint32_t publish = 0;
      if (sync && prev >= 0) publish = (resets_upto(i) == resets_upto((uint32_t)prev)) ? 1u : 0u;
      uint32_t tot = 0;
      const uint32_t ex = block_excl_scan(publish, s_warp, &tot);  // syncs
      if (publish) {
        const uint32_t k = s_published + ex;
        if (k < a.max_scans) desc[k] = make_uint2((uint32_t)prev, i - (uint32_t)prev);
      }
      __syncthreads();
      if (tid == 0) {
        int last = s_last_sync;
        for (int w = 0; w < AT / 32; ++w) last = max(last, s_wmax[w]);
        s_last_sync = last;
        s_published += tot;
      }
      __syncthreads();
    }
    const uint32_t total = s_published;
    if (tid == 0) a.scans_per_stream[s] = total;
    // ---- copy the published scans (descriptors were written by this CTA: visible after the barrier) ---
    const uint32_t stored = min(total, a.max_scans);
    for (uint32_t k = 0; k < stored; ++k) {
      const uint2 d = desc[k];
      const uint32_t cnt = min(d.y, a.max_nodes);
      uint2* o = out + (size_t)k * a.scan_stride;
      for (uint32_t q = tid; q < cnt; q += AT) {
        // a scan that hit the cap kept overwriting its last entry: it ends with the scan's last node
        const uint32_t src = (q == cnt - 1 && d.y > cnt) ? d.x + d.y - 1 : d.x + q;
        o[q] = nodes[src];
      }
      if (tid == 0) out_len[k] = cnt;
    }
    __syncthreads();
  }
}

}  // namespace

cudaError_t launch_assemble(const AssembleArgs& a, int grid, cudaStream_t stream) {
  if (a.n_streams == 0) return cudaSuccess;
  assemble_kernel<<<grid, AT, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace rpl
