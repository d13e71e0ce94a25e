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
// One CTA per stream.  Scan starts are rare (one per revolution), so the flag pass just appends
// their positions to a small shared-memory list (unordered, one atomic per scan start) and rank-sorts
// it; reset counts come from a prefix over the capsule flags + binary search; published scans are
// compacted into descriptors and then copied coalesced.  A stream with more scan starts than the
// list holds takes the chunked block-scan path instead (same result, more barriers).
#include "decode_args.h"
#include "rpl_device.cuh"

namespace rpl {

namespace {

constexpr int AT = 512;
constexpr uint32_t kListCap = 4096, kResetCap = 1024;  // scan starts per stream handled by the list path
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

// ascending rank sort of a short list (duplicates keep their list order)
__device__ __forceinline__ void rank_sort(const uint32_t* in, uint32_t* out, uint32_t k) {
  for (uint32_t e = threadIdx.x; e < k; e += AT) {
    const uint32_t v = in[e];
    uint32_t r = 0;
    for (uint32_t j = 0; j < k; ++j) {
      const uint32_t w = in[j];
      r += (w < v || (w == v && j < e)) ? 1u : 0u;
    }
    out[r] = v;
  }
}

__global__ void __launch_bounds__(AT) assemble_kernel(AssembleArgs a) {
  __shared__ uint32_t s_list[kListCap], s_sorted[kListCap];      // scan-start positions
  __shared__ uint32_t s_rlist[kResetCap], s_rsorted[kResetCap];  // reset positions
  __shared__ uint32_t s_cnt, s_rcnt;
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

    if (tid == 0) {
      s_cnt = 0;
      s_rcnt = 0;
      s_last_sync = -1;
      s_published = 0;
    }
    __syncthreads();
    // ---- flag pass: positions of the scan-start nodes and of the reset requests -----------------------
    for (uint32_t j = tid; j < ncap; j += AT) {
      if (cst[j] & kStSync) {
        const uint32_t idx = atomicAdd(&s_rcnt, 1u);
        if (idx < kResetCap) s_rlist[idx] = coff[j];
      }
    }
    // the decoder may have handed the scan-start positions over (rpl_decode_dense_batch_starts_dev): then the node
    // stream is not read again at all
    const uint32_t n_listed = (a.scan_starts && a.scan_start_counts) ? a.scan_start_counts[s] : 0xFFFFFFFFu;
    if (n_listed <= a.starts_stride && n_listed <= kListCap) {
      const uint32_t* lst = a.scan_starts + (size_t)s * a.starts_stride;
      for (uint32_t e = tid; e < n_listed; e += AT) s_list[e] = lst[e];
      if (tid == 0) s_cnt = n_listed;
    } else {
      const uint32_t* flags = reinterpret_cast<const uint32_t*>(nodes) + 1;  // word 1 of every node
      uint32_t i = tid;
      for (; i + 3 * AT < n; i += 4 * AT) {
        const uint32_t y0 = __ldg(flags + 2 * (size_t)i), y1 = __ldg(flags + 2 * (size_t)(i + AT));
        const uint32_t y2 = __ldg(flags + 2 * (size_t)(i + 2 * AT)), y3 = __ldg(flags + 2 * (size_t)(i + 3 * AT));
        if (((y0 | y1 | y2 | y3) >> 24) & 1u) {
          const uint32_t yy[4] = {y0, y1, y2, y3};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if ((yy[u] >> 24) & 1u) {
              const uint32_t idx = atomicAdd(&s_cnt, 1u);
              if (idx < kListCap) s_list[idx] = i + u * AT;
            }
          }
        }
      }
      for (; i < n; i += AT) {
        if ((__ldg(flags + 2 * (size_t)i) >> 24) & 1u) {
          const uint32_t idx = atomicAdd(&s_cnt, 1u);
          if (idx < kListCap) s_list[idx] = i;
        }
      }
    }
    __syncthreads();
    const uint32_t K = s_cnt, RK = s_rcnt;
    if (K <= kListCap && RK <= kResetCap) {
      // ---- list path ---------------------------------------------------------------------------------
      rank_sort(s_list, s_sorted, K);
      rank_sort(s_rlist, s_rsorted, RK);
      __syncthreads();
      auto resets_upto = [&](uint32_t x) -> uint32_t {  // number of reset positions <= x
        uint32_t lo = 0, hi = RK;
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (s_rsorted[mid] <= x) lo = mid + 1;
          else hi = mid;
        }
        return lo;
      };
      for (uint32_t k0 = 0; k0 + 1 < K; k0 += AT) {
        const uint32_t k = k0 + tid;
        uint32_t publish = 0, st = 0, en = 0;
        if (k + 1 < K) {
          st = s_sorted[k];
          en = s_sorted[k + 1];
          publish = (resets_upto(en) == resets_upto(st)) ? 1u : 0u;
        }
        uint32_t tot = 0;
        const uint32_t ex = block_excl_scan(publish, s_warp, &tot);
        if (publish) {
          const uint32_t slot = s_published + ex;
          if (slot < a.max_scans) desc[slot] = make_uint2(st, en - st);
        }
        __syncthreads();
        if (tid == 0) s_published += tot;
        __syncthreads();
      }
    } else {
      // ---- many scan starts: chunked block scans over the flags ---------------------------------------
      uint32_t carry = 0;
      for (uint32_t c0 = 0; c0 < ncap; c0 += AT) {
        const uint32_t j = c0 + tid;
        const uint32_t v = (j < ncap && (cst[j] & kStSync)) ? 1u : 0u;
        uint32_t tot = 0;
        const uint32_t ex = block_excl_scan(v, s_warp, &tot);
        if (j < ncap) rs[j] = carry + ex + v;
        carry += tot;
      }
      __syncthreads();
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
    }
    __syncthreads();
    const uint32_t total = s_published;
    if (tid == 0) a.scans_per_stream[s] = total;
    // ---- copy the published scans (descriptors were written by this CTA: visible after the barrier) ---
    const uint32_t stored = min(total, a.max_scans);
    if (a.views_out) {
      // view mode: no copy.  A published scan is handed on as (first node, count) into the node buffer itself;
      // a scan that hit the holder's capacity gets the holder's "replace the last entry" applied in place
      // (the nodes behind the cap are dropped either way).
      uint2* vout = a.views_out + (size_t)s * a.max_scans;
      for (uint32_t k = tid; k < a.max_scans; k += AT) {
        uint2 v = make_uint2(0u, 0u);
        if (k < stored) {
          const uint2 d = desc[k];
          const uint32_t cnt = min(d.y, a.max_nodes);
          if (d.y > cnt) a.nodes_mut[(size_t)s * a.stride_nodes + d.x + cnt - 1] = nodes[d.x + d.y - 1];
          v = make_uint2((uint32_t)((size_t)s * a.stride_nodes + d.x), cnt);
          if (a.scan_begin_ts_us)
            a.scan_begin_ts_us[(size_t)s * a.max_scans + k] =
                a.node_ts_us ? a.node_ts_us[(size_t)s * a.stride_nodes + d.x] : 0ull;
        }
        vout[k] = v;
        out_len[k] = v.y;
      }
      __syncthreads();
      continue;
    }
    for (uint32_t k = 0; k < stored; ++k) {
      const uint2 d = desc[k];
      const uint32_t cnt = min(d.y, a.max_nodes);
      uint2* o = out + (size_t)k * a.scan_stride;
      const uint2* src = nodes + d.x;
      // a scan that hit the cap kept overwriting its last entry: its last slot takes the scan's last node.
      // The copy leaves that slot alone (one writer per output element, no ordering between threads needed).
      const bool capped = d.y > cnt;
      const uint32_t ncopy = capped ? cnt - 1 : cnt;
      uint32_t q = tid;
      for (; q + 3 * AT < ncopy; q += 4 * AT) {  // four independent loads in flight per thread
        const uint2 v0 = ld_hint_v2(src + q, l2_policy_evict_first()), v1 = ld_hint_v2(src + q + AT, l2_policy_evict_first());
        const uint2 v2 = ld_hint_v2(src + q + 2 * AT, l2_policy_evict_first()), v3 = ld_hint_v2(src + q + 3 * AT, l2_policy_evict_first());
        o[q] = v0;
        o[q + AT] = v1;
        o[q + 2 * AT] = v2;
        o[q + 3 * AT] = v3;
      }
      for (; q < ncopy; q += AT) o[q] = src[q];
      if (tid == 0) {
        if (capped) o[cnt - 1] = src[d.y - 1];
        out_len[k] = cnt;
        // the scan-start node's stamp (ScanDataHolder::_scan_begin_timestamp_uS, sl_lidar_driver.cpp:293)
        if (a.scan_begin_ts_us)
          a.scan_begin_ts_us[(size_t)s * a.max_scans + k] =
              a.node_ts_us ? a.node_ts_us[(size_t)s * a.stride_nodes + d.x] : 0ull;
      }
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
