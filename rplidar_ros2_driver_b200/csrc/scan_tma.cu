// scan_tma.cu -- fast scan kernel v2: TMA-staged scan tiles (producer warp + mbarrier ring).
//
// Same contract and the same two passes as scan_fast.cu (mark keys, fold to a rank table,
// place by rank; ascendScanData_ + publish_scan, reference
// src/sdk/src/sl_lidar_driver.cpp:128-184 and src/rplidar_node.cpp:581-677), restructured so
// that no consumer warp ever waits on a global load:
//
//   * a dedicated producer warp streams the scan tile through a shared-memory ring with
//     1-D bulk TMA copies (cp.async.bulk, 8 KB chunks, mbarrier full/empty handshake).
//     The chunk sequence is [scan s pass 1][scan s pass 2][scan s+1 pass 1]...: the producer
//     runs ahead across pass and scan boundaries, so HBM/L2 latency is hidden behind the ring
//     depth instead of behind resident warps.  Pass-1 copies carry an L2 evict_last hint, the
//     pass-2 copies (L2 hits) and all output stores evict_first.
//   * marking keeps v1's 64 KB byte map (plain byte stores, ~10 instructions per node).  A
//     first version of this kernel ORed key bits into the bitmap with warp-aggregated
//     shared-memory atomics instead; ncu showed ~100 instructions per node in that loop
//     (profiles/ncu_r1_scan_tma_atomics.txt), so it was dropped.
//
// Serves launches without the ascended node buffer (that variant needs two more shared
// tables and stays on scan_fast.cu).  Requires every scan base to be 16-byte aligned (nodes
// pointer 16 B aligned, even stride); the host falls back to scan_fast.cu otherwise.
#include <type_traits>

#include "rpl_device.cuh"
#include "scan_args.h"
#include "scan_common.cuh"

namespace rpl {

namespace {

#ifndef RPL_TMA_TC
#define RPL_TMA_TC 512
#endif
#ifndef RPL_TMA_CTAS
#define RPL_TMA_CTAS 2
#endif
constexpr int TC = RPL_TMA_TC;         // consumer threads (512 or 1024)
constexpr int kCtasPerSm = RPL_TMA_CTAS;
constexpr int kCWarps = TC / 32;
constexpr uint32_t kRowBytes = kKeySpace / TC;   // byte-map bytes folded by one thread
constexpr uint32_t kCols = kRowBytes / 16;       // 16-byte columns per row
constexpr uint32_t kWordsPerThread = kWords / TC;
static_assert(TC == 512 || TC == 1024, "fold layout is written for 512 or 1024 consumer threads");
constexpr int kBlock = TC + 32;        // + one producer warp
#ifndef RPL_TMA_CH
#define RPL_TMA_CH 1024
#endif
#ifndef RPL_TMA_STAGES
#define RPL_TMA_STAGES 4
#endif
constexpr uint32_t CH = RPL_TMA_CH;          // nodes per chunk (8 bytes each)
constexpr int kStages = RPL_TMA_STAGES;      // ring depth; CH * kStages * 8 = 32 KB
constexpr int kRounds = CH / TC;       // nodes per consumer thread per chunk

struct __align__(128) TmaSmem {
  uint8_t bytemap[kKeySpace];                // presence map (swizzled)
  uint2 ring[kStages][CH];
  uint2 rankV[kWords];                       // {bits, exclusive prefix} over measured keys
  unsigned long long full[kStages];
  unsigned long long empty[kStages];
  uint32_t red[4 * kCWarps];
  uint32_t valid_count;
  uint32_t totV;
  uint32_t fallback;
};

// byte-map swizzle: within the row a thread folds, the 16-byte column is XORed with row bits so
// that the 128-bit reads of 8 neighbouring threads hit 8 different bank groups
__device__ __forceinline__ uint32_t swz_x(uint32_t x) {
  if (TC == 512) return (x ^ ((x >> 3) & 0x70u)) & 0xFFFFu;  // 128-byte rows: column ^= row & 7
  return (x ^ ((x >> 3) & 0x30u)) & 0xFFFFu;                 // 64-byte rows: column ^= (row >> 1) & 3
}
__device__ __forceinline__ uint32_t gather4(uint32_t x) { return (x * 0x10204080u) >> 28; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// producer-side wait: the ring is usually full, so let the hardware park the thread (suspend
// time hint, ns) instead of spinning through issue slots the consumer warps need
__device__ __forceinline__ void mbar_wait_relaxed(void* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(20000u)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// 1-D bulk TMA copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, void* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC) : "memory"); }

// which scans run through the ring (producer and consumers must agree)
__device__ __forceinline__ bool scan_is_streamed(uint32_t n, uint32_t stride, uint32_t max_nodes) {
  return n != 0 && n <= stride && n <= max_nodes && n <= kMaxFastNodes;
}

// MODE: 0 = LaserScan Mode B, 1 = LaserScan Mode A, 2 = PointCloud2 (window + polar->xyz)
template <int MODE>
__global__ void __launch_bounds__(kBlock, kCtasPerSm) scan_tma_kernel(ScanBatchArgs a, FastWorkspace ws) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TmaSmem& sm = *reinterpret_cast<TmaSmem*>(smem_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&sm.full[i], 1);         // the producer's arrive.expect_tx
      mbar_init(&sm.empty[i], kCWarps);  // one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  // =========================== producer warp ===========================================
  if (warp == kCWarps) {
    if (lane == 0) {
      const uint64_t pol_keep = l2_policy_evict_last();
      const uint64_t pol_stream = l2_policy_evict_first();
      uint32_t stage = 0, parity = 1;  // empty-barrier parity: first round passes at once
      for (uint32_t s = blockIdx.x; s < a.n_scans; s += gridDim.x) {
        const uint32_t n = a.counts[s];
        if (!scan_is_streamed(n, a.stride, ws.max_nodes)) continue;
        const uint2* base = a.nodes + (size_t)s * a.stride;
        const uint32_t nch = (n + CH - 1) / CH;
        for (int pass = 0; pass < 2; ++pass) {
          for (uint32_t c = 0; c < nch; ++c) {
            mbar_wait_relaxed(&sm.empty[stage], parity);
            // an odd tail is rounded up to a whole 16 bytes; the extra node lies inside the
            // scan's stride (even stride, n odd => n + 1 <= stride) and is masked by consumers
            const uint32_t cn = min(CH, n - c * CH);
            const uint32_t bytes = ((cn + 1u) & ~1u) * 8u;
            mbar_expect_tx(&sm.full[stage], bytes);
            tma_load_1d(&sm.ring[stage][0], base + (size_t)c * CH, bytes, &sm.full[stage],
                        pass == 0 ? pol_keep : pol_stream);
            if (++stage == kStages) {
              stage = 0;
              parity ^= 1u;
            }
          }
        }
      }
    }
    return;
  }

  // =========================== consumer warps ==========================================
  const bool new_proto = a.is_new_protocol != 0;
  const bool inverted = a.inverted != 0;
  const uint64_t pol_stream = l2_policy_evict_first();
  const uint32_t q_shift = new_proto ? 16u : 18u, q_mask = new_proto ? 0xFFu : 0x3Fu;
  constexpr bool MODE_A = (MODE == 1);
  constexpr bool CLOUD = (MODE == 2);
  const float w_rmin = a.range_min, w_rmax = a.range_max, w_imin = a.intensity_min;
  // intensity straight from word y without the conversion pipe (exact for values < 2^23)
  auto intensity_of = [&](uint32_t y) {
    return __fsub_rn(__uint_as_float(((y >> q_shift) & q_mask) | 0x4B000000u), 8388608.0f);
  };
  uint32_t stage = 0, parity = 0;  // ring position of the next chunk to consume
  using Checked = std::integral_constant<bool, true>;
  using Unchecked = std::integral_constant<bool, false>;
  auto advance = [&]() {
    if (++stage == kStages) {
      stage = 0;
      parity ^= 1u;
    }
  };

  // hand the current ring slot back to the producer.  Called after the memory operations
  // (byte-map / output stores) whose addresses depend on the values loaded from the slot, so
  // the arrive cannot issue before those loads have returned.
  auto release = [&]() {
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[stage]);
  };
  auto drain = [&](uint32_t nch) {  // consume chunks without looking at them
    for (uint32_t c = 0; c < nch; ++c) {
      mbar_wait(&sm.full[stage], parity);
      release();
      advance();
    }
  };

  for (uint32_t s = blockIdx.x; s < a.n_scans; s += gridDim.x) {
    const uint32_t n = a.counts[s];

    if (n > a.stride || n > ws.max_nodes) {  // caller error: report, touch nothing
      if (tid == 0) {
        if (a.status) a.status[s] = 0x80008000u;  // SL_RESULT_INVALID_DATA
        if (a.path) a.path[s] = 0u;
        if (a.beam_counts) a.beam_counts[s] = 0u;
        if (a.angle_inc) a.angle_inc[s] = 0.0f;
      }
      continue;
    }
    if (n > kMaxFastNodes) {  // cannot be tie-free: general kernel
      if (tid == 0) a.fallback_list[atomicAdd(a.fallback_count, 1u)] = s;
      continue;
    }
    if (n == 0) {  // ascendScanData: OPERATION_FAIL; publish_scan: nodes.empty() -> return
      if (tid == 0) {
        if (a.status) a.status[s] = a.apply_ascend ? kResultOperationFail : kResultOk;
        if (a.path) a.path[s] = 0u;
        if (a.beam_counts) a.beam_counts[s] = 0u;
        if (a.angle_inc) a.angle_inc[s] = 0.0f;
      }
      continue;
    }
    const uint32_t nch = (n + CH - 1) / CH;

    // ---- phase 0: clear the presence map ------------------------------------------------
    {
      uint4* bm = reinterpret_cast<uint4*>(sm.bytemap);
      const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (uint32_t j = 0; j < kKeySpace / 16 / TC; ++j) bm[j * TC + tid] = z;
      if (tid == 0) {
        sm.fallback = 0;
      }
    }
    consumer_sync();

    // ---- phase 1 (mark): one byte store per measured key ---------------------------------
    uint32_t cnt = 0;
    uint8_t* const bmap = sm.bytemap;
    auto mark_chunk = [&](auto checked, uint32_t c) {
      mbar_wait(&sm.full[stage], parity);
      const uint2* slot = sm.ring[stage];
      uint2 v[kRounds];
#pragma unroll
      for (int r = 0; r < kRounds; ++r) v[r] = slot[r * TC + tid];
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const uint32_t dist = __funnelshift_r(v[r].x, v[r].y, 16);
        bool valid = dist != 0;
        if (CLOUD) valid = valid && cloud_keep(dist_to_m(dist), intensity_of(v[r].y), w_rmin, w_rmax, w_imin);
        if (decltype(checked)::value && c * CH + r * TC + tid >= n) valid = false;
        if (valid) bmap[swz_x(v[r].x)] = 1;
        cnt += valid ? 1u : 0u;
      }
      release();
      advance();
    };
    const uint32_t nfull = n / CH;
    for (uint32_t c = 0; c < nfull; ++c) mark_chunk(Unchecked{}, c);
    if (nfull < nch) mark_chunk(Checked{}, nfull);
    cnt = warp_sum(cnt);
    if (lane == 0) sm.red[warp] = cnt;
    consumer_sync();

    // ---- fold: byte map -> bitmap + exclusive popcount prefix ------------------------------
    {
      // thread t owns keys [kRowBytes t, kRowBytes (t+1)) = kWordsPerThread bitmap words
      uint32_t wv[kWordsPerThread];
#pragma unroll
      for (uint32_t j = 0; j < kWordsPerThread; ++j) wv[j] = 0;
      const uint4* bm = reinterpret_cast<const uint4*>(sm.bytemap);
      const uint32_t colx = (TC == 512) ? (tid & 7u) : ((tid >> 1) & 3u);
#pragma unroll
      for (uint32_t c = 0; c < kCols; ++c) {
        const uint4 q = bm[tid * kCols + (c ^ colx)];  // physical column of logical chunk c
        const uint32_t x[4] = {q.x, q.y, q.z, q.w};
        uint32_t bv = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) bv |= gather4(x[j] & 0x01010101u) << (4 * j);
        wv[c >> 1] |= bv << (16 * (c & 1));
      }
      uint32_t sv = 0;
#pragma unroll
      for (uint32_t j = 0; j < kWordsPerThread; ++j) sv += __popc(wv[j]);
      const uint32_t iv = warp_inclusive_scan(sv);
      if (lane == 31) sm.red[2 * kCWarps + warp] = iv;
      consumer_sync();
      if (warp == 0) {
        uint32_t tv = lane < kCWarps ? sm.red[2 * kCWarps + lane] : 0u;
        uint32_t cc = lane < kCWarps ? sm.red[lane] : 0u;
        const uint32_t cv = warp_inclusive_scan(tv);
        cc = warp_sum(cc);
        if (lane < kCWarps) sm.red[2 * kCWarps + lane] = cv - tv;
        if (lane == 31) {
          sm.totV = cv;
          sm.valid_count = cc;
        }
      }
      consumer_sync();
      uint32_t pv = sm.red[2 * kCWarps + warp] + iv - sv;
#pragma unroll
      for (uint32_t j = 0; j < kWordsPerThread; ++j) {
        sm.rankV[tid * kWordsPerThread + j] = make_uint2(wv[j], pv);
        pv += __popc(wv[j]);
      }
    }
    consumer_sync();
    const uint32_t M = sm.valid_count;

    if (M == 0) {
      // ascendScanData: OPERATION_FAIL, buffer untouched; publish_scan: nothing to publish
      if (tid == 0) {
        if (a.status) a.status[s] = a.apply_ascend ? kResultOperationFail : kResultOk;
        if (a.path) a.path[s] = 0u;
        if (a.beam_counts) a.beam_counts[s] = 0u;
        if (a.angle_inc) a.angle_inc[s] = 0.0f;
      }
      drain(nch);
      consumer_sync();
      continue;
    }
    // duplicate keys (fewer distinct keys than measured nodes) -> general kernel (stable rule);
    // so do Mode A scans too large for the shared-memory index map
    if (sm.totV != M || (MODE_A && M > kModeASmemMaxPoints)) {
      if (tid == 0) a.fallback_list[atomicAdd(a.fallback_count, 1u)] = s;
      drain(nch);
      consumer_sync();
      continue;
    }

    // ---- phase 2 (place): rank and place --------------------------------------------------
    float* ranges = CLOUD ? nullptr : a.ranges + (size_t)s * a.stride;
    float* intens = CLOUD ? nullptr : a.intensities + (size_t)s * a.stride;
    float4* cloud = CLOUD ? a.xyzi + (size_t)s * a.stride : nullptr;
    uint16_t* sidx = reinterpret_cast<uint16_t*>(sm.bytemap);  // Mode A: u-rank -> node index (map is dead)
    const uint2* base = a.nodes + (size_t)s * a.stride;
    const float inc = angle_increment(M, MODE_A);
    const bool has0 = (sm.rankV[0].x & 1u) != 0;
    // Mode B output slot = ob + os * rank in wrapping u32 arithmetic (reference
    // rplidar_node.cpp:673); intensities[] sits at a fixed byte distance from ranges[]
    const uint32_t ob = inverted ? M - 1u : 0u, os = inverted ? 0xFFFFFFFFu : 1u;
    const ptrdiff_t i_minus_r = reinterpret_cast<char*>(intens) - reinterpret_cast<char*>(ranges);

    auto place_chunk = [&](auto checked, uint32_t c) {
      mbar_wait(&sm.full[stage], parity);
      const uint2* slot = sm.ring[stage];
      uint2 v[kRounds];
#pragma unroll
      for (int r = 0; r < kRounds; ++r) v[r] = slot[r * TC + tid];
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const uint2 nd = v[r];
        const uint32_t k = nd.x & 0xFFFFu;
        const uint32_t dist = __funnelshift_r(nd.x, nd.y, 16);
        uint32_t measured = dist != 0 ? 1u : 0u;
        if (decltype(checked)::value && c * CH + r * TC + tid >= n) measured = 0;
        const uint32_t rk = rank_of(sm.rankV, k);
        const float dm = dist_to_m(dist);
        if (CLOUD) {  // polar -> xyz at the rank among kept points (oracle/cloud_oracle.cpp 1-3)
          const float it = intensity_of(nd.y);
          if (!cloud_keep(dm, it, w_rmin, w_rmax, w_imin)) measured = 0;
          const float2 cs = __ldg(a.trig + k);
          st_f32x4_if(cloud + rk, make_float4(__fmul_rn(dm, cs.x), __fmul_rn(dm, cs.y), 0.0f, it), pol_stream,
                      measured);
        } else if (!MODE_A) {  // Mode B: reference rplidar_node.cpp:661-677
          const uint32_t o = ob + os * rk;
          const float it = intensity_of(nd.y);
          float* pr = ranges + o;
          st_f32_if(pr, dm, pol_stream, measured);
          st_f32_if(reinterpret_cast<float*>(reinterpret_cast<char*>(pr) + i_minus_r), it, pol_stream, measured);
        } else if (measured) {  // Mode A: remember which node sits at this u-rank (mode_a_emit_smem)
          sidx[mode_a_urank(k, rk, M, inverted, has0)] = (uint16_t)(c * CH + r * TC + tid);
        }
      }
      release();
      advance();
    };
    for (uint32_t c = 0; c < nfull; ++c) place_chunk(Unchecked{}, c);
    if (nfull < nch) place_chunk(Checked{}, nfull);
    consumer_sync();

    // ---- phase 3 (Mode A): resolve bins that hold several points --------------------------
    if (MODE_A) {
      ModeAOut mo;
      mo.ranges = ranges;
      mo.intens = intens;
      mo.angle = a.angle;
      mo.M = M;
      mo.inc = inc;
      mo.inverted = inverted;
      mo.new_proto = new_proto;
      mo.policy = pol_stream;
      // the rank table is dead by now: each warp stages a batch of bins in its own slice of it
      static_assert(kEmit2Stage * 2 * kCWarps <= sizeof(sm.rankV), "bin staging must fit the rank table");
      mode_a_emit_smem(mo, sidx, base, warp, kCWarps,
                       reinterpret_cast<uint16_t*>(sm.rankV) + warp * ((kEmit2Stage + 7u) & ~7u));
    }
    consumer_sync();
    if (tid == 0) {
      if (sm.fallback) {
        a.fallback_list[atomicAdd(a.fallback_count, 1u)] = s;
      } else {
        if (a.status) a.status[s] = kResultOk;
        if (a.path) a.path[s] = 0u;
        if (a.beam_counts) a.beam_counts[s] = M;
        if (a.angle_inc) a.angle_inc[s] = inc;
      }
    }
    consumer_sync();
  }
}

}  // namespace

cudaError_t launch_scan_tma(const ScanBatchArgs& a, const FastWorkspace& ws, int grid, cudaStream_t stream) {
  if (a.xyzi) scan_tma_kernel<2><<<grid, kBlock, sizeof(TmaSmem), stream>>>(a, ws);
  else if (a.mode_a) scan_tma_kernel<1><<<grid, kBlock, sizeof(TmaSmem), stream>>>(a, ws);
  else scan_tma_kernel<0><<<grid, kBlock, sizeof(TmaSmem), stream>>>(a, ws);
  return cudaGetLastError();
}

cudaError_t scan_tma_configure() {
  const int sh = (int)sizeof(TmaSmem);
  cudaError_t e = cudaFuncSetAttribute(scan_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sh);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(scan_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sh);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(scan_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sh);
}

int scan_tma_max_ctas_per_sm() {
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, scan_tma_kernel<0>, kBlock, sizeof(TmaSmem));
  return nb;
}

}  // namespace rpl
