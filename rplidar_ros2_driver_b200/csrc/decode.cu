// decode.cu -- dense-capsule decode on the GPU (the step immediately
// before the hot path).
//
// Replaces, for framed S2/S3 DenseBoost capsules (answer type 0x85, 84 bytes, reference
// src/sdk/include/sl_lidar_cmd.h:223-234):
//   UnpackerHandler_DenseCapsuleNode::onData                       handler_capsules.cpp:639-734
//   UnpackerHandler_DenseCapsuleNode::_onScanNodeDenseCapsuleData   handler_capsules.cpp:736-791
// (reference src/sdk/src/dataunpacker/unpacker/).  One capsule carries a start angle and 40 raw
// distances; its 40 HQ nodes can only be produced when the NEXT capsule arrives (angles are
// interpolated between the two start angles), and only if both checksums hold, the next capsule
// is not a scan start, and the angular step is below the 100 Hz bound.  The reference does this
// byte by byte in a state machine; here every capsule is a thread:
//
//   * the stream is staged through shared memory in tiles of 256 capsules (coalesced 128-bit
//     loads; 21-word capsule stride is conflict free),
//   * "does capsule j release its predecessor?" depends only on capsules j-1 and j, so the output
//     position of every node is an exclusive block scan of the release flags,
//   * the one genuinely sequential piece of state, the reference's function-static
//     lastNodeSyncBit (sync_i = raw_i & ~sync_{i-1}), is carried across capsules by scanning
//     2-bit transfer functions (out(0), out(1)) under composition,
//   * nodes are written by a thread per PAIR of samples (one 32-bit word of the capsule, shared look-ups, one
//     16-byte store), 40 consecutive nodes per capsule: fully coalesced.
//
// Wire bytes in: 84 per capsule (2.1 B per point); nodes out: 320 per released capsule.
#include "decode_args.h"
#include "rpl_device.cuh"

namespace rpl {

namespace {

constexpr int DT = 256;              // threads = capsules per tile
constexpr int kCapWords = 21;        // 84 bytes
constexpr uint32_t kStOk = 1, kStSync = 2, kStEmit = 4, kStDiscard = 8, kStChecksum = 16, kStEncReset = 32,
                   kStBadFrame = 64;

struct __align__(16) DecodeSmem {
  uint32_t cap[2][DT * kCapWords];     // double-buffered tiles (cp.async prefetch of the next tile)
  uint32_t carry[kCapWords];           // last capsule of the previous tile
  unsigned long long smask[DT];        // final sync bits of the 40 nodes each capsule releases
  uint32_t start_q8[DT + 1];           // (start & 0x7FFF) << 2, slot 0 = carry
  int inc_q16[DT];                     // angular step per sample of the capsule each capsule releases
  uint32_t okflag[DT + 1];             // checksum + frame ok, slot 0 = carry
  uint32_t emit_list[DT];              // releasing capsules of the tile, in order (compacted)
  uint32_t warp_a[DT / 32], warp_b[DT / 32];
  uint32_t carry_nodes;                // nodes written so far in this stream
  uint32_t carry_sync;                 // lastNodeSyncBit entering the tile
  uint32_t tile_nodes;
  uint32_t red_sync;                   // lastNodeSyncBit leaving the tile
  uint32_t n_starts;
};

// raw scan-start test of the 40 interpolated samples (reference :768) as a bit mask
// One real modulo, then a running remainder: inc_q16 < 360 deg (the jump threshold keeps the
// step per sample far below a revolution), so a single conditional subtraction per step suffices.
__device__ __forceinline__ unsigned long long raw_sync_mask(int prev_q8, int inc_q16) {
  const int kFull = 360 << 16;
  unsigned long long m = 0;
  int rem = ((prev_q8 << 8) + inc_q16) % kFull;  // (cur + inc) % full for pos = 0
  const int lim = inc_q16 << 1;
  // no wrap inside the capsule and already past the two steps after the last one: nothing to mark
  // (79 of 80 capsules of a revolution)
  if (rem >= lim && rem + 39 * inc_q16 < kFull) return 0ull;
#pragma unroll 8
  for (int pos = 0; pos < 40; ++pos) {
    if (rem < lim) m |= 1ull << pos;
    rem += inc_q16;
    if (rem >= kFull) rem -= kFull;
  }
  return m;
}
// sync_i = raw_i & ~sync_{i-1} (reference :769), sync_{-1} = s_in; only set raw bits matter
__device__ __forceinline__ unsigned long long resolve_sync(unsigned long long raw, uint32_t s_in) {
  unsigned long long s = 0, r = raw;
  while (r) {
    const int i = __ffsll((long long)r) - 1;
    r &= r - 1;
    const uint32_t prev = (i == 0) ? s_in : (uint32_t)((s >> (i - 1)) & 1ull);
    if (!prev) s |= 1ull << i;
  }
  return s;
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}

__global__ void __launch_bounds__(DT) decode_dense_kernel(DecodeArgs a) {
  extern __shared__ __align__(16) unsigned char decode_smem_raw[];
  DecodeSmem& sm = *reinterpret_cast<DecodeSmem*>(decode_smem_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int thr_q8 = (360 * 100 * 40 / (int)(1000000u / a.sample_duration_us)) << 8;

  for (uint32_t s = blockIdx.x; s < a.n_streams; s += gridDim.x) {
    const uint32_t n = a.counts[s];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.capsules + (size_t)s * a.stride_capsules * 84);
    uint2* out = a.nodes_out + (size_t)s * a.stride_capsules * 40;
    uint32_t* st_out = a.capsule_status ? a.capsule_status + (size_t)s * a.stride_capsules : nullptr;
    uint32_t* off_out = a.capsule_node_offset ? a.capsule_node_offset + (size_t)s * a.stride_capsules : nullptr;
    uint32_t* starts = a.scan_starts ? a.scan_starts + (size_t)s * a.starts_stride : nullptr;
    if (tid == 0) {
      sm.n_starts = 0;
      sm.carry_nodes = 0;
      sm.carry_sync = a.sync_state_in ? (a.sync_state_in[s] & 1u) : 0u;
      sm.okflag[0] = 0;  // no previous capsule
      sm.start_q8[0] = 0;
    }
    __syncthreads();

    // stage a tile into buffer `b` (asynchronously when the source is 16-byte aligned)
    auto stage = [&](uint32_t c0, uint32_t b) {
      const uint32_t live = min((uint32_t)DT, n - c0);
      const uint32_t words = live * kCapWords;
      const uint32_t* g = src + (size_t)c0 * kCapWords;
      uint32_t done = 0;
      if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
        const uint32_t quads = words >> 2;
        for (uint32_t q = tid; q < quads; q += DT) cp_async16(&sm.cap[b][4 * q], g + 4 * q);
        done = quads << 2;
      }
      for (uint32_t w = done + tid; w < words; w += DT) sm.cap[b][w] = __ldg(g + w);
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (n > 0) stage(0, 0);
    uint32_t buf = 0;
    for (uint32_t c0 = 0; c0 < n; c0 += DT, buf ^= 1u) {
      const uint32_t live = min((uint32_t)DT, n - c0);
      // ---- wait for this tile, start fetching the next one ------------------------------------------
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      if (c0 + DT < n) stage(c0 + DT, buf ^ 1u);
      const uint32_t* tile = sm.cap[buf];
      // ---- per capsule: frame, checksum, start angle --------------------------------------------
      uint32_t st = 0, ok = 0, sync = 0, start = 0;
      if (tid < live) {
        const uint32_t* c = &tile[tid * kCapWords];
        const uint32_t w0 = c[0];
        const uint32_t b0 = w0 & 0xFF, b1 = (w0 >> 8) & 0xFF;
        start = w0 >> 16;
        if ((b0 >> 4) != 0xA || (b1 >> 4) != 0x5) {
          st = kStBadFrame;
        } else {
          uint32_t x = w0 >> 16;  // bytes 2, 3
#pragma unroll
          for (int w = 1; w < kCapWords; ++w) x ^= c[w];
          x ^= x >> 16;
          const uint32_t sum = (x ^ (x >> 8)) & 0xFF;
          const uint32_t recv = ((b0 & 0xF) | (b1 << 4)) & 0xFF;
          if (recv != sum) {
            st = kStChecksum;
          } else {
            ok = 1;
            st = kStOk;
            sync = (start >> 15) & 1u;
            if (sync) st |= kStSync;
          }
        }
        sm.okflag[tid + 1] = ok;
        sm.start_q8[tid + 1] = (start & 0x7FFFu) << 2;
      }
      __syncthreads();
      // ---- does this capsule release its predecessor? ---------------------------------------------
      uint32_t emit = 0;
      int prev_q8 = 0, inc_q16 = 0;
      if (tid < live && ok) {
        const uint32_t prev_ok = sm.okflag[tid];
        if (sync) {
          if (prev_ok) st |= kStEncReset;
        } else if (prev_ok) {
          const int cur_q8 = (int)sm.start_q8[tid + 1];
          prev_q8 = (int)sm.start_q8[tid];
          int diff = cur_q8 - prev_q8;
          if (prev_q8 > cur_q8) diff += (360 << 8);
          if (diff > thr_q8) {
            st |= kStDiscard;
          } else {
            emit = 1;
            st |= kStEmit;
            inc_q16 = (diff << 8) / 40;
            sm.inc_q16[tid] = inc_q16;
          }
        }
      }
      // exclusive scan of the release flags -> node offsets
      const uint32_t inc_scan = warp_inclusive_scan(emit);
      if (lane == 31) sm.warp_a[warp] = inc_scan;
      // transfer function of the sync bit through this capsule: f(s_in) = s_out
      unsigned long long raw = 0;
      uint32_t f = 0x2;  // identity: f(0)=0 (bit0), f(1)=1 (bit1)
      if (emit) {
        raw = raw_sync_mask(prev_q8, inc_q16);
        const uint32_t o0 = (uint32_t)(resolve_sync(raw, 0) >> 39) & 1u;
        const uint32_t o1 = (uint32_t)(resolve_sync(raw, 1) >> 39) & 1u;
        f = o0 | (o1 << 1);
      }
      // inclusive scan of the functions under composition: (g o f)(x) = g(f(x))
      uint32_t F = f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t p = __shfl_up_sync(0xffffffffu, F, o);  // earlier capsules' function
        if (lane >= (uint32_t)o) F = ((F >> (p & 1u)) & 1u) | (((F >> ((p >> 1) & 1u)) & 1u) << 1);
      }
      if (lane == 31) sm.warp_b[warp] = F;
      __syncthreads();
      uint32_t base_off = 0, s_state = sm.carry_sync;
      for (uint32_t w = 0; w < warp; ++w) {
        base_off += sm.warp_a[w];
        s_state = (sm.warp_b[w] >> s_state) & 1u;
      }
      const uint32_t my_off = base_off + inc_scan - emit;  // releasing capsules before me in the tile
      // state entering this capsule: everything before it in the warp
      const uint32_t Fprev = __shfl_up_sync(0xffffffffu, F, 1);
      const uint32_t s_in = (lane == 0) ? s_state : ((Fprev >> s_state) & 1u);
      if (tid < live) {
        sm.smask[tid] = emit ? resolve_sync(raw, s_in) : 0ull;
        if (emit) sm.emit_list[my_off] = tid;
        const uint32_t node_off = sm.carry_nodes + 40u * my_off;
        if (st_out) st_out[c0 + tid] = st;
        if (off_out) off_out[c0 + tid] = node_off;
        if (starts) {  // the scan-start nodes of this capsule (at most a few per revolution), for the assembler
          unsigned long long m = sm.smask[tid];
          while (m) {
            const int pos = __ffsll((long long)m) - 1;
            m &= m - 1;
            const uint32_t idx = atomicAdd(&sm.n_starts, 1u);
            if (idx < a.starts_stride) starts[idx] = node_off + (uint32_t)pos;
          }
        }
      }
      if (tid == DT - 1) {
        uint32_t tot = 0, st2 = sm.carry_sync;
        for (uint32_t w = 0; w < DT / 32; ++w) {
          tot += sm.warp_a[w];
          st2 = (sm.warp_b[w] >> st2) & 1u;
        }
        sm.tile_nodes = 40u * tot;
        sm.red_sync = st2;
      }
      __syncthreads();
      // ---- write the nodes: the tile's 40 * E nodes are one contiguous run of the output; a thread handles a PAIR of
      // samples (one 32-bit word of the capsule): the capsule look-ups are shared and the two nodes leave as one
      // 16-byte store (a capsule's run starts on a multiple of 320 bytes)
      {
        const uint32_t n_pairs = sm.tile_nodes / 2u;
        uint2* o = out + sm.carry_nodes;
        const bool wide = (reinterpret_cast<uintptr_t>(o) & 15u) == 0;
        auto node_of = [](int angle_q6, int dist_q2, uint32_t syncb) {
          if (angle_q6 < 0) angle_q6 += (360 << 6);
          if (angle_q6 >= (360 << 6)) angle_q6 -= (360 << 6);
          // (angle_q6 >= 0 here for every wire input: start angles are 15-bit q6 values and steps are >= -152 deg)
          const uint32_t key = ((uint32_t)(angle_q6 << 8) / 90u) & 0xFFFFu;
          const uint32_t quality = dist_q2 ? (0x2Fu << 2) : 0u;
          const uint32_t flag = syncb | ((syncb ^ 1u) << 1);
          uint2 nd;
          nd.x = key | ((uint32_t)dist_q2 << 16);
          nd.y = ((uint32_t)dist_q2 >> 16) | (quality << 16) | (flag << 24);
          return nd;
        };
        for (uint32_t p = tid; p < n_pairs; p += DT) {
          const uint32_t e = p / 20u, pp = p - e * 20u, pos = 2u * pp;
          const uint32_t j = sm.emit_list[e];
          const uint32_t* pc = (j == 0) ? sm.carry : &tile[(j - 1) * kCapWords];  // the predecessor capsule
          const int pq8 = (int)sm.start_q8[j];
          const int inc = sm.inc_q16[j];
          const uint32_t wq = pc[1 + pp];
          const uint32_t sy = (uint32_t)(sm.smask[j] >> pos) & 3u;
          const int a0 = (pq8 << 8) + (int)pos * inc;
          const uint2 na = node_of(a0 >> 10, (int)((wq & 0xFFFFu) << 2), sy & 1u);
          const uint2 nb = node_of((a0 + inc) >> 10, (int)((wq >> 16) << 2), sy >> 1);
          if (wide) {
            *reinterpret_cast<uint4*>(o + 2u * p) = make_uint4(na.x, na.y, nb.x, nb.y);
          } else {
            o[2u * p] = na;
            o[2u * p + 1u] = nb;
          }
        }
      }
      __syncthreads();
      // ---- carry into the next tile -----------------------------------------------------------------
      if (tid < kCapWords) sm.carry[tid] = tile[(live - 1) * kCapWords + tid];
      if (tid == 0) {
        sm.okflag[0] = sm.okflag[live];
        sm.start_q8[0] = sm.start_q8[live];
        sm.carry_nodes += sm.tile_nodes;
        sm.carry_sync = sm.red_sync;
      }
      __syncthreads();
    }
    if (tid == 0) {
      if (a.node_counts) a.node_counts[s] = sm.carry_nodes;
      if (a.sync_state_out) a.sync_state_out[s] = sm.carry_sync;
      if (a.scan_start_counts) a.scan_start_counts[s] = sm.n_starts;
    }
    __syncthreads();
  }
}

}  // namespace

cudaError_t launch_decode_dense(const DecodeArgs& a, int grid, cudaStream_t stream) {
  if (a.n_streams == 0) return cudaSuccess;
  decode_dense_kernel<<<grid, DT, sizeof(DecodeSmem), stream>>>(a);
  return cudaGetLastError();
}

cudaError_t decode_configure() {
  return cudaFuncSetAttribute(decode_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)sizeof(DecodeSmem));
}

}  // namespace rpl
