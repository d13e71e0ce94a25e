// scan_common.cuh -- device code shared by the two fast scan kernels (scan_fast.cu: register-
// streamed v1, scan_tma.cu: TMA-ring v2): rank lookup in the key bitmap and the Mode A
// bin-ownership logic.
#pragma once
#include <type_traits>

#include "rpl_device.cuh"

namespace rpl {
namespace {

constexpr uint32_t kWords = kKeySpace / 32;     // 2048 bitmap words
constexpr uint32_t kMaxFastNodes = kKeySpace;   // more nodes cannot be tie-free

__device__ __forceinline__ uint32_t rank_of(const uint2* rk, uint32_t key) {
  const uint2 e = rk[key >> 5];
  return e.y + __popc(e.x & ((1u << (key & 31)) - 1u));
}
// largest set key < k (or -1) / smallest set key > k (or -1)
__device__ __forceinline__ int prev_set(const uint2* rk, uint32_t k) {
  int w = (int)(k >> 5);
  uint32_t m = rk[w].x & ((1u << (k & 31)) - 1u);
  for (;;) {
    if (m) return (w << 5) + 31 - __clz(m);
    if (w == 0) return -1;
    m = rk[--w].x;
  }
}
__device__ __forceinline__ int next_set(const uint2* rk, uint32_t k) {
  uint32_t w = k >> 5;
  uint32_t m = rk[w].x & ~((2u << (k & 31)) - 1u);
  for (;;) {
    if (m) return (int)((w << 5) + __ffs(m) - 1);
    if (++w == kWords) return -1;
    m = rk[w].x;
  }
}

// predicated streaming stores (no branch around them)
__device__ __forceinline__ void st_f32_if(float* p, float v, uint64_t pol, uint32_t pred) {
  asm volatile(
      "{ .reg .pred q; setp.ne.u32 q, %3, 0;\n\t"
      "@q st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2; }" ::"l"(p),
      "f"(v), "l"(pol), "r"(pred));
}

// ---- Mode A (reference rplidar_node.cpp:630-660) ------------------------------------------
// beam_count = M bins; every measured point goes to bin (int)(angle / angle_increment) and the
// bin keeps the smallest dist_m (strict '<': on equal dist_m the first point in ascending key
// order, i.e. the smallest key).
//
// Bins grow with the key -- for inverted scans along "key 0 first, then descending keys" -- so in
// that order (the "u-order") the points of a bin are neighbours.  The place pass therefore
// writes one packed entry per measured point at its u-rank into a per-CTA scratch
//     entry = dist_m bits << 32 | key << 8 | quality          (u64 min = min (dist_m, key))
// and mode_a_emit() walks the entries in order: every warp owns a contiguous slice, finds the
// runs of equal bins, takes their minimum with a segmented shuffle scan, and writes each bin
// exactly once (empty bins included) with converged, mostly coalesced stores.  No atomics, no
// block-wide barriers inside the walk.
__device__ __forceinline__ uint32_t mode_a_urank(uint32_t key, uint32_t rank, uint32_t M, bool inverted,
                                                 bool has0) {
  if (!inverted) return rank;
  if (key == 0u) return 0u;
  return (M - 1u - rank) + (has0 ? 1u : 0u);
}
__device__ __forceinline__ unsigned long long mode_a_entry(float dist_m, uint32_t key, uint32_t quality) {
  return ((unsigned long long)__float_as_uint(dist_m) << 32) | ((unsigned long long)key << 8) |
         (unsigned long long)quality;
}

struct ModeAOut {
  float* ranges;
  float* intens;
  const float2* angle;  // [65536] (angle, inverted angle) table
  uint32_t M;
  float inc;
  bool inverted, new_proto;
  uint64_t policy;
};

__device__ __forceinline__ void mode_a_store_bin(const ModeAOut& o, int bin, unsigned long long best, uint32_t pred) {
  const float dm = __uint_as_float((uint32_t)(best >> 32));
  const float it = quality_to_intensity((uint32_t)best & 0xFFu, o.new_proto);
  st_f32_if(o.ranges + bin, dm, o.policy, pred);
  st_f32_if(o.intens + bin, it, o.policy, pred);
}
__device__ __forceinline__ void mode_a_fill_empty(const ModeAOut& o, int from, int to) {  // bins [from, to)
  const float kInf = __int_as_float(0x7f800000);
  for (int e = from; e < to; ++e) {
    o.ranges[e] = kInf;
    o.intens[e] = 0.0f;
  }
}

// One warp of `nwarps`; E = entries in u-order.  Call with all 32 lanes converged.
//
// Ownership: a run (maximal stretch of equal bins) belongs to the warp whose slice holds its
// first entry.  The owner writes the run's bin and the empty bins in front of it; the owner of
// the last run also writes the empty bins behind it.  A warp therefore may read past the end of
// its slice to finish a run it owns, and skips leading entries that continue a run of the
// previous slice.
constexpr uint32_t kEmitBatch = 256;                 // entries a warp stages per batch
constexpr uint32_t kEmitStage = kEmitBatch + 33;     // + the entry before and 32 after (look-ahead)
constexpr uint32_t kEmitStageBytes = (kEmitStage * 12 + 15) & ~15u;  // u64 entry + i32 bin per staged entry

// Every entry is handled by one lane, independently of all others: the lane that holds the
// first entry of a run (bin differs from the previous entry's) takes the run's minimum by
// looking ahead, writes the bin, and fills the empty bins in front of it; the lane holding the
// very last entry fills the empty bins behind it.  A warp stages 256 entries (+ one before,
// 32 after) and their bins in its private slice of shared memory, so all loads and bin
// evaluations of a batch are in flight together and the per-entry work reads shared memory.
// `stage`: kEmitStageBytes of shared memory private to this warp.  Call converged.
__device__ __forceinline__ void mode_a_emit(const ModeAOut& o, const unsigned long long* E, uint32_t warp,
                                            uint32_t nwarps, unsigned char* stage) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t M = o.M;
  const uint32_t len = ((M + nwarps - 1u) / nwarps + kEmitBatch - 1u) & ~(kEmitBatch - 1u);  // whole batches
  const uint32_t r_begin = min(M, warp * len), r_end = min(M, r_begin + len);
  if (r_begin >= r_end) return;
  unsigned long long* se = reinterpret_cast<unsigned long long*>(stage);
  int* sb = reinterpret_cast<int*>(stage + kEmitStage * 8);
  // (int)((angle - angle_min) / angle_increment) with the reference's own float angle from the
  // table: exact, no double precision, no divergence
  auto bin_of = [&](unsigned long long e) {
    const float2 a = __ldg(o.angle + ((uint32_t)(e >> 8) & 0xFFFFu));
    return __float2int_rz(__fdiv_rn(o.inverted ? a.y : a.x, o.inc));
  };
  const int kNoBin = 0x7fffffff;
  const float kInf = __int_as_float(0x7f800000);

  for (uint32_t base = r_begin; base < r_end; base += kEmitBatch) {
    __syncwarp();
    for (uint32_t t = lane; t < kEmitStage; t += 32) {  // staged index t <-> rank base - 1 + t
      const long long r = (long long)base - 1 + t;
      unsigned long long e = ~0ull;
      int b = (r < 0) ? -1 : kNoBin;
      if (r >= 0 && r < (long long)M) {
        e = E[r];
        b = bin_of(e);
      }
      se[t] = e;
      sb[t] = b;
    }
    __syncwarp();
#pragma unroll 2
    for (uint32_t w = 0; w < kEmitBatch / 32; ++w) {
      const uint32_t i = 1 + w * 32 + lane;
      const uint32_t r = base + w * 32 + lane;
      const bool live = r < M;
      const int b = sb[i], bprev = sb[i - 1];
      const bool head = live && (b != bprev);
      unsigned long long v = se[i];
      if (head && sb[i + 1] == b) {
        v = min(v, se[i + 1]);
        if (sb[i + 2] == b) {  // a run of three or more: walk it (rare with M points in M bins)
          uint32_t j = i + 2;
          while (j < kEmitStage && sb[j] == b) {
            v = min(v, se[j]);
            ++j;
          }
          if (j == kEmitStage) {
            for (uint32_t rr = base - 1 + j; rr < M; ++rr) {
              const unsigned long long ee = E[rr];
              if (bin_of(ee) != b) break;
              v = min(v, ee);
            }
          }
        }
      }
      mode_a_store_bin(o, head ? b : 0, v, head ? 1u : 0u);
      if (head && b - bprev > 1) {  // empty bins in front of this run
        if (b - bprev == 2) {
          o.ranges[b - 1] = kInf;
          o.intens[b - 1] = 0.0f;
        } else {
          mode_a_fill_empty(o, bprev + 1, b);
        }
      }
      if (live && r == M - 1u) mode_a_fill_empty(o, b + 1, (int)M);  // empty bins behind the last run
    }
  }
}

// ---- Mode A emit, shared-memory variant (scan_tma.cu) ----------------------------------------
// ncu on the variant above showed the scratch doubling DRAM traffic (8 B written + 8 B read per
// point through a 76 MB working set).  Here the place pass only records WHICH node sits at each
// u-rank, as a u16 node index in the 64 KB the dead presence map leaves free (so M <= 32768),
// and the emit pass gathers the nodes from the tile again (L2 hits, coalesced for a sorted
// revolution).  Bins of a batch are staged per warp in the (equally dead) rank table.
constexpr uint32_t kEmit2Batch = 256;
constexpr uint32_t kEmit2Stage = kEmit2Batch + 33;  // staged bins: one entry before, 32 after
constexpr uint32_t kModeASmemMaxPoints = kKeySpace / 2;

__device__ __forceinline__ void mode_a_emit_smem(const ModeAOut& o, const uint16_t* sidx, const uint2* tile,
                                                 uint32_t warp, uint32_t nwarps, uint16_t* sb) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t M = o.M;
  const uint32_t len = ((M + nwarps - 1u) / nwarps + kEmit2Batch - 1u) & ~(kEmit2Batch - 1u);
  const uint32_t r_begin = min(M, warp * len), r_end = min(M, r_begin + len);
  if (r_begin >= r_end) return;
  const bool inverted = o.inverted;
  const float inc = o.inc;
  // integer quotient where the float chain provably agrees, else the exact chain in registers
  // (FP64, no memory access): this pass is bound by load latency, not by issue slots, so the
  // table lookup of mode_a_emit would only add a dependent L2 access per entry
  auto bin_of_key = [&](uint32_t key) { return (uint32_t)mode_a_bin_fast(key, M, inc, inverted); };
  const uint32_t kNoBin = 0xFFFFu, kBeforeFirst = 0xFFFEu;  // real bins are < 32768
  const float kInf = __int_as_float(0x7f800000);
  auto entry_of = [&](uint2 nd) {
    return mode_a_entry(dist_to_m(__funnelshift_r(nd.x, nd.y, 16)), nd.x & 0xFFFFu, (nd.y >> 16) & 0xFFu);
  };
  constexpr int kW = kEmit2Batch / 32;
  using Checked = std::integral_constant<bool, true>;
  using Unchecked = std::integral_constant<bool, false>;

  auto batch = [&](auto checked, uint32_t base) {
    constexpr bool CK = decltype(checked)::value;
    __syncwarp();
    // stage: this lane's own eight entries (kept in registers) plus one halo entry (the entry
    // before the batch for lane 0, the 32 after it for the others); all gathers issued together
    uint2 nd[kW];
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      const uint32_t r = base + w * 32 + lane;
      nd[w] = (!CK || r < M) ? tile[sidx[r]] : make_uint2(0, 0);
    }
    const uint32_t th = (lane == 0) ? 0u : (kEmit2Batch + lane);  // lane 0 -> before; 1..31 -> after
    const long long rh = (long long)base - 1 + th;
    const bool halo_ok = !CK || (rh >= 0 && rh < (long long)M);
    const uint2 ndh = halo_ok ? tile[sidx[halo_ok ? rh : 0]] : make_uint2(0, 0);
    const uint32_t r2 = base - 1 + kEmit2Batch + 32;  // last look-ahead slot (lane 31)
    const bool last_ok = (lane == 31) && (!CK || r2 < M);
    const uint2 nd2 = last_ok ? tile[sidx[last_ok ? r2 : 0]] : make_uint2(0, 0);
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      const uint32_t r = base + w * 32 + lane;
      sb[1 + w * 32 + lane] = (!CK || r < M) ? (uint16_t)bin_of_key(nd[w].x & 0xFFFFu) : (uint16_t)kNoBin;
    }
    {
      uint32_t b = (CK && rh < 0) ? kBeforeFirst : kNoBin;
      if (halo_ok) b = bin_of_key(ndh.x & 0xFFFFu);
      sb[th] = (uint16_t)b;
      if (lane == 31) sb[kEmit2Batch + 32] = last_ok ? (uint16_t)bin_of_key(nd2.x & 0xFFFFu) : (uint16_t)kNoBin;
    }
    __syncwarp();
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      const uint32_t i = 1 + w * 32 + lane;
      const uint32_t r = base + w * 32 + lane;
      const bool live = !CK || r < M;
      const uint32_t b = sb[i], bprev = sb[i - 1];
      const bool head = live && (b != bprev);
      unsigned long long v = entry_of(nd[w]);
      // the next entry's node sits in the neighbouring lane (or lane 0 of the next window):
      // a bin shared by two points -- the usual collision -- costs two shuffles, no memory access
      uint2 nxt;
      nxt.x = __shfl_down_sync(0xffffffffu, nd[w].x, 1);
      nxt.y = __shfl_down_sync(0xffffffffu, nd[w].y, 1);
      if (w + 1 < kW) {
        const uint32_t fx = __shfl_sync(0xffffffffu, nd[(w + 1) % kW].x, 0);
        const uint32_t fy = __shfl_sync(0xffffffffu, nd[(w + 1) % kW].y, 0);
        if (lane == 31) nxt = make_uint2(fx, fy);
      }
      if (head && sb[i + 1] == b) {
        if (w + 1 == kW && lane == 31) nxt = tile[sidx[r + 1]];  // first entry of the next batch
        v = min(v, entry_of(nxt));
        if (sb[i + 2] == b) {  // three or more points in the bin: walk on (rare: M points, M bins)
          for (uint32_t rr = r + 2; rr < M; ++rr) {
            const uint2 other = tile[sidx[rr]];
            if (bin_of_key(other.x & 0xFFFFu) != b) break;
            v = min(v, entry_of(other));
          }
        }
      }
      mode_a_store_bin(o, head ? (int)b : 0, v, head ? 1u : 0u);
      const int gap_from = (CK && bprev == kBeforeFirst) ? 0 : (int)bprev + 1;
      if (head && (int)b > gap_from) {  // empty bins in front of this run
        if ((int)b - gap_from == 1) {
          o.ranges[gap_from] = kInf;
          o.intens[gap_from] = 0.0f;
        } else {
          mode_a_fill_empty(o, gap_from, (int)b);
        }
      }
      if (CK && live && r == M - 1u) mode_a_fill_empty(o, (int)b + 1, (int)M);  // empty bins behind the last run
    }
  };
  for (uint32_t base = r_begin; base < r_end; base += kEmit2Batch) {
    // interior batches (entry before and all 256 + 32 look-ahead entries exist) skip every range check
    if (base >= 1u && base + kEmit2Batch + 32u <= M) batch(Unchecked{}, base);
    else batch(Checked{}, base);
  }
}

}  // namespace
}  // namespace rpl
