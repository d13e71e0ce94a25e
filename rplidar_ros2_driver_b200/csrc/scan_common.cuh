// scan_common.cuh -- device code shared by the two fast scan kernels (scan_fast.cu: register-
// streamed v1, scan_tma.cu: TMA-ring v2): rank lookup in the key bitmap and the Mode A
// bin-ownership logic.
#pragma once
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
// bin keeps the smallest dist_m (strict '<': on equal dist_m the first point in ascending
// key order).  Bins grow with the key (for inverted scans: key 0 first, then descending
// keys), so the points of a bin are neighbours in that order and the presence bitmap alone
// tells a point whether it is the first (head) / last (tail) of its bin and which empty bins
// lie before it.  Single-point bins are written directly; shared bins go through a small
// per-CTA scratch and are resolved after the pass by the warp that saw their head (every warp
// keeps its own head list and count: no shared counter, no atomics).
struct ModeACtx {
  const uint2* rankV;
  float* ranges;
  float* intens;
  unsigned long long* gscratch;
  uint32_t* fallback;
  uint32_t M;
  float inc;
  bool inverted, has0, new_proto;
};

// Returns 0 for nothing left to do, 1 when this point is alone in bin b_out (the caller stores
// it with one converged, coalesced warp store), 2 when it heads a bin shared by several
// points ((ru_out, b_out) then name the group's first scratch slot and its bin).
__device__ __forceinline__ uint32_t mode_a_place(const ModeACtx& c, uint32_t k, uint32_t r, float dm,
                                                 uint32_t q, uint32_t& ru_out, uint32_t& b_out) {
  const uint32_t M = c.M;
  const float kInf = __int_as_float(0x7f800000);
  int pk, nk;
  uint32_t ru;
  if (!c.inverted) {
    pk = prev_set(c.rankV, k);
    nk = next_set(c.rankV, k);
    ru = r;
  } else if (k == 0) {
    pk = -1;
    nk = (c.rankV[kWords - 1].x >> 31) ? (int)(kKeySpace - 1) : prev_set(c.rankV, kKeySpace - 1);
    if (nk == 0) nk = -1;
    ru = 0;
  } else {
    pk = next_set(c.rankV, k);
    if (pk < 0 && c.has0) pk = 0;
    nk = prev_set(c.rankV, k);
    if (nk == 0) nk = -1;  // key 0 comes first in the inverted order, never after
    ru = (M - 1 - r) + (c.has0 ? 1u : 0u);
  }
  const int b = mode_a_bin_fast(k, M, c.inc, c.inverted);
  if (b < 0 || b >= (int)M) {  // never taken for u16 keys; the reference's guard, kept
    *c.fallback = 1;
    return 0u;
  }
  const int bp = pk >= 0 ? mode_a_bin_fast((uint32_t)pk, M, c.inc, c.inverted) : -1;
  const int bn = nk >= 0 ? mode_a_bin_fast((uint32_t)nk, M, c.inc, c.inverted) : (int)M;
  const bool head = bp != b, tail = bn != b;
  if (head)
    for (int e = bp + 1; e < b; ++e) {  // empty bins in front of this group
      c.ranges[e] = kInf;
      c.intens[e] = 0.0f;
    }
  if (nk < 0)
    for (int e = b + 1; e < (int)M; ++e) {  // empty bins behind the last group
      c.ranges[e] = kInf;
      c.intens[e] = 0.0f;
    }
  b_out = (uint32_t)b;
  if (head && tail) return 1u;
  // several points share the bin: keep the smallest (dist_m, key)
  c.gscratch[ru] = ((unsigned long long)__float_as_uint(dm) << 32) | ((unsigned long long)k << 16) |
                   ((unsigned long long)q << 8) | (tail ? 1ull : 0ull);
  ru_out = ru;
  return head ? 2u : 0u;
}

// warp-private list of shared-bin heads: converged call, one entry per lane with `is_head`
__device__ __forceinline__ void mode_a_push_heads(uint2* wlist, uint32_t cap, uint32_t& wcount, bool is_head,
                                                  uint32_t ru, uint32_t b, uint32_t* fallback) {
  const uint32_t m = __ballot_sync(0xffffffffu, is_head);
  if (m == 0) return;
  const uint32_t lane = threadIdx.x & 31u;
  if (is_head) {
    const uint32_t slot = wcount + __popc(m & ((1u << lane) - 1u));
    if (slot < cap) wlist[slot] = make_uint2(ru, b);
    else *fallback = 1;
  }
  wcount += __popc(m);
}

// resolver: the minimum (dist_m, key) of every listed group -> ranges / intensities
__device__ __forceinline__ void mode_a_resolve(const ModeACtx& c, const uint2* wlist, uint32_t wcount) {
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t e = lane; e < wcount; e += 32) {
    const uint2 h = wlist[e];
    unsigned long long best = ~0ull;
    for (uint32_t slot = h.x; slot < c.M; ++slot) {
      const unsigned long long g = c.gscratch[slot];
      best = min(best, g);
      if (g & 1ull) break;
    }
    c.ranges[h.y] = __uint_as_float((uint32_t)(best >> 32));
    c.intens[h.y] = quality_to_intensity((uint32_t)(best >> 8) & 0xFFu, c.new_proto);
  }
}

}  // namespace
}  // namespace rpl
