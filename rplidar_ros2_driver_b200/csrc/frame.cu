// frame.cu -- byte-level framing of raw capsule streams with the SDK's resynchronisation.
//
// The capsule unpackers of the SDK do not receive framed capsules: they hunt for the two sync nibbles byte by
// byte (reference src/sdk/src/dataunpacker/unpacker/handler_capsules.cpp:107-135 express, :324-353 ultra,
// :639-668 dense, :852-880 ultra-dense -- the same little machine four times, only the frame size differs):
//   waiting for byte 0:  high nibble 0xA -> byte 1;  anything else is skipped and `_is_previous_capsuledataRdy`
//                        is cleared
//   waiting for byte 1:  high nibble 0x5 -> collect; anything else (that byte is consumed, not looked at again)
//                        goes back to byte 0 and clears the flag as well
//   collecting:          bytes 2 .. size-1 are taken blindly; the frame is then checked (checksum) and decoded
// As a function of the stream this is a chain of jumps from one "waiting for byte 0" position to the next:
// +1, +2 or +frame size.  This kernel walks that chain and writes the frames it visits back to back
// ([capsules][frame size], the input format of decode.cu / decode_formats.cu); every stretch of skipped bytes
// becomes ONE all-zero capsule in the output.  The decoders report such a capsule as RPL_CAPSULE_BAD_FRAME and
// forget the capsule before it -- precisely the effect the skipped bytes have in the SDK -- so framing + decoding
// reproduces the SDK's node stream on damaged input (dropped, inserted, corrupted bytes), pinned against the
// SDK's own unpacker in tests/test_framing_vs_ref.py.
//
// One CTA per stream, tiles of 16 KB staged through shared memory.  A tile whose frames all sit where the chain
// expects them (the normal case) is recognised by one marker test per frame, in parallel, and copied as a block;
// only a tile with a broken marker is walked serially by one thread (a dependent shared-memory load per jump).
#include "decode_args.h"
#include "rpl_device.cuh"

namespace rpl {

namespace {

constexpr int FT = 256;
constexpr uint32_t kTileBytes = 16384;
constexpr uint32_t kMaxFrame = 176;
constexpr uint32_t kMaxEntries = 2 * (kTileBytes / 84) + 8;
constexpr uint32_t kDummy = 0xFFFFFFFFu;

struct FrameSmem {
  uint32_t words[(kTileBytes + kMaxFrame + 8) / 4 + 2];
  uint32_t entry[kMaxEntries];  // tile offset of a frame, or kDummy
  uint32_t n_entries, new_pos, lost, done;
};

__global__ void __launch_bounds__(FT) frame_capsules_kernel(FrameArgs a) {
  __shared__ FrameSmem sm;
  const uint32_t tid = threadIdx.x;
  const uint32_t cb = a.capsule_bytes;
  const uint8_t* sb = reinterpret_cast<const uint8_t*>(sm.words);
  for (uint32_t s = blockIdx.x; s < a.n_streams; s += gridDim.x) {
    const uint8_t* in = a.bytes + (size_t)s * a.stride_bytes;
    const uint32_t n = a.byte_counts[s];
    uint8_t* out = a.capsules_out + (size_t)s * a.stride_capsules * cb;
    uint32_t pos = 0, count = 0;
    bool lost = false, done = false;
    while (!done && pos < n) {
      // ---- stage [pos, pos + L) (word loads from the aligned-down address) -----------------------------
      const uint32_t L = min(kTileBytes + cb, n - pos);
      const uintptr_t addr = reinterpret_cast<uintptr_t>(in + pos);
      const uint32_t sh = (uint32_t)(addr & 3u);
      const uint32_t* gw = reinterpret_cast<const uint32_t*>(addr - sh);
      const uint32_t nwords = (sh + L + 3u) >> 2;
      __syncthreads();  // the previous tile's readers are done
      for (uint32_t w = tid; w < nwords; w += FT) sm.words[w] = __ldg(gw + w);
      __syncthreads();
      const uint8_t* t = sb + sh;  // t[i] = stream byte pos + i
      // ---- fast path: every frame of the tile where the chain expects it --------------------------------
      const uint32_t K = min(kTileBytes / cb, L / cb);
      int ok = 1;
      for (uint32_t j = tid; j < K; j += FT)
        ok = ok && ((t[j * cb] >> 4) == 0xAu) && ((t[j * cb + 1] >> 4) == 0x5u);
      const int all_ok = __syncthreads_and(ok);
      if (K > 0 && all_ok) {
        const uint32_t first = count + (lost ? 1u : 0u);
        const uint32_t room = first < a.stride_capsules ? a.stride_capsules - first : 0u;
        const uint32_t kw = min(K, room);
        if (lost && count < a.stride_capsules)
          for (uint32_t i = tid; i < cb; i += FT) out[(size_t)count * cb + i] = 0;
        uint8_t* dst = out + (size_t)first * cb;
        const uint32_t bytes = kw * cb;
        if (sh == 0 && (reinterpret_cast<uintptr_t>(dst) & 3u) == 0 && (bytes & 3u) == 0) {
          uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
          for (uint32_t w = tid; w < (bytes >> 2); w += FT) d4[w] = sm.words[w];
        } else {
          for (uint32_t i = tid; i < bytes; i += FT) dst[i] = t[i];
        }
        count = first + K;
        lost = false;
        pos += K * cb;
        continue;
      }
      // ---- a marker is broken (or the stream ends inside this tile): walk the chain ------------------------
      if (tid == 0) {
        uint32_t q = 0, ne = 0, l = lost ? 1u : 0u, fin = 0;
        const uint32_t lim = min(kTileBytes, L);
        while (q < lim) {
          if ((t[q] >> 4) != 0xAu) {
            l = 1;
            q += 1;
            continue;
          }
          if (q + 1 >= L) {  // the stream ends on a lone first marker byte
            fin = 1;
            break;
          }
          if ((t[q + 1] >> 4) != 0x5u) {
            l = 1;
            q += 2;
            continue;
          }
          if (q + cb > L) {  // unfinished frame at the end of the stream
            fin = 1;
            break;
          }
          if (l) {
            sm.entry[ne++] = kDummy;
            l = 0;
          }
          sm.entry[ne++] = q;
          q += cb;
        }
        // L < tile + frame means the stream ends in this tile: whatever is left over is an unfinished frame
        if (!fin && pos + q >= n) fin = 1;
        sm.n_entries = ne;
        sm.new_pos = pos + min(q, L);
        sm.lost = l;
        sm.done = fin;
      }
      __syncthreads();
      const uint32_t ne = sm.n_entries;
      for (uint32_t e = 0; e < ne; ++e) {
        if (count + e >= a.stride_capsules) break;
        const uint32_t off = sm.entry[e];
        uint8_t* dst = out + (size_t)(count + e) * cb;
        if (off == kDummy) {
          for (uint32_t i = tid; i < cb; i += FT) dst[i] = 0;
        } else {
          for (uint32_t i = tid; i < cb; i += FT) dst[i] = t[off + i];
        }
      }
      count += ne;
      pos = sm.new_pos;
      lost = sm.lost != 0;
      done = sm.done != 0;
    }
    if (tid == 0) {
      a.capsule_counts_out[s] = count;  // > stride_capsules: the output overflowed (frames past it were dropped)
      if (a.bytes_left_out) a.bytes_left_out[s] = n - min(pos, n);
    }
    __syncthreads();
  }
}

}  // namespace

cudaError_t launch_frame_capsules(const FrameArgs& a, int grid, cudaStream_t stream) {
  if (a.n_streams == 0) return cudaSuccess;
  frame_capsules_kernel<<<grid, FT, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace rpl
