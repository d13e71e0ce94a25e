// cloud.cu -- PointCloud2 post-processing (north-star extensions; no reference counterpart).
//
// oracle/cloud_oracle.cpp is the definition (PARITY UNPINNED: the reference has no
// polar->Cartesian, outlier or voxel code).  Steps 1-3 (window, stable key order, polar->xyz)
// run inside the scan kernels (scan_tma.cu / scan_general.cu, `xyzi` payload) and leave, per
// scan, the kept points in angle order as (x, y, 0, intensity).  This file holds
//   * the trig table those kernels read: (float)cos((double)angle_rad(key)), built on the host
//     with the same libm the oracle uses, so x = r * c is bit-identical on both sides,
//   * step 4, statistical outlier removal over the +-16 angular neighbours, statistics taken
//     with exact integer sums (order independent => bit reproducible),
//   * step 5, the voxel grid: shared hash table per scan, centroids from exact integer sums,
//     cells emitted in order of their first member,
//   * the fuse step that packs the per-scan clouds of a batch into one dense cloud (the per-GPU
//     payload of the multi-GPU all-gather).
#include <cmath>
#include <vector>

#include "cloud_args.h"
#include "rpl_device.cuh"

namespace rpl {

namespace {

constexpr int CT = 256;  // threads per CTA of the post-processing kernels
constexpr int kHalfWindow = 16;

struct CloudScratch {  // per CTA, sized for max_nodes
  unsigned long long* q;  // [max_nodes] SOR: fixed-point mean neighbour distance (llrintf)
  uint32_t* slot;      // [max_nodes] voxel: hash slot of every point
  unsigned long long* hkey;  // [hsize] packed cell (EMPTY = ~0)
  long long* hsx;            // [hsize]
  long long* hsy;
  unsigned long long* hsi;
  uint32_t* hn;
  uint32_t* hfirst;
  uint32_t* horder;
};

__device__ __forceinline__ CloudScratch carve(void* base, size_t per_cta, uint32_t max_nodes, uint32_t hsize) {
  unsigned char* p = static_cast<unsigned char*>(base) + (size_t)blockIdx.x * per_cta;
  CloudScratch s;
  s.hkey = reinterpret_cast<unsigned long long*>(p); p += (size_t)hsize * 8;
  s.hsx = reinterpret_cast<long long*>(p); p += (size_t)hsize * 8;
  s.hsy = reinterpret_cast<long long*>(p); p += (size_t)hsize * 8;
  s.hsi = reinterpret_cast<unsigned long long*>(p); p += (size_t)hsize * 8;
  s.hn = reinterpret_cast<uint32_t*>(p); p += (size_t)hsize * 4;
  s.hfirst = reinterpret_cast<uint32_t*>(p); p += (size_t)hsize * 4;
  s.horder = reinterpret_cast<uint32_t*>(p); p += (size_t)hsize * 4;
  s.q = reinterpret_cast<unsigned long long*>(p); p += (size_t)max_nodes * 8;
  s.slot = reinterpret_cast<uint32_t*>(p);
  return s;
}

__host__ __device__ inline uint32_t hash_size_for(uint32_t max_nodes) {
  uint32_t h = 64;
  while (h < 2u * max_nodes) h <<= 1;
  return h;
}

__device__ __forceinline__ uint32_t block_exclusive_scan_u32(uint32_t v, uint32_t* smem_warp, uint32_t* total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t inc = warp_inclusive_scan(v);
  if (lane == 31) smem_warp[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < CT / 32; ++w) {
    const uint32_t t = smem_warp[w];
    if ((uint32_t)w < warp) base += t;
    tot += t;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

// ---- step 4: statistical outlier removal (oracle/cloud_oracle.cpp step 4) -----------------
__global__ void __launch_bounds__(CT) cloud_sor_kernel(float4* xyzi, uint32_t* point_counts, uint32_t n_scans,
                                                       uint32_t stride, uint32_t sor_k, float sor_alpha,
                                                       void* scratch, size_t per_cta, uint32_t max_nodes) {
  __shared__ uint32_t s_warp[CT / 32];
  __shared__ long long s_s1[CT / 32];
  __shared__ unsigned long long s_s2[CT / 32];
  __shared__ double s_thr;
  const CloudScratch sc = carve(scratch, per_cta, max_nodes, hash_size_for(max_nodes));
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (uint32_t s = blockIdx.x; s < n_scans; s += gridDim.x) {
    float4* pts = xyzi + (size_t)s * stride;
    const uint32_t m = point_counts[s];
    if (m < 2 || m > max_nodes) continue;  // fewer than 2 points keep everything
    const bool all_others = (m - 1) <= 2u * kHalfWindow;
    long long s1 = 0;
    unsigned long long s2 = 0;
    for (uint32_t i = tid; i < m; i += CT) {
      const float4 me = pts[i];
      float d[2 * kHalfWindow];
      uint32_t nd = 0;
      auto add = [&](uint32_t j) {
        const float4 o = pts[j];
        const float dx = __fsub_rn(o.x, me.x), dy = __fsub_rn(o.y, me.y);
        d[nd++] = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      };
      if (all_others) {
        for (uint32_t j = 0; j < m; ++j)
          if (j != i) add(j);
      } else {
        for (uint32_t o = 1; o <= (uint32_t)kHalfWindow; ++o) {
          add(i >= o ? i - o : i + m - o);
          add(i + o < m ? i + o : i + o - m);
        }
      }
      // ascending insertion sort (values are distances: no NaN), then the k smallest in order
      for (uint32_t x = 1; x < nd; ++x) {
        const float v = d[x];
        uint32_t y = x;
        while (y > 0 && d[y - 1] > v) {
          d[y] = d[y - 1];
          --y;
        }
        d[y] = v;
      }
      const uint32_t k = min(sor_k, nd);
      float sum = 0.0f;
      for (uint32_t t = 0; t < k; ++t) sum = __fadd_rn(sum, d[t]);
      const float mean = __fdiv_rn(sum, __uint2float_rn(k));
      const long long q = __float2ll_rn(__fmul_rn(mean, 65536.0f));  // llrintf
      sc.q[i] = (unsigned long long)q;
      s1 += q;
      s2 += (unsigned long long)q * (unsigned long long)q;
    }
    // exact block sums
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
      s_s1[warp] = s1;
      s_s2[warp] = s2;
    }
    __syncthreads();
    if (tid == 0) {
      long long t1 = 0;
      unsigned long long t2 = 0;
      for (int w = 0; w < CT / 32; ++w) {
        t1 += s_s1[w];
        t2 += s_s2[w];
      }
      const double dn = (double)m;
      const double mean = __ddiv_rn((double)t1, dn);
      const double sq = __ddiv_rn(__dmul_rn((double)t1, (double)t1), dn);
      double var = __ddiv_rn(__dsub_rn((double)t2, sq), __dsub_rn(dn, 1.0));
      if (!(var > 0.0)) var = 0.0;
      s_thr = __dadd_rn(mean, __dmul_rn((double)sor_alpha, __dsqrt_rn(var)));
    }
    __syncthreads();
    const double thr = s_thr;
    // in-place stable compaction, chunk by chunk (destinations never pass sources)
    uint32_t done = 0;
    for (uint32_t c0 = 0; c0 < m; c0 += CT) {
      const uint32_t i = c0 + tid;
      float4 me = make_float4(0, 0, 0, 0);
      uint32_t keep = 0;
      if (i < m) {
        me = pts[i];
        keep = ((double)(long long)sc.q[i] <= thr) ? 1u : 0u;
      }
      uint32_t tot = 0;
      const uint32_t pos = block_exclusive_scan_u32(keep, s_warp, &tot);  // syncs: chunk fully read
      if (keep) pts[done + pos] = me;
      done += tot;
      __syncthreads();
    }
    if (tid == 0) point_counts[s] = done;
    __syncthreads();
  }
}

// ---- step 5: voxel grid (oracle/cloud_oracle.cpp step 5) -----------------------------------
__global__ void __launch_bounds__(CT) cloud_voxel_kernel(float4* xyzi, uint32_t* point_counts, uint32_t n_scans,
                                                         uint32_t stride, float voxel, void* scratch, size_t per_cta,
                                                         uint32_t max_nodes) {
  __shared__ uint32_t s_warp[CT / 32];
  const uint32_t hsize = hash_size_for(max_nodes);
  const CloudScratch sc = carve(scratch, per_cta, max_nodes, hsize);
  const uint32_t tid = threadIdx.x;
  const unsigned long long kEmpty = ~0ull;
  for (uint32_t s = blockIdx.x; s < n_scans; s += gridDim.x) {
    float4* pts = xyzi + (size_t)s * stride;
    const uint32_t m = point_counts[s];
    if (m == 0 || m > max_nodes) continue;
    uint32_t hs = 64;  // table for this scan: >= 2 m slots
    while (hs < 2u * m) hs <<= 1;
    for (uint32_t j = tid; j < hs; j += CT) {
      sc.hkey[j] = kEmpty;
      sc.hsx[j] = 0;
      sc.hsy[j] = 0;
      sc.hsi[j] = 0;
      sc.hn[j] = 0;
      sc.hfirst[j] = 0xFFFFFFFFu;
    }
    __syncthreads();
    // insert: exact integer sums, first member by atomicMin
    for (uint32_t i = tid; i < m; i += CT) {
      const float4 p = pts[i];
      const int ix = __float2int_rd(__fdiv_rn(p.x, voxel));  // floorf(x / voxel)
      const int iy = __float2int_rd(__fdiv_rn(p.y, voxel));
      const unsigned long long key = ((unsigned long long)(uint32_t)ix << 32) | (uint32_t)iy;
      uint32_t h = (uint32_t)(mix64(key) & (hs - 1));
      for (;;) {
        const unsigned long long prev = atomicCAS(&sc.hkey[h], kEmpty, key);
        if (prev == kEmpty || prev == key) break;
        h = (h + 1) & (hs - 1);
      }
      sc.slot[i] = h;
      atomicAdd(reinterpret_cast<unsigned long long*>(&sc.hsx[h]),
                (unsigned long long)__float2ll_rn(__fmul_rn(p.x, 65536.0f)));
      atomicAdd(reinterpret_cast<unsigned long long*>(&sc.hsy[h]),
                (unsigned long long)__float2ll_rn(__fmul_rn(p.y, 65536.0f)));
      atomicAdd(&sc.hsi[h], (unsigned long long)(long long)__float2ll_rn(p.w));
      atomicAdd(&sc.hn[h], 1u);
      atomicMin(&sc.hfirst[h], i);
    }
    __syncthreads();
    // cells in order of their first member: exclusive scan over "i is a first member"
    uint32_t done = 0;
    for (uint32_t c0 = 0; c0 < m; c0 += CT) {
      const uint32_t i = c0 + tid;
      uint32_t rep = 0, h = 0;
      if (i < m) {
        h = sc.slot[i];
        rep = (sc.hfirst[h] == i) ? 1u : 0u;
      }
      uint32_t tot = 0;
      const uint32_t pos = block_exclusive_scan_u32(rep, s_warp, &tot);
      if (rep) sc.horder[h] = done + pos;
      done += tot;
    }
    __syncthreads();
    // emit centroids (all reads of the point array happened before the first barrier above)
    for (uint32_t i = tid; i < m; i += CT) {
      const uint32_t h = sc.slot[i];
      if (sc.hfirst[h] != i) continue;
      const double cnt = (double)sc.hn[h];
      const double den = __dmul_rn(65536.0, cnt);
      float4 o;
      o.x = __double2float_rn(__ddiv_rn((double)sc.hsx[h], den));
      o.y = __double2float_rn(__ddiv_rn((double)sc.hsy[h], den));
      o.z = 0.0f;
      o.w = __double2float_rn(__ddiv_rn((double)(long long)sc.hsi[h], cnt));
      pts[sc.horder[h]] = o;
    }
    if (tid == 0) point_counts[s] = done;
    __syncthreads();
  }
}

// ---- fuse: per-scan clouds -> one dense cloud ------------------------------------------------
__global__ void cloud_offsets_kernel(const uint32_t* counts, uint32_t n, uint32_t* offsets, uint32_t* total) {
  // single CTA exclusive scan (n_scans is small next to the point data)
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t c0 = 0; c0 < n; c0 += blockDim.x) {
    const uint32_t i = c0 + threadIdx.x;
    const uint32_t v = i < n ? counts[i] : 0u;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_inclusive_scan(v);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    for (uint32_t w = 0; w < blockDim.x / 32; ++w) {
      if (w < warp) base += s_warp[w];
      tot += s_warp[w];
    }
    const uint32_t carry = s_carry;
    if (i < n) offsets[i] = carry + base + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

__global__ void cloud_pack_kernel(const float4* xyzi, const uint32_t* counts, const uint32_t* offsets,
                                  uint32_t n_scans, uint32_t stride, float4* fused) {
  const uint64_t pol = l2_policy_evict_first();
  for (uint32_t s = blockIdx.y; s < n_scans; s += gridDim.y) {
    const uint32_t m = counts[s], off = offsets[s];
    const float4* src = xyzi + (size_t)s * stride;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
      st_f32x4_if(fused + off + i, src[i], pol, 1u);
  }
}

}  // namespace

cudaError_t cloud_configure() { return cudaSuccess; }

cudaError_t cloud_workspace_alloc(CloudWorkspace& ws, int num_sms, uint32_t max_nodes) {
  // trig table: (float)cos((double)angle_rad), (float)sin(...) with angle_rad exactly as
  // publish_scan computes it (reference rplidar_node.cpp:586-587)
  std::vector<float2> h(65536), ang(65536);
  const double two_pi = 2.0 * 3.14159265358979323846;
  for (uint32_t k = 0; k < 65536; ++k) {
    const float deg = static_cast<float>(k) * 90.0f / 16384.0f;
    const float rad = static_cast<float>(static_cast<double>(deg) * (3.14159265358979323846 / 180.0));
    h[k].x = static_cast<float>(std::cos(static_cast<double>(rad)));
    h[k].y = static_cast<float>(std::sin(static_cast<double>(rad)));
    // Mode A angles (reference rplidar_node.cpp:641-649): plain and inverted
    float inv = static_cast<float>(two_pi - static_cast<double>(rad));
    if (static_cast<double>(inv) >= two_pi) inv = static_cast<float>(static_cast<double>(inv) - two_pi);
    ang[k].x = rad;
    ang[k].y = inv;
  }
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ws.trig), h.size() * sizeof(float2));
  if (e != cudaSuccess) return e;
  e = cudaMemcpy(ws.trig, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  e = cudaMalloc(reinterpret_cast<void**>(&ws.angle), ang.size() * sizeof(float2));
  if (e != cudaSuccess) return e;
  e = cudaMemcpy(ws.angle, ang.data(), ang.size() * sizeof(float2), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  ws.max_nodes = max_nodes;
  ws.ctas = num_sms * 2;
  const uint32_t hsize = hash_size_for(max_nodes);
  ws.scratch_per_cta = (size_t)hsize * (8 * 4 + 4 * 3) + (size_t)max_nodes * 12 + 64;
  ws.scratch_per_cta = (ws.scratch_per_cta + 255) & ~(size_t)255;
  return cudaMalloc(&ws.scratch, ws.scratch_per_cta * ws.ctas);
}

void cloud_workspace_free(CloudWorkspace& ws) {
  cudaFree(ws.trig);
  cudaFree(ws.angle);
  cudaFree(ws.scratch);
  ws = CloudWorkspace{};
}

cudaError_t launch_cloud_post(float4* xyzi, uint32_t* point_counts, uint32_t n_scans, uint32_t stride,
                              uint32_t sor_k, float sor_alpha, float voxel, const CloudWorkspace& ws,
                              cudaStream_t stream, int* launched) {
  const int grid = (int)min((uint32_t)ws.ctas, n_scans);
  if (sor_k > 0) {
    cloud_sor_kernel<<<grid, CT, 0, stream>>>(xyzi, point_counts, n_scans, stride, sor_k, sor_alpha, ws.scratch,
                                              ws.scratch_per_cta, ws.max_nodes);
    if (launched) ++*launched;
  }
  if (voxel > 0.0f) {
    cloud_voxel_kernel<<<grid, CT, 0, stream>>>(xyzi, point_counts, n_scans, stride, voxel, ws.scratch,
                                                ws.scratch_per_cta, ws.max_nodes);
    if (launched) ++*launched;
  }
  return cudaGetLastError();
}

cudaError_t launch_cloud_fuse(const float4* xyzi, const uint32_t* point_counts, uint32_t n_scans,
                              uint32_t stride, float4* fused, uint32_t* offsets, uint32_t* total,
                              cudaStream_t stream, int* launched) {
  if (n_scans == 0) return cudaSuccess;
  cloud_offsets_kernel<<<1, 1024, 0, stream>>>(point_counts, n_scans, offsets, total);
  const uint32_t gy = min(n_scans, 65535u);
  const uint32_t gx = max(1u, min(32u, (stride + 255u) / 256u));
  cloud_pack_kernel<<<dim3(gx, gy), 256, 0, stream>>>(xyzi, point_counts, offsets, n_scans, stride, fused);
  if (launched) *launched += 2;
  return cudaGetLastError();
}

}  // namespace rpl
