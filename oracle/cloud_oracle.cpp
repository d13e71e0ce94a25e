// oracle/cloud_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
//
// PARITY UNPINNED.  The reference has no polar->Cartesian, PointCloud2, range/quality
// window, voxel-grid or outlier-removal code at all (a grep for
// PointCloud|voxel|outlier|cos( over /root/reference hits nothing on this path).  These
// are the north-star's extensions; this file is the self-authored DEFINITION the CUDA
// path is tested against.  Semantic anchors (not in the reference tree, not installed):
// laser_geometry::projectLaser (x=r cos a, y=r sin a, z=0, drop r outside
// [range_min,range_max]) and PCL VoxelGrid / StatisticalOutlierRemoval, adapted so that
// every step is order-independent and therefore bit-reproducible on a parallel machine:
//
//  1. window   keep nodes with dist_mm_q2 != 0, range_min <= r <= range_max,
//              intensity >= intensity_min.  r and intensity are unpacked exactly as
//              publish_scan does (reference src/rplidar_node.cpp:586-590).
//  2. order    stable sort by angle_z_q14 (angle_rad is strictly monotonic in it).
//  3. xyz      a = angle_rad (float, as publish_scan); c = (float)cos((double)a),
//              s = (float)sin((double)a); x = r*c, y = r*s (one float rounding each), z = 0.
//  4. SOR      (sor_k > 0) neighbourhood of point i = the 16 points before and the 16
//              after it in the angle order, circular (all other points when fewer than
//              33 remain).  d_ij = sqrtf(dx*dx + dy*dy), products and sum rounded
//              separately.  m_i = (sum of the k smallest d_ij, added in ascending order)
//              / k.  Statistics over the scan are taken on q_i = llrintf(m_i * 65536)
//              with exact integer sums S1 = sum q_i, S2 = sum q_i^2, then in double:
//              mean = S1/n, var = (S2 - S1*S1/n)/(n-1), keep i iff q_i <= mean +
//              alpha*sqrt(var).  Scans with fewer than 2 points keep everything.
//  5. voxel    (voxel_size > 0) cell = (floorf(x/voxel), floorf(y/voxel)).  One output
//              point per occupied cell, cells emitted in the order of their first member
//              in the (post-SOR) angle order.  Centroid and mean intensity come from exact
//              integer sums of llrintf(v * 65536): out = (float)((double)sum /
//              (65536.0 * count)); intensity (an integer 0..255) likewise without scale.
//  Output: 4 floats per point (x, y, z, intensity) = PointCloud2 point_step 16.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <utility>
#include <vector>

#include "oracle.h"

namespace {
struct Pt {
  uint32_t key;
  float r, inten, x, y;
};
constexpr int kSorHalfWindow = 16;
}  // namespace

extern "C" uint32_t orc_cloud_scan(const orc_node_hq* nodes, size_t count,
                                   const orc_cloud_params* p, float* xyzi) {
  std::vector<Pt> pts;
  pts.reserve(count);
  for (size_t i = 0; i < count; ++i) {
    const orc_node_hq& n = nodes[i];
    if (n.dist_mm_q2 == 0) continue;
    float r = n.dist_mm_q2 / 4000.0f;
    float inten = p->is_new_protocol ? static_cast<float>(n.quality)
                                     : static_cast<float>(n.quality >> 2);
    if (r < p->range_min || r > p->range_max || inten < p->intensity_min) continue;
    pts.push_back({n.angle_z_q14, r, inten, 0.f, 0.f});
  }
  std::stable_sort(pts.begin(), pts.end(), [](const Pt& a, const Pt& b) { return a.key < b.key; });
  for (Pt& q : pts) {
    float deg = q.key * 90.0f / 16384.0f;
    float rad = deg * (M_PI / 180.0f);
    float c = static_cast<float>(std::cos(static_cast<double>(rad)));
    float s = static_cast<float>(std::sin(static_cast<double>(rad)));
    q.x = q.r * c;
    q.y = q.r * s;
  }

  if (p->sor_k > 0 && pts.size() >= 2) {
    const long n = static_cast<long>(pts.size());
    std::vector<int64_t> q(n);
    int64_t s1 = 0;
    // S2 can reach 2^62 for 65536 points of 128 m: exact in unsigned 64-bit
    uint64_t s2 = 0;
    for (long i = 0; i < n; ++i) {
      std::vector<float> d;
      auto add = [&](long j) {
        float dx = pts[j].x - pts[i].x, dy = pts[j].y - pts[i].y;
        float a = dx * dx, b = dy * dy;
        d.push_back(std::sqrt(a + b));
      };
      if (n - 1 <= 2 * kSorHalfWindow) {
        for (long j = 0; j < n; ++j)
          if (j != i) add(j);
      } else {
        for (int o = 1; o <= kSorHalfWindow; ++o) {
          add(((i - o) % n + n) % n);
          add((i + o) % n);
        }
      }
      std::sort(d.begin(), d.end());
      const size_t k = std::min<size_t>(p->sor_k, d.size());
      float sum = 0.f;
      for (size_t t = 0; t < k; ++t) sum += d[t];
      float m = sum / static_cast<float>(k);
      q[i] = llrintf(m * 65536.0f);
      s1 += q[i];
      s2 += static_cast<uint64_t>(q[i]) * static_cast<uint64_t>(q[i]);
    }
    const double dn = static_cast<double>(n);
    const double mean = static_cast<double>(s1) / dn;
    const double var =
        (static_cast<double>(s2) - static_cast<double>(s1) * static_cast<double>(s1) / dn) / (dn - 1.0);
    const double thr = mean + static_cast<double>(p->sor_alpha) * std::sqrt(var > 0.0 ? var : 0.0);
    std::vector<Pt> kept;
    kept.reserve(pts.size());
    for (long i = 0; i < n; ++i)
      if (static_cast<double>(q[i]) <= thr) kept.push_back(pts[i]);
    pts.swap(kept);
  }

  if (p->voxel_size > 0.0f) {
    struct Acc {
      int64_t sx = 0, sy = 0, si = 0;
      uint32_t n = 0, order = 0;
    };
    std::map<std::pair<int32_t, int32_t>, Acc> cells;
    uint32_t next = 0;
    for (const Pt& q : pts) {
      int32_t ix = static_cast<int32_t>(std::floor(q.x / p->voxel_size));
      int32_t iy = static_cast<int32_t>(std::floor(q.y / p->voxel_size));
      auto it = cells.find({ix, iy});
      if (it == cells.end()) {
        it = cells.emplace(std::make_pair(ix, iy), Acc{}).first;
        it->second.order = next++;
      }
      Acc& a = it->second;
      a.sx += llrintf(q.x * 65536.0f);
      a.sy += llrintf(q.y * 65536.0f);
      a.si += static_cast<int64_t>(q.inten);
      a.n += 1;
    }
    for (const auto& kv : cells) {
      const Acc& a = kv.second;
      float* o = xyzi + 4 * static_cast<size_t>(a.order);
      o[0] = static_cast<float>(static_cast<double>(a.sx) / (65536.0 * a.n));
      o[1] = static_cast<float>(static_cast<double>(a.sy) / (65536.0 * a.n));
      o[2] = 0.0f;
      o[3] = static_cast<float>(static_cast<double>(a.si) / static_cast<double>(a.n));
    }
    return next;
  }

  for (size_t i = 0; i < pts.size(); ++i) {
    xyzi[4 * i + 0] = pts[i].x;
    xyzi[4 * i + 1] = pts[i].y;
    xyzi[4 * i + 2] = 0.0f;
    xyzi[4 * i + 3] = pts[i].inten;
  }
  return static_cast<uint32_t>(pts.size());
}
