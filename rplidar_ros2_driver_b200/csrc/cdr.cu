// cdr.cu -- LaserScan / PointCloud2 -> wire (the step AFTER the hot path).
//
// The reference hands a sensor_msgs::msg::LaserScan to rclcpp (src/rplidar_node.cpp:679) and the RMW
// layer serialises it to CDR before it reaches the wire.  Here the serialised message is produced on the
// device, straight from the scan kernels' outputs, so that the host can publish the bytes as they are
// (rclcpp::SerializedMessage) without building the message or copying the arrays
// again.  Encoding: XCDR version 1, little endian, the representation ROS 2's RMW implementations put
// on the wire -- 4-byte encapsulation header {0x00,0x01,0x00,0x00}, then the members in declaration
// order, every primitive aligned to its size relative to the first byte after that header, strings as
// uint32 length (terminating NUL included) + bytes, sequences as uint32 count + elements.
//   sensor_msgs/LaserScan:   header{stamp{int32 sec, uint32 nanosec}, string frame_id}, 7 x float32,
//                            float32[] ranges, float32[] intensities
//   sensor_msgs/PointCloud2: header, uint32 height, width, PointField[] fields{string name, uint32
//                            offset, uint8 datatype, uint32 count}, bool is_bigendian, uint32
//                            point_step, row_step, uint8[] data, bool is_dense
// The fixed part of each message (everything but the arrays) is a host-built template with a handful
// of 4-byte fields patched per message; the arrays are copied coalesced.
// PARITY UNPINNED: /root/reference holds no serialiser (it lives in the RMW dependency); the format
// follows the OMG CDR rules above and oracle/cdr_oracle.py restates it independently in numpy.
#include "cdr_args.h"
#include "rpl_device.cuh"

namespace rpl {

namespace {

constexpr int CT = 256;

__device__ __forceinline__ void put32(uint8_t* p, uint32_t v) { *reinterpret_cast<uint32_t*>(p) = v; }

__global__ void __launch_bounds__(CT) laserscan_cdr_kernel(LaserScanCdrArgs a, CdrTemplate t) {
  const uint32_t s = blockIdx.y;
  const uint32_t n = a.beam_counts[s];
  uint8_t* msg = a.cdr_out + (size_t)s * a.cdr_stride;
  const uint32_t P = t.prefix_bytes;  // multiple of 4; ends with the ranges count
  if (blockIdx.x == 0) {
    for (uint32_t w = threadIdx.x; w < P / 4; w += CT) put32(msg + 4 * w, reinterpret_cast<const uint32_t*>(t.prefix)[w]);
    __syncthreads();
    if (threadIdx.x == 0) {
      const LaserScanMeta m = a.meta[s];
      put32(msg + 4, (uint32_t)m.stamp_sec);
      put32(msg + 8, m.stamp_nanosec);
      uint8_t* f = msg + P - 32;  // 7 floats, then the ranges count
      put32(f + 0, __float_as_uint(m.angle_min));
      put32(f + 4, __float_as_uint(m.angle_max));
      put32(f + 8, __float_as_uint(a.angle_increment ? a.angle_increment[s] : m.angle_increment));
      put32(f + 12, __float_as_uint(m.time_increment));
      put32(f + 16, __float_as_uint(m.scan_time));
      put32(f + 20, __float_as_uint(m.range_min));
      put32(f + 24, __float_as_uint(m.range_max));
      put32(f + 28, n);
      put32(msg + P + 4 * (size_t)n, n);  // intensities count
      if (a.cdr_sizes) a.cdr_sizes[s] = P + 8 * n + 4;
    }
  }
  uint32_t* r_out = reinterpret_cast<uint32_t*>(msg + P);
  uint32_t* i_out = r_out + n + 1;
  const uint32_t* r_in = reinterpret_cast<const uint32_t*>(a.ranges + (size_t)s * a.stride);
  const uint32_t* i_in = reinterpret_cast<const uint32_t*>(a.intensities + (size_t)s * a.stride);
  for (uint32_t i = blockIdx.x * CT + threadIdx.x; i < n; i += gridDim.x * CT) {
    r_out[i] = __ldg(r_in + i);
    i_out[i] = __ldg(i_in + i);
  }
}

__global__ void __launch_bounds__(CT) pointcloud2_cdr_kernel(PointCloudCdrArgs a, CdrTemplate t) {
  const uint32_t s = blockIdx.y;
  const uint32_t n = a.point_counts[s];
  uint8_t* msg = a.cdr_out + (size_t)s * a.cdr_stride;
  const uint32_t P = t.prefix_bytes;  // multiple of 4; ends with the data length
  if (blockIdx.x == 0) {
    for (uint32_t w = threadIdx.x; w < P / 4; w += CT) put32(msg + 4 * w, reinterpret_cast<const uint32_t*>(t.prefix)[w]);
    __syncthreads();
    if (threadIdx.x == 0) {
      put32(msg + 4, (uint32_t)a.stamps[2 * s]);
      put32(msg + 8, a.stamps[2 * s + 1]);
      put32(msg + t.patch_width, n);         // width (height = 1: unorganised cloud)
      put32(msg + t.patch_row_step, 16u * n);
      put32(msg + P - 4, 16u * n);           // data length
      msg[P + 16 * (size_t)n] = 1;           // is_dense: the cloud path drops unmeasured points
      if (a.cdr_sizes) a.cdr_sizes[s] = P + 16 * n + 1;
    }
  }
  // the prefix length follows from the frame id: 16-byte stores only when it happens to be a multiple of 16
  uint4* d_out = reinterpret_cast<uint4*>(msg + P);
  const uint4* d_in = reinterpret_cast<const uint4*>(a.xyzi + (size_t)s * a.stride * 4);
  if ((P & 15u) == 0) {
    for (uint32_t i = blockIdx.x * CT + threadIdx.x; i < n; i += gridDim.x * CT) d_out[i] = __ldg(d_in + i);
  } else {
    uint32_t* o = reinterpret_cast<uint32_t*>(msg + P);
    const uint32_t* in = reinterpret_cast<const uint32_t*>(d_in);
    for (uint32_t i = blockIdx.x * CT + threadIdx.x; i < 4 * n; i += gridDim.x * CT) o[i] = __ldg(in + i);
  }
}

}  // namespace

cudaError_t launch_laserscan_cdr(const LaserScanCdrArgs& a, const CdrTemplate& t, cudaStream_t stream) {
  if (a.n_scans == 0) return cudaSuccess;
  const uint32_t bx = std::max(1u, std::min((a.stride + CT * 4 - 1) / (CT * 4), 32u));
  for (uint32_t s0 = 0; s0 < a.n_scans; s0 += 65535) {
    LaserScanCdrArgs b = a;
    b.meta += s0;
    b.ranges += (size_t)s0 * a.stride;
    b.intensities += (size_t)s0 * a.stride;
    b.beam_counts += s0;
    if (b.angle_increment) b.angle_increment += s0;
    b.cdr_out += (size_t)s0 * a.cdr_stride;
    if (b.cdr_sizes) b.cdr_sizes += s0;
    laserscan_cdr_kernel<<<dim3(bx, std::min(65535u, a.n_scans - s0)), CT, 0, stream>>>(b, t);
  }
  return cudaGetLastError();
}

cudaError_t launch_pointcloud2_cdr(const PointCloudCdrArgs& a, const CdrTemplate& t, cudaStream_t stream) {
  if (a.n_clouds == 0) return cudaSuccess;
  const uint32_t bx = std::max(1u, std::min((a.stride + CT * 4 - 1) / (CT * 4), 32u));
  for (uint32_t s0 = 0; s0 < a.n_clouds; s0 += 65535) {
    PointCloudCdrArgs b = a;
    b.stamps += 2 * (size_t)s0;
    b.xyzi += (size_t)s0 * a.stride * 4;
    b.point_counts += s0;
    b.cdr_out += (size_t)s0 * a.cdr_stride;
    if (b.cdr_sizes) b.cdr_sizes += s0;
    pointcloud2_cdr_kernel<<<dim3(bx, std::min(65535u, a.n_clouds - s0)), CT, 0, stream>>>(b, t);
  }
  return cudaGetLastError();
}

}  // namespace rpl
