#!/usr/bin/env python
"""Hand-derived CDR bytes of one sensor_msgs/LaserScan and one sensor_msgs/PointCloud2 -> cdr_laserscan.bin,
cdr_pointcloud2.bin.

Why: the serialiser the reference relies on lives in its RMW dependency (rmw_fastrtps / rmw_cyclonedds), not under
/root/reference, and no ROS 2 is installed here, so there is no vector from a real RMW.  Until one is available
this file is the third, spec-derived source next to the device writer (csrc/cdr.cu) and the numpy writer
(oracle/cdr_oracle.py): every byte below is laid out BY HAND from the rules, offset by offset -- there is no
alignment routine in this script, each padding run is a literal whose length is justified in the comment.

Rules used (OMG "Extended CDR", XCDR version 1 = classic PLAIN_CDR, little endian; ROS 2 rosidl type mapping):
  R1  4-byte encapsulation header 00 01 00 00 (CDR_LE, options 0); offsets below count from the byte AFTER it
  R2  a primitive of size s starts at an offset that is a multiple of s; padding bytes are 0
  R3  string  = uint32 (length including the terminating NUL) + characters + NUL
  R4  sequence<T> = uint32 element count + elements (each aligned by R2); uint8[] is sequence<uint8>
  R5  struct = its members in declaration order, no padding of its own
Message definitions: std_msgs/Header {builtin_interfaces/Time stamp {int32 sec; uint32 nanosec}; string frame_id},
sensor_msgs/LaserScan {Header header; float32 angle_min, angle_max, angle_increment, time_increment, scan_time,
range_min, range_max; float32[] ranges; float32[] intensities},
sensor_msgs/PointField {string name; uint32 offset; uint8 datatype; uint32 count},
sensor_msgs/PointCloud2 {Header header; uint32 height, width; PointField[] fields; bool is_bigendian;
uint32 point_step, row_step; uint8[] data; bool is_dense}.
"""
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))

u32 = lambda v: struct.pack("<I", v)
i32 = lambda v: struct.pack("<i", v)
f32 = lambda v: struct.pack("<f", v)
PAD = lambda n: b"\x00" * n

# ---------------------------------------------------------------------------------------------- LaserScan
LASERSCAN = dict(sec=1_700_000_000, nanosec=123_456_789, frame_id="laser",
                 scalars=[0.0, 6.2831854820251465, 0.0019640722312033176, 3.125e-05, 0.10000000149011612, 0.15000000596046448, 40.0],
                 ranges=[1.0, float("inf"), 2.5], intensities=[47.0, 0.0, 12.0])
ls = b"".join([
    b"\x00\x01\x00\x00",                 # R1
    i32(1_700_000_000),                  # off  0  header.stamp.sec
    u32(123_456_789),                    # off  4  header.stamp.nanosec
    u32(6), b"laser\x00",                # off  8  frame_id: length 5 + NUL = 6, then 6 bytes -> ends at 18
    PAD(2),                              # off 18  next member is a float32: 18 -> 20 (R2)
    f32(0.0),                            # off 20  angle_min
    f32(6.2831854820251465),             # off 24  angle_max      = (float)(2 pi)
    f32(0.0019640722312033176),          # off 28  angle_increment
    f32(3.125e-05),                      # off 32  time_increment
    f32(0.10000000149011612),            # off 36  scan_time
    f32(0.15000000596046448),            # off 40  range_min
    f32(40.0),                           # off 44  range_max
    u32(3), f32(1.0), f32(float("inf")), f32(2.5),   # off 48  ranges: count, then 3 floats -> ends at 64
    u32(3), f32(47.0), f32(0.0), f32(12.0),          # off 64  intensities -> ends at 80
])
assert len(ls) == 4 + 80

# -------------------------------------------------------------------------------------------- PointCloud2
POINTS = [(1.0, -2.0, 0.0, 47.0), (0.5, 0.25, 0.0, 3.0)]
PC2 = dict(sec=12, nanosec=999_999_999, frame_id="lidar_3", points=POINTS)
pc = b"".join([
    b"\x00\x01\x00\x00",                 # R1
    i32(12),                             # off   0  stamp.sec
    u32(999_999_999),                    # off   4  stamp.nanosec
    u32(8), b"lidar_3\x00",              # off   8  frame_id: 7 + NUL = 8 -> ends at 20 (already 4-aligned)
    u32(1),                              # off  20  height
    u32(2),                              # off  24  width
    u32(4),                              # off  28  fields: 4 elements
    u32(2), b"x\x00",                    # off  32  fields[0].name -> ends at 38
    PAD(2),                              # off  38  -> 40 for the uint32
    u32(0),                              # off  40  fields[0].offset
    b"\x07",                             # off  44  fields[0].datatype FLOAT32 = 7
    PAD(3),                              # off  45  -> 48
    u32(1),                              # off  48  fields[0].count
    u32(2), b"y\x00",                    # off  52  fields[1].name -> 58
    PAD(2),                              # off  58  -> 60
    u32(4),                              # off  60  fields[1].offset
    b"\x07", PAD(3),                     # off  64  datatype, -> 68
    u32(1),                              # off  68  count
    u32(2), b"z\x00",                    # off  72  fields[2].name -> 78
    PAD(2),                              # off  78  -> 80
    u32(8),                              # off  80  offset
    b"\x07", PAD(3),                     # off  84  -> 88
    u32(1),                              # off  88
    u32(10), b"intensity\x00",           # off  92  fields[3].name: 9 + NUL = 10 -> ends at 106
    PAD(2),                              # off 106  -> 108
    u32(12),                             # off 108  offset
    b"\x07", PAD(3),                     # off 112  -> 116
    u32(1),                              # off 116  count
    b"\x00",                             # off 120  is_bigendian = false
    PAD(3),                              # off 121  -> 124
    u32(16),                             # off 124  point_step
    u32(32),                             # off 128  row_step = point_step * width
    u32(32),                             # off 132  data: 32 bytes follow (uint8 elements need no alignment)
    *[f32(v) for p in POINTS for v in p],  # off 136  -> 168
    b"\x01",                             # off 168  is_dense = true
])
assert len(pc) == 4 + 169

if __name__ == "__main__":
    with open(os.path.join(HERE, "cdr_laserscan.bin"), "wb") as f:
        f.write(ls)
    with open(os.path.join(HERE, "cdr_pointcloud2.bin"), "wb") as f:
        f.write(pc)
    print(len(ls), len(pc))
