"""oracle/cdr_oracle.py -- TEST INFRASTRUCTURE ONLY.

SURVEY.md 8(f) rank 3: the serialised form of the messages the reference publishes
(scan_pub_->publish, reference src/rplidar_node.cpp:679).  PARITY UNPINNED: the serialiser is in the
RMW dependency (rmw_fastrtps / rmw_cyclonedds, not vendored under /root/reference), so there is no
golden vector to hold this against.  What is restated here is the published encoding those
implementations share for ROS 2 messages: OMG CDR (XCDR version 1), little endian --
  * 4-byte encapsulation header 00 01 00 00,
  * members in declaration order, each primitive aligned to its own size counted from the first byte
    after the encapsulation header, padding bytes zero,
  * string = uint32 length including the terminating NUL + bytes + NUL,
  * sequence<T> = uint32 element count + elements,
and the message definitions sensor_msgs/msg/LaserScan, sensor_msgs/msg/PointCloud2, PointField,
std_msgs/msg/Header, builtin_interfaces/msg/Time.  A writer and an independent reader: the tests
check writer -> reader round trips and the device output byte for byte against the writer."""
import struct

import numpy as np


class _W:
    def __init__(self):
        self.b = bytearray(b"\x00\x01\x00\x00")

    def align(self, a):
        while (len(self.b) - 4) % a:
            self.b.append(0)

    def prim(self, fmt, v):
        self.align(struct.calcsize(fmt))
        self.b += struct.pack("<" + fmt, v)

    def string(self, s: str):
        raw = s.encode()
        self.prim("I", len(raw) + 1)
        self.b += raw + b"\x00"

    def header(self, sec, nanosec, frame_id):
        self.prim("i", sec)
        self.prim("I", nanosec)
        self.string(frame_id)


def laserscan_cdr(sec, nanosec, frame_id, scalars7, ranges, intensities) -> bytes:
    w = _W()
    w.header(sec, nanosec, frame_id)
    for v in scalars7:  # angle_min, angle_max, angle_increment, time_increment, scan_time, range_min, range_max
        w.align(4)
        w.b += np.float32(v).tobytes()
    for arr in (ranges, intensities):
        a = np.ascontiguousarray(arr, dtype="<f4")
        w.prim("I", a.shape[0])
        w.b += a.tobytes()
    return bytes(w.b)


def pointcloud2_cdr(sec, nanosec, frame_id, xyzi) -> bytes:
    pts = np.ascontiguousarray(xyzi, dtype="<f4").reshape(-1, 4)
    n = pts.shape[0]
    w = _W()
    w.header(sec, nanosec, frame_id)
    w.prim("I", 1)  # height
    w.prim("I", n)  # width
    w.prim("I", 4)  # fields
    for k, name in enumerate(("x", "y", "z", "intensity")):
        w.string(name)
        w.prim("I", 4 * k)
        w.prim("B", 7)  # FLOAT32
        w.prim("I", 1)
    w.prim("B", 0)      # is_bigendian
    w.prim("I", 16)     # point_step
    w.prim("I", 16 * n)  # row_step
    w.prim("I", 16 * n)  # data
    w.b += pts.tobytes()
    w.prim("B", 1)      # is_dense
    return bytes(w.b)


class _R:
    def __init__(self, b: bytes):
        assert b[:4] == b"\x00\x01\x00\x00", "not XCDR1 little endian"
        self.b, self.p = b, 4

    def prim(self, fmt):
        size = struct.calcsize(fmt)
        while (self.p - 4) % size:
            assert self.b[self.p] == 0, "non-zero padding"
            self.p += 1
        (v,) = struct.unpack_from("<" + fmt, self.b, self.p)
        self.p += size
        return v

    def string(self):
        n = self.prim("I")
        raw = self.b[self.p: self.p + n]
        self.p += n
        assert raw[-1] == 0
        return raw[:-1].decode()

    def f32seq(self):
        n = self.prim("I")
        a = np.frombuffer(self.b, "<f4", n, self.p)
        self.p += 4 * n
        return a


def parse_laserscan(b: bytes) -> dict:
    r = _R(b)
    out = dict(sec=r.prim("i"), nanosec=r.prim("I"), frame_id=r.string())
    for k in ("angle_min", "angle_max", "angle_increment", "time_increment", "scan_time", "range_min", "range_max"):
        out[k] = np.float32(r.prim("f"))
    out["ranges"] = r.f32seq()
    out["intensities"] = r.f32seq()
    assert r.p == len(b), "trailing bytes"
    return out


def parse_pointcloud2(b: bytes) -> dict:
    r = _R(b)
    out = dict(sec=r.prim("i"), nanosec=r.prim("I"), frame_id=r.string(), height=r.prim("I"), width=r.prim("I"))
    fields = []
    for _ in range(r.prim("I")):
        fields.append((r.string(), r.prim("I"), r.prim("B"), r.prim("I")))
    out["fields"] = fields
    out["is_bigendian"] = r.prim("B")
    out["point_step"] = r.prim("I")
    out["row_step"] = r.prim("I")
    n = r.prim("I")
    out["data"] = np.frombuffer(b, np.uint8, n, r.p)
    r.p += n
    out["is_dense"] = r.prim("B")
    assert r.p == len(b), "trailing bytes"
    return out
