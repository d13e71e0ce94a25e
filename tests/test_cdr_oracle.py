"""SURVEY.md 8(f) rank 3 -- the CDR restatement (oracle/cdr_oracle.py, parity unpinned: see its header):
writer/reader round trips, alignment rules, and the size functions the C-ABI exports.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from oracle import cdr_oracle as cdr


@pytest.mark.parametrize("frame_id", ["", "a", "laser", "laser_frame", "base_link/lidar_0007", "x" * 255])
@pytest.mark.parametrize("n", [0, 1, 7, 360])
def test_laserscan_round_trip_and_size(frame_id, n):
    rng = np.random.default_rng(n)
    r, it = rng.random(n).astype(np.float32), rng.random(n).astype(np.float32)
    sc = rng.random(7).astype(np.float32)
    b = cdr.laserscan_cdr(-5, 999999999, frame_id, sc, r, it)
    m = cdr.parse_laserscan(b)
    assert m["sec"] == -5 and m["nanosec"] == 999999999 and m["frame_id"] == frame_id
    assert (m["ranges"].view(np.uint32) == r.view(np.uint32)).all()
    assert (m["intensities"].view(np.uint32) == it.view(np.uint32)).all()
    assert m["angle_increment"].tobytes() == sc[2].tobytes()
    import rplidar_ros2_driver_b200 as R

    assert R.lib().rpl_laserscan_cdr_size(len(frame_id), n) == len(b)


@pytest.mark.parametrize("frame_id", ["", "laser", "laser_frame", "abc", "x" * 255])
@pytest.mark.parametrize("n", [0, 1, 5, 3200])
def test_pointcloud2_round_trip_and_size(frame_id, n):
    rng = np.random.default_rng(n + 1)
    pts = rng.random((n, 4)).astype(np.float32)
    b = cdr.pointcloud2_cdr(17, 5, frame_id, pts)
    m = cdr.parse_pointcloud2(b)
    assert m["frame_id"] == frame_id and m["height"] == 1 and m["width"] == n
    assert m["fields"] == [("x", 0, 7, 1), ("y", 4, 7, 1), ("z", 8, 7, 1), ("intensity", 12, 7, 1)]
    assert m["point_step"] == 16 and m["row_step"] == 16 * n and m["is_dense"] == 1 and m["is_bigendian"] == 0
    assert m["data"].tobytes() == pts.tobytes()
    import rplidar_ros2_driver_b200 as R

    assert R.lib().rpl_pointcloud2_cdr_size(len(frame_id), n) == len(b)


def test_known_bytes():
    """A message small enough to check by hand against the CDR rules."""
    b = cdr.laserscan_cdr(1, 2, "ab", [0, 0, 0, 0, 0, 0, 0], [1.0], [2.0])
    expect = (b"\x00\x01\x00\x00" + b"\x01\x00\x00\x00" + b"\x02\x00\x00\x00" + b"\x03\x00\x00\x00ab\x00" + b"\x00"
              + b"\x00" * 28 + b"\x01\x00\x00\x00\x00\x00\x80\x3f" + b"\x01\x00\x00\x00\x00\x00\x00\x40")
    assert b == expect
