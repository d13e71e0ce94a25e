"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref).

Run in the authoring container only (needs /root/reference to build oracle/_ref):

    python tests/golden/make_golden.py

Fixtures:
  dummy_scans.npz
    raw[16,360]          nodes captured from the reference's DummyLidarDriver::grab_scan_data
                         calls 1..16 (reference src/lidar_driver_wrapper.cpp:441-471; captured,
                         not regenerated, because the generator uses libm sin)
    variants[V,360]      raw[0], raw[0] with nodes {0,1,100,359} zeroed, raw[0] rotated by
                         120 positions, the zeroed one rotated, raw[0] with a leading run of
                         7 unmeasured nodes, raw[0] with a trailing run of 5
    variants_ascended    the reference's own ascendScanData output for each variant
    variants_rc          its sl_result
  ascend_cases.npz
    small hand-made edge cases (count 1/2/3, all invalid, wrap past 360, clamp at 0) with
    the reference's ascendScanData outputs
  laserscan_golden.npz
    for every variant x {is_new_protocol} x {Mode A, Mode B} x {inverted}:
    ranges / intensities / header scalars of the reference's REAL RPlidarNode::publish_scan
    (src/rplidar_node.cpp compiled in place against the ROS API stubs in oracle/ros_stubs/ ->
    oracle/_ref/libref_node.so) applied to the reference-ascended buffer and to the raw buffer.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def variants_of(scan: np.ndarray) -> np.ndarray:
    v = [scan.copy()]
    z = scan.copy()
    for i in (0, 1, 100, 359):
        z[i]["dist_mm_q2"] = 0
        z[i]["quality"] = 0
    v.append(z)
    v.append(np.roll(scan, -120))
    v.append(np.roll(z, -120))
    lead = scan.copy()
    lead["dist_mm_q2"][:7] = 0
    v.append(lead)
    trail = scan.copy()
    trail["dist_mm_q2"][-5:] = 0
    v.append(trail)
    return np.stack(v)


def edge_cases():
    mk = O.make_nodes
    cases = []
    cases.append(mk([100], [4000], [40]))
    cases.append(mk([100], [0]))
    cases.append(mk([5, 3], [4000, 8000], [1, 2]))
    cases.append(mk([0, 0, 0], [0, 0, 0]))
    cases.append(mk([10, 20, 30, 65000], [0, 0, 0, 4000], [0, 0, 0, 9]))  # head chain clamps at 0
    cases.append(mk([65000, 1, 2, 3], [4000, 0, 0, 0], [9, 0, 0, 0]))  # fill wraps past 360
    cases.append(mk([30000, 0, 0, 40000, 0, 50000, 0, 0], [400, 0, 0, 500, 0, 600, 0, 0],
                    [1, 0, 0, 2, 0, 3, 0, 0]))
    cases.append(mk([0, 0, 40000, 0, 10, 0], [0, 0, 777, 0, 888, 0], [0, 0, 5, 0, 6, 0]))
    # 17..40 nodes, some unmeasured, unique keys (no tie ambiguity)
    rng = np.random.default_rng(7)
    for n in (17, 33, 40):
        keys = rng.choice(65536, size=n, replace=False)
        dist = rng.integers(1, 200000, size=n)
        dist[rng.random(n) < 0.3] = 0
        if not dist.any():
            dist[n // 2] = 5
        cases.append(mk(keys, dist, rng.integers(0, 256, size=n)))
    return cases


def main():
    O.build(ref=True)
    assert O.have_ref() and O.have_ref_node(), "oracle/_ref missing: run in the authoring container"
    assert O.ref().ref_sizeof_node() == 8

    raw = np.stack([O.ref_dummy_grab() for _ in range(16)])
    var = variants_of(raw[0])
    asc = np.zeros_like(var)
    rcs = np.zeros(len(var), dtype=np.uint32)
    for i, v in enumerate(var):
        rcs[i], asc[i] = O.ref_ascend(v)
    np.savez_compressed(os.path.join(OUT, "dummy_scans.npz"), raw=raw.view(np.uint8).reshape(16, 360, 8),
                        variants=var.view(np.uint8).reshape(len(var), 360, 8),
                        variants_ascended=asc.view(np.uint8).reshape(len(var), 360, 8), variants_rc=rcs)

    cases = edge_cases()
    d = {}
    for i, c in enumerate(cases):
        rc, out = O.ref_ascend(c)
        d[f"in_{i}"] = c.view(np.uint8).reshape(-1, 8)
        d[f"out_{i}"] = out.view(np.uint8).reshape(-1, 8)
        d[f"rc_{i}"] = np.uint32(rc)
    d["n_cases"] = np.int32(len(cases))
    np.savez_compressed(os.path.join(OUT, "ascend_cases.npz"), **d)

    g = {}
    idx = 0
    for vi in range(len(var)):
        for use_asc in (0, 1):
            nodes = asc[vi] if use_asc else var[vi]
            for newp in (0, 1):
                for mode_a in (0, 1):
                    for inv in (0, 1):
                        prm = O.scan_params(newp, mode_a, inv, use_asc, 12.0, 0.1)
                        published, h7, r, it = O.ref_publish(nodes, prm)  # the node's own publish_scan
                        assert published
                        g[f"cfg_{idx}"] = np.array([vi, use_asc, newp, mode_a, inv], dtype=np.int32)
                        g[f"ranges_{idx}"] = r
                        g[f"intens_{idx}"] = it
                        g[f"hdr_{idx}"] = h7.astype(np.float32)
                        g[f"beams_{idx}"] = np.uint32(len(r))
                        idx += 1
    g["n"] = np.int32(idx)
    np.savez_compressed(os.path.join(OUT, "laserscan_golden.npz"), **g)
    print("wrote golden fixtures:", os.listdir(OUT))


if __name__ == "__main__":
    main()
