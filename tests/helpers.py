"""Shared test helpers: an independent numpy restatement of publish_scan (second opinion
on oracle/scan_oracle.cpp) and scan builders."""
from __future__ import annotations

import numpy as np

F32 = np.float32
TWO_PI_D = 2.0 * np.pi  # the reference's `2.0f * M_PI` is a double product


def np_publish(nodes, is_new_protocol, scan_processing, inverted, stable_keys=True):
    """numpy float32/float64 restatement of reference src/rplidar_node.cpp:581-677.

    Requires tie-free keys among measured nodes (or accepts stable order on ties).
    Returns (ranges, intensities, angle_increment).
    """
    valid = nodes["dist_mm_q2"] != 0
    v = nodes[valid]
    if len(v) == 0:
        return np.zeros(0, F32), np.zeros(0, F32), F32(0)
    deg = (v["angle_z_q14"].astype(F32) * F32(90.0)) / F32(16384.0)
    rad = (deg.astype(np.float64) * (np.pi / 180.0)).astype(F32)
    dist = v["dist_mm_q2"].astype(F32) / F32(4000.0)
    inten = (v["quality"] if is_new_protocol else (v["quality"] >> 2)).astype(F32)
    order = np.argsort(rad, kind="stable")
    rad, dist, inten = rad[order], dist[order], inten[order]
    m = len(rad)
    if scan_processing:
        inc = F32(TWO_PI_D / float(m))
        a = rad
        if inverted:
            a = (TWO_PI_D - a.astype(np.float64)).astype(F32)
            wrap = a.astype(np.float64) >= TWO_PI_D
            a = np.where(wrap, (a.astype(np.float64) - TWO_PI_D).astype(F32), a)
        idx = ((a - F32(0.0)) / inc).astype(np.int32)
        ranges = np.full(m, np.inf, F32)
        intens = np.zeros(m, F32)
        for i in range(m):
            b = idx[i]
            if 0 <= b < m and dist[i] < ranges[b]:
                ranges[b] = dist[i]
                intens[b] = inten[i]
        return ranges, intens, inc
    denom = float(m - 1 if m > 1 else 1)
    inc = F32(TWO_PI_D / denom)
    if inverted:
        return dist[::-1].copy(), inten[::-1].copy(), inc
    return dist, inten, inc


def bits(a):
    """float32 array -> uint32 view for bit-exact comparison."""
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
