"""Independent pins for the PointCloud2 extensions (SURVEY.md 8(c) lists them as "parity unpinned": the
reference has no such code, oracle/cloud_oracle.cpp is a self-authored definition).  These tests check that
definition -- and, in tests/test_gpu_cloud.py, the CUDA path directly -- against checkers that share no code
and no arithmetic with it:

  * polar -> Cartesian against float64 numpy with the TRUE angle of the key (key * 2 pi / 65536) and the exact
    range (dist_mm_q2 / 4000): laser_geometry::projectLaser semantics (x = r cos a, y = r sin a, z = 0, points
    outside [range_min, range_max] dropped).  Tolerance: BASELINE.json's north_star, 1e-6 relative -- taken
    relative to the point's range, |p32 - p64| <= 1e-6 * r (a per-component relative bound is meaningless where
    cos or sin cross zero).  Observed: <= 4.2e-7 (float32 range, float32 angle, one rounding per product).
  * statistical outlier removal against brute-force k-nearest-neighbour search over ALL points of the scan
    (PCL StatisticalOutlierRemoval semantics: mean distance to the k nearest neighbours, threshold mean +
    alpha * stddev).  The definition here searches only the 16 points before and the 16 after in angle order.
    Where the k nearest neighbours of a point all lie inside that window the two mean distances are the SAME
    number (asserted, exact); that holds for ~99 % of the points of a room scan.  The rest are points whose
    nearest neighbours in space are far away in angle (a wall seen again behind a corner, both sides of a thin
    gap): the window form then over-estimates their mean distance (asserted: never under-estimates), which can
    only move a point towards "outlier".  On the room scans the keep/drop decisions are identical.
"""
import numpy as np
import pytest

from test_gpu_cloud import room_scans


def float64_projection(nodes, range_min, range_max, intensity_min=0.0, new_protocol=0):
    """laser_geometry-style projection in float64, kept points in angle order."""
    d = nodes["dist_mm_q2"].astype(np.float64) / 4000.0
    inten = (nodes["quality"] if new_protocol else (nodes["quality"] >> 2)).astype(np.float64)
    # the window is applied to the float32 range the driver publishes (publish_scan's dist_m)
    r32 = (nodes["dist_mm_q2"].astype(np.float32) / np.float32(4000.0))
    keep = (nodes["dist_mm_q2"] != 0) & (r32 >= np.float32(range_min)) & (r32 <= np.float32(range_max)) & (inten >= intensity_min)
    k = nodes["angle_z_q14"][keep].astype(np.int64)
    order = np.argsort(k, kind="stable")
    th = k[order] * (2.0 * np.pi / 65536.0)
    r = d[keep][order]
    return np.stack([r * np.cos(th), r * np.sin(th), np.zeros_like(r), inten[keep][order]], axis=1), r


def check_projection(points32, nodes, **kw):
    exp, r = float64_projection(nodes, **kw)
    assert points32.shape[0] == exp.shape[0]
    err = np.hypot(points32[:, 0].astype(np.float64) - exp[:, 0], points32[:, 1].astype(np.float64) - exp[:, 1])
    assert (err <= 1e-6 * r).all(), float((err / r).max())
    assert (points32[:, 2] == 0).all() and (points32[:, 3].astype(np.float64) == exp[:, 3]).all()
    return float((err / r).max())


@pytest.mark.parametrize("variant,n", [(0, 3200), (1, 3200), (4, 3200), (3, 8192), (1, 32768), (0, 360)])
def test_polar_to_xyz_within_1e6_of_float64(oracle, variant, n):
    nodes = oracle.synth_batch(9000 + variant, 3, n, variant)
    worst = 0.0
    for s in range(3):
        pts = oracle.cloud(nodes[s], oracle.cloud_params(range_min=0.15, range_max=40.0))
        worst = max(worst, check_projection(pts, nodes[s], range_min=0.15, range_max=40.0))
    assert worst < 1e-6


def test_projection_at_the_extremes_of_the_key_and_range_space(oracle):
    """every key once, ranges from 1/4 mm to the 40 m window edge and beyond"""
    keys = np.arange(65536)
    rng = np.random.default_rng(5)
    dist = np.concatenate([[1, 2, 3, 599, 600, 160000, 160001, 2**31], rng.integers(1, 200000, 65536 - 8)])
    nodes = oracle.make_nodes(keys, dist, rng.integers(0, 256, 65536), 2)
    pts = oracle.cloud(nodes, oracle.cloud_params(range_min=0.0, range_max=1e9))
    check_projection(pts, nodes, range_min=0.0, range_max=1e9)
    pts = oracle.cloud(nodes, oracle.cloud_params(range_min=0.15, range_max=40.0, intensity_min=17.0))
    check_projection(pts, nodes, range_min=0.15, range_max=40.0, intensity_min=17.0)


def knn_and_window(xy, k):
    m = len(xy)
    d = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
    np.fill_diagonal(d, np.inf)
    idx = np.argsort(d, axis=1)[:, :k]
    mean_knn = np.take_along_axis(d, idx, 1).mean(1)
    offs = np.concatenate([np.arange(-16, 0), np.arange(1, 17)])
    widx = (np.arange(m)[:, None] + offs[None, :]) % m
    mean_win = np.sort(np.take_along_axis(d, widx, 1), axis=1)[:, :k].mean(1)
    circ = np.minimum((idx - np.arange(m)[:, None]) % m, (np.arange(m)[:, None] - idx) % m)
    return mean_knn, mean_win, (circ <= 16).all(1)


@pytest.mark.parametrize("seed,n", [(1, 3200), (2, 3200), (3, 800), (4, 360)])
def test_sor_window_against_brute_force_knn(oracle, seed, n):
    nodes = room_scans(oracle, 1, n, seed)[0]
    base = oracle.cloud(nodes, oracle.cloud_params(range_min=0.15, range_max=40.0))
    xy = base[:, :2].astype(np.float64)
    k, alpha = 8, 1.0
    mean_knn, mean_win, inside = knn_and_window(xy, k)
    # where the k nearest neighbours lie inside the angular window the two definitions are the same number
    assert inside.mean() > 0.98
    assert (mean_knn[inside] == mean_win[inside]).all()
    # elsewhere the window can only over-estimate
    assert (mean_win >= mean_knn).all()

    def keep(q):
        return q <= q.mean() + alpha * q.std(ddof=1)

    knn_keep, win_keep = keep(mean_knn), keep(mean_win)
    assert (knn_keep == win_keep).mean() >= 0.995
    # and the library's definition (float32 distances, fixed-point statistics) takes the same decisions as the
    # float64 window form, up to points that sit on the threshold
    got = oracle.cloud(nodes, oracle.cloud_params(range_min=0.15, range_max=40.0, sor_k=k, sor_alpha=alpha))
    kept = np.zeros(len(base), bool)
    j = 0
    for i in range(len(base)):  # `got` is a subsequence of `base`
        if j < len(got) and (got[j].view(np.uint32) == base[i].view(np.uint32)).all():
            kept[i] = True
            j += 1
    assert j == len(got)
    assert (kept != win_keep).sum() <= 2
    assert (kept != knn_keep).sum() <= max(2, int(0.005 * len(base)))
