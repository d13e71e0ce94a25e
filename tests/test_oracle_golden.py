"""Pins the oracle: restatement vs fixtures captured from the compiled reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from helpers import bits, np_publish


def _nodes(a, O):
    return np.ascontiguousarray(a).view(O.NODE_DTYPE).reshape(a.shape[:-1])


def test_node_layout(oracle):
    O = oracle
    assert O.NODE_DTYPE.itemsize == 8
    n = O.make_nodes([0x1234], [0xAABBCCDD], [0x56], [0x78])
    # reference sl_lidar_cmd.h:272-278: packed LE, dist at unaligned offset 2
    assert n.view(np.uint8).tolist() == [0x34, 0x12, 0xDD, 0xCC, 0xBB, 0xAA, 0x56, 0x78]


def test_dummy_generator_matches_captured_reference(oracle, golden_dir):
    g = np.load(f"{golden_dir}/dummy_scans.npz")
    raw = _nodes(g["raw"], oracle)
    for c in range(raw.shape[0]):
        got = oracle.dummy_scan(c + 1)
        assert (got.view(np.uint8) == raw[c].view(np.uint8)).all(), f"dummy call {c + 1}"
    assert raw[0][0].tolist() == (0, 8199, 200, 0)


def test_ascend_matches_reference_on_dummy_variants(oracle, golden_dir):
    g = np.load(f"{golden_dir}/dummy_scans.npz")
    var, asc = _nodes(g["variants"], oracle), _nodes(g["variants_ascended"], oracle)
    for i in range(var.shape[0]):
        for stable in (False, True):  # tie-free after fill here -> both orders agree
            rc, out = oracle.ascend(var[i], stable=stable)
            assert rc == g["variants_rc"][i]
            assert (out.view(np.uint8) == asc[i].view(np.uint8)).all(), (i, stable)


def test_ascend_edge_cases_match_reference(oracle, golden_dir):
    g = np.load(f"{golden_dir}/ascend_cases.npz")
    for i in range(int(g["n_cases"])):
        inp = _nodes(g[f"in_{i}"], oracle)
        rc, out = oracle.ascend(inp)
        assert rc == int(g[f"rc_{i}"]), i
        assert (out.view(np.uint8).reshape(-1, 8) == g[f"out_{i}"]).all(), i
    # all-invalid: OPERATION_FAIL and buffer untouched (reference sl_lidar_driver.cpp:150)
    rc, out = oracle.ascend(oracle.make_nodes([3, 2, 1], [0, 0, 0]))
    assert rc == oracle.RESULT_OPERATION_FAIL and out["angle_z_q14"].tolist() == [3, 2, 1]
    rc, _ = oracle.ascend(oracle.make_nodes([], []))
    assert rc == oracle.RESULT_OPERATION_FAIL


def test_laserscan_golden(oracle, golden_dir):
    g = np.load(f"{golden_dir}/laserscan_golden.npz")
    d = np.load(f"{golden_dir}/dummy_scans.npz")
    var, asc = _nodes(d["variants"], oracle), _nodes(d["variants_ascended"], oracle)
    for k in range(int(g["n"])):
        vi, use_asc, newp, mode_a, inv = g[f"cfg_{k}"].tolist()
        nodes = asc[vi] if use_asc else var[vi]
        hdr, r, it = oracle.publish(nodes, oracle.scan_params(newp, mode_a, inv, use_asc, 12.0, 0.1))
        assert hdr.beam_count == int(g[f"beams_{k}"])
        assert (bits(r) == bits(g[f"ranges_{k}"])).all(), k
        assert (bits(it) == bits(g[f"intens_{k}"])).all(), k
        h = np.array([hdr.angle_min, hdr.angle_max, hdr.angle_increment, hdr.time_increment,
                      hdr.scan_time, hdr.range_min, hdr.range_max], dtype=np.float32)
        assert (bits(h) == bits(g[f"hdr_{k}"])).all(), k
        # second opinion: independent numpy restatement of publish_scan
        r2, it2, inc2 = np_publish(nodes, newp, mode_a, inv)
        assert (bits(r2) == bits(r)).all() and (bits(it2) == bits(it)).all(), k
        assert bits(np.float32(inc2)) == bits(np.float32(hdr.angle_increment)), k


def test_dummy_a1_known_answers(oracle, golden_dir):
    """the dummy A1 grid in Mode A leaves 8 empty bins (+inf) and
    8 double-hit bins; intensity = quality>>2 = 50 in dummy mode."""
    d = np.load(f"{golden_dir}/dummy_scans.npz")
    raw = _nodes(d["raw"], oracle)
    hdr, r, it = oracle.publish(raw[0], oracle.scan_params(0, 1, 0, 0, 40.0, 0.1))
    assert hdr.beam_count == 360 and hdr.published == 1
    assert int(np.isinf(r).sum()) == 8
    assert set(np.unique(it).tolist()) == {0.0, 50.0}
    assert bits(np.float32(hdr.angle_max)) == bits(np.float32(2.0 * np.pi))
    assert hdr.range_min == np.float32(0.15)
    # inverted(0) lands at 1.748e-7, i.e. bin 0, not bin M-1
    hdr, r_inv, _ = oracle.publish(raw[0], oracle.scan_params(0, 1, 1, 0, 40.0, 0.1))
    assert r_inv[0] == raw[0]["dist_mm_q2"][0] / np.float32(4000.0) or r_inv[0] < np.inf


def test_angle_rad_strictly_monotonic_and_below_two_pi(oracle):
    """The wrap branches at reference rplidar_node.cpp:592-597 are unreachable for u16
    input, and the float sort key is strictly monotonic in angle_z_q14."""
    k = np.arange(65536, dtype=np.float32)
    rad = ((k * np.float32(90.0) / np.float32(16384.0)).astype(np.float64) * (np.pi / 180.0)).astype(np.float32)
    assert (np.diff(rad) > 0).all()
    assert rad.max().astype(np.float64) < 2.0 * np.pi and rad.min() >= 0
