"""GPU tests of the message -> CDR kernels against oracle/cdr_oracle.py (parity
unpinned, see its header): byte-for-byte equality with the numpy writer and a parse-back of every message."""
import numpy as np
import pytest

from oracle import cdr_oracle as cdr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import rplidar_ros2_driver_b200 as R

    return R


@pytest.mark.parametrize("frame_id", ["laser_frame", "", "x" * 255, "lidar_3"])
def test_laserscan_batch_from_the_scan_kernel(R, oracle, frame_id):
    """scan kernel -> CDR on the device; each serialised message must parse back to the oracle's LaserScan."""
    import torch

    S, N = 48, 3200
    ctx = R.Context(0, N, S)
    dev = torch.device("cuda")
    hnodes = oracle.synth_batch(7, S, N, variant=1)
    counts_h = np.full(S, N, np.uint32)
    counts_h[3], counts_h[9] = 0, 17
    nodes = torch.from_numpy(hnodes.view(np.uint8).reshape(S, N, 8)).to(dev)
    counts = torch.from_numpy(counts_h.view(np.int32)).to(dev)
    ranges = torch.zeros((S, N), dtype=torch.float32, device=dev)
    intens = torch.zeros((S, N), dtype=torch.float32, device=dev)
    beams = torch.zeros(S, dtype=torch.int32, device=dev)
    inc = torch.zeros(S, dtype=torch.float32, device=dev)
    params = R.scan_params(0, 0, 0, 1)
    torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
    ctx.scan_batch_dev(nodes.data_ptr(), counts.data_ptr(), S, N, params, ranges=ranges.data_ptr(),
                       intensities=intens.data_ptr(), beam_counts=beams.data_ptr(), angle_increment=inc.data_ptr())
    meta_h = np.zeros(S, R.capi.LASERSCAN_META_DTYPE)
    rng = np.random.default_rng(1)
    meta_h["stamp_sec"] = rng.integers(-5, 2_000_000_000, S)
    meta_h["stamp_nanosec"] = rng.integers(0, 1_000_000_000, S)
    for k in ("angle_min", "angle_max", "angle_increment", "time_increment", "scan_time", "range_min", "range_max"):
        meta_h[k] = rng.random(S).astype(np.float32)
    meta = torch.from_numpy(meta_h.view(np.uint8)).to(dev)
    cdr_stride = (R.lib().rpl_laserscan_cdr_size(len(frame_id), N) + 15) & ~15
    out = torch.full((S, cdr_stride), 0xEE, dtype=torch.uint8, device=dev)
    sizes = torch.zeros(S, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
    ctx.laserscan_cdr_batch_dev(meta.data_ptr(), frame_id, ranges.data_ptr(), intens.data_ptr(), beams.data_ptr(), S, N,
                                out.data_ptr(), cdr_stride, cdr_sizes=sizes.data_ptr(), angle_increment=inc.data_ptr())
    ctx.synchronize()
    torch.cuda.synchronize()
    ho, hs, hb = out.cpu().numpy(), sizes.cpu().numpy(), beams.cpu().numpy()
    hr, hi, hinc = ranges.cpu().numpy(), intens.cpu().numpy(), inc.cpu().numpy()
    for s in range(S):
        n = int(hb[s])
        m = meta_h[s]
        expect = cdr.laserscan_cdr(int(m["stamp_sec"]), int(m["stamp_nanosec"]), frame_id,
                                   [m["angle_min"], m["angle_max"], hinc[s], m["time_increment"], m["scan_time"],
                                    m["range_min"], m["range_max"]], hr[s, :n], hi[s, :n])
        assert hs[s] == len(expect)
        assert ho[s, : hs[s]].tobytes() == expect, s
        assert (ho[s, hs[s]:] == 0xEE).all()  # nothing written past the message
        parsed = cdr.parse_laserscan(ho[s, : hs[s]].tobytes())
        assert parsed["frame_id"] == frame_id and len(parsed["ranges"]) == n
    ctx.close()


@pytest.mark.parametrize("frame_id", ["laser_frame", "abc", ""])
def test_pointcloud2_batch_from_the_cloud_path(R, oracle, frame_id):
    import torch

    S, N = 20, 3200
    ctx = R.Context(0, N, S)
    dev = torch.device("cuda")
    hnodes = oracle.synth_batch(11, S, N, variant=4)
    nodes = torch.from_numpy(hnodes.view(np.uint8).reshape(S, N, 8)).to(dev)
    counts = torch.full((S,), N, dtype=torch.int32, device=dev)
    xyzi = torch.zeros((S, N, 4), dtype=torch.float32, device=dev)
    pcount = torch.zeros(S, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
    ctx.cloud_batch_dev(nodes.data_ptr(), counts.data_ptr(), S, N, R.cloud_params(0.15, 30.0, 0.0, 0.0, 0, 0.0, 1),
                        xyzi.data_ptr(), pcount.data_ptr())
    stamps_h = np.stack([np.arange(S) + 100, np.arange(S) * 1000], axis=1).astype(np.uint32)
    stamps = torch.from_numpy(stamps_h.view(np.int32)).to(dev)
    cdr_stride = (R.lib().rpl_pointcloud2_cdr_size(len(frame_id), N) + 15) & ~15
    out = torch.full((S, cdr_stride), 0xEE, dtype=torch.uint8, device=dev)
    sizes = torch.zeros(S, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
    ctx.pointcloud2_cdr_batch_dev(stamps.data_ptr(), frame_id, xyzi.data_ptr(), pcount.data_ptr(), S, N,
                                  out.data_ptr(), cdr_stride, cdr_sizes=sizes.data_ptr())
    ctx.synchronize()
    torch.cuda.synchronize()
    ho, hs, hp, hx = out.cpu().numpy(), sizes.cpu().numpy(), pcount.cpu().numpy(), xyzi.cpu().numpy()
    assert hp.sum() > 0
    for s in range(S):
        n = int(hp[s])
        expect = cdr.pointcloud2_cdr(int(stamps_h[s, 0]), int(stamps_h[s, 1]), frame_id, hx[s, :n])
        assert hs[s] == len(expect)
        assert ho[s, : hs[s]].tobytes() == expect, s
        assert (ho[s, hs[s]:] == 0xEE).all()
        assert cdr.parse_pointcloud2(ho[s, : hs[s]].tobytes())["width"] == n
    ctx.close()


def test_device_writer_matches_the_hand_derived_bytes(R):
    """tests/golden/cdr_*.bin are laid out by hand from the XCDR1 rules (tests/golden/make_golden_cdr.py)."""
    import os

    import torch

    from test_cdr_golden import GOLD, _load

    m = _load()
    dev = torch.device("cuda")
    ctx = R.Context(0, 64, 4)
    g = m.LASERSCAN
    stride = 8
    ranges = torch.zeros((1, stride), dtype=torch.float32, device=dev)
    intens = torch.zeros((1, stride), dtype=torch.float32, device=dev)
    ranges[0, :3] = torch.tensor(g["ranges"], dtype=torch.float32)
    intens[0, :3] = torch.tensor(g["intensities"], dtype=torch.float32)
    beams = torch.tensor([3], dtype=torch.int32, device=dev)
    meta_h = np.zeros(1, R.capi.LASERSCAN_META_DTYPE)
    meta_h["stamp_sec"], meta_h["stamp_nanosec"] = g["sec"], g["nanosec"]
    for k, v in zip(("angle_min", "angle_max", "angle_increment", "time_increment", "scan_time", "range_min", "range_max"),
                    g["scalars"]):
        meta_h[k] = np.float32(v)
    meta = torch.from_numpy(meta_h.view(np.uint8)).to(dev)
    cdr_stride = (R.lib().rpl_laserscan_cdr_size(len(g["frame_id"]), stride) + 15) & ~15
    out = torch.full((1, cdr_stride), 0xEE, dtype=torch.uint8, device=dev)
    sizes = torch.zeros(1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
    ctx.laserscan_cdr_batch_dev(meta.data_ptr(), g["frame_id"], ranges.data_ptr(), intens.data_ptr(), beams.data_ptr(), 1,
                                stride, out.data_ptr(), cdr_stride, cdr_sizes=sizes.data_ptr())
    ctx.synchronize()
    torch.cuda.synchronize()
    gold = open(os.path.join(GOLD, "cdr_laserscan.bin"), "rb").read()
    assert int(sizes[0]) == len(gold) and out[0, : len(gold)].cpu().numpy().tobytes() == gold

    g = m.PC2
    xyzi = torch.zeros((1, stride, 4), dtype=torch.float32, device=dev)
    xyzi[0, :2] = torch.tensor(g["points"], dtype=torch.float32)
    pcount = torch.tensor([2], dtype=torch.int32, device=dev)
    stamps = torch.from_numpy(np.array([[g["sec"], g["nanosec"]]], np.uint32).view(np.int32)).to(dev)
    cdr_stride = (R.lib().rpl_pointcloud2_cdr_size(len(g["frame_id"]), stride) + 15) & ~15
    out = torch.full((1, cdr_stride), 0xEE, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
    ctx.pointcloud2_cdr_batch_dev(stamps.data_ptr(), g["frame_id"], xyzi.data_ptr(), pcount.data_ptr(), 1, stride,
                                  out.data_ptr(), cdr_stride, cdr_sizes=sizes.data_ptr())
    ctx.synchronize()
    torch.cuda.synchronize()
    gold = open(os.path.join(GOLD, "cdr_pointcloud2.bin"), "rb").read()
    assert int(sizes[0]) == len(gold) and out[0, : len(gold)].cpu().numpy().tobytes() == gold
    ctx.close()
