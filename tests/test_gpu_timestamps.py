"""GPU tests of the per-sample timestamps against oracle/timestamp_oracle.cpp,
which tests/test_timestamps_vs_ref.py pins against the SDK's unpackers on a settable clock.
64-bit integer work: bit-exact."""
import numpy as np
import pytest

from test_capsule_oracle_vs_ref import make_capsules
from test_decode_oracle_vs_ref import make_stream
from test_timestamps_vs_ref import TIMINGS, rx_times

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import rplidar_ros2_driver_b200 as R

    return R


@pytest.fixture(scope="module")
def ctx(R):
    c = R.Context(0, 8192, 64)
    yield c
    c.close()


@pytest.mark.parametrize("ans", [0x82, 0x83, 0x84, 0x85, 0x86])
def test_capsule_streams_decode_stamp_assemble(R, oracle, ctx, ans):
    """decode -> stamps -> scan assembly with scan-begin stamps, batched and ragged, one timing per run."""
    import torch

    O = oracle
    dev = torch.device("cuda")
    cb, per = O.capsule_bytes(ans), O.capsule_nodes(ans)
    n_streams, n_caps, max_nodes, max_scans = 12, 400, 8192, 8
    mk = (lambda s: make_stream(O, n_caps, 80.0, seed=40 + s, sync_every=170)) if ans == 0x85 else \
         (lambda s: make_capsules(O, ans, n_caps, 3200.0 / per, seed=40 + s, sync_every=170))
    host = np.stack([mk(s) for s in range(n_streams)])
    host[2, 33, 20] ^= 0x20
    counts_h = np.full(n_streams, n_caps, np.uint32)
    counts_h[5] = 77
    rx_h = np.stack([rx_times(n_caps, 90 + s) for s in range(n_streams)])
    for timing in TIMINGS[:3]:
        t = R.Timing(*timing)
        t4 = O.timing4(*timing)
        caps = torch.from_numpy(host).to(dev)
        counts = torch.from_numpy(counts_h.view(np.int32)).to(dev)
        rx = torch.from_numpy(rx_h.view(np.int64)).to(dev)
        nodes = torch.zeros((n_streams, n_caps * per, 8), dtype=torch.uint8, device=dev)
        ncount = torch.zeros(n_streams, dtype=torch.int32, device=dev)
        status = torch.zeros((n_streams, n_caps), dtype=torch.int32, device=dev)
        offs = torch.zeros((n_streams, n_caps), dtype=torch.int32, device=dev)
        ts = torch.zeros((n_streams, n_caps * per), dtype=torch.int64, device=dev)
        scans = torch.zeros((n_streams, max_scans, max_nodes, 8), dtype=torch.uint8, device=dev)
        slen = torch.zeros((n_streams, max_scans), dtype=torch.int32, device=dev)
        sps = torch.zeros(n_streams, dtype=torch.int32, device=dev)
        sts = torch.zeros((n_streams, max_scans), dtype=torch.int64, device=dev)
        torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
        ctx.decode_capsules_batch_dev(ans, caps.data_ptr(), counts.data_ptr(), n_streams, n_caps, timing[0],
                                      nodes.data_ptr(), ncount.data_ptr(), capsule_status=status.data_ptr(),
                                      capsule_node_offset=offs.data_ptr())
        torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
        ctx.node_timestamps_dev(ans, t, rx.data_ptr(), status.data_ptr(), offs.data_ptr(), counts.data_ptr(),
                                n_streams, n_caps, ts.data_ptr())
        torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
        ctx.assemble_scans_dev(nodes.data_ptr(), ncount.data_ptr(), n_streams, n_caps * per, max_nodes, max_scans,
                               max_nodes, scans.data_ptr(), slen.data_ptr(), sps.data_ptr(),
                               capsule_status=status.data_ptr(), capsule_node_offset=offs.data_ptr(),
                               capsule_counts=counts.data_ptr(), stride_capsules=n_caps, node_ts_us=ts.data_ptr(),
                               scan_begin_ts_us=sts.data_ptr())
        ctx.synchronize()
        torch.cuda.synchronize()
        hts = ts.cpu().numpy().view(np.uint64)
        hsts = sts.cpu().numpy().view(np.uint64)
        hsps = sps.cpu().numpy()
        published = 0
        for s in range(n_streams):
            k = int(counts_h[s])
            en, es, eo, _ = O.decode_capsules(ans, host[s, :k], timing[0])
            ets = O.node_timestamps(ans, t4, rx_h[s, :k], es, eo, len(en))
            assert int(ncount[s]) == len(en)
            assert (hts[s, : len(en)] == ets).all(), (hex(ans), s)
            _, elen, ek, escan_ts = O.assemble_scans_ts(en, ets, O.resets_from_capsules(es, eo), max_nodes, max_scans)
            assert hsps[s] == ek
            assert (hsts[s, : min(ek, max_scans)] == escan_ts[: min(ek, max_scans)]).all()
            published += ek
        assert published > 0 or ans == 0x83  # HQ payload is random: scan-start flags everywhere or nowhere


def test_standard_node_stamps(R, oracle, ctx):
    import torch

    O = oracle
    dev = torch.device("cuda")
    rng = np.random.default_rng(5)
    n_streams, n = 6, 3000
    streams = []
    for s in range(n_streams):
        rec = np.zeros((n, 5), np.uint8)
        sb = (np.arange(n) % 360 == 0).astype(np.uint8)
        rec[:, 0] = (rng.integers(0, 64, n).astype(np.uint8) << 2) | ((1 - sb) << 1) | sb
        w = (rng.integers(0, 360 * 64, n).astype(np.uint16) << 1) | 1
        rec[:, 1], rec[:, 2] = w & 0xFF, w >> 8
        rec[:, 3:] = rng.integers(0, 256, (n, 2))
        b = rec.reshape(-1).copy()
        b[rng.choice(len(b), 60, replace=False)] ^= 0xFF
        streams.append(b)
    host = np.stack(streams)
    stride = host.shape[1]
    chunk = 64
    n_chunks = (stride + chunk - 1) // chunk
    rx_h = np.stack([rx_times(n_chunks, 7 + s) for s in range(n_streams)])
    wire = torch.from_numpy(host).to(dev)
    counts = torch.full((n_streams,), stride, dtype=torch.int32, device=dev)
    nodes = torch.zeros((n_streams, stride // 5, 8), dtype=torch.uint8, device=dev)
    ncount = torch.zeros(n_streams, dtype=torch.int32, device=dev)
    ends = torch.zeros((n_streams, stride // 5), dtype=torch.int32, device=dev)
    rx = torch.from_numpy(rx_h.view(np.int64)).to(dev)
    ts = torch.zeros((n_streams, stride // 5), dtype=torch.int64, device=dev)
    timing = TIMINGS[3]
    torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
    ctx.decode_normal_batch_dev(wire.data_ptr(), counts.data_ptr(), n_streams, stride, nodes.data_ptr(),
                                ncount.data_ptr(), node_end=ends.data_ptr())
    torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
    ctx.normal_timestamps_dev(R.Timing(*timing), ends.data_ptr(), ncount.data_ptr(), n_streams, stride // 5, chunk,
                              rx.data_ptr(), n_chunks, ts.data_ptr())
    ctx.synchronize()
    torch.cuda.synchronize()
    hts = ts.cpu().numpy().view(np.uint64)
    hends = ends.cpu().numpy().view(np.uint32)
    for s in range(n_streams):
        en, eend, _ = O.decode_normal(host[s])
        assert int(ncount[s]) == len(en) and (hends[s, : len(en)] == eend).all()
        ets = O.normal_timestamps(O.timing4(*timing), eend, chunk, rx_h[s])
        assert (hts[s, : len(en)] == ets).all()
