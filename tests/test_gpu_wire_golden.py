"""The device path of the wire-side steps against the vectors captured from the compiled reference
(tests/golden/wire_golden.npz): decode -> per-node stamps -> scan assembly with scan-begin stamps, for
every measurement answer type, through the C-ABI."""
import numpy as np
import pytest

from test_wire_golden import ANS, compare, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ans", ANS)
def test_device_chain_reproduces_the_captured_reference_outputs(oracle, golden_dir, ans):
    import torch

    import rplidar_ros2_driver_b200 as R

    g = load(golden_dir)
    t = f"{ans:02x}"
    wire, rx_h, t4 = g[f"wire_{t}"], g[f"rx_{t}"], g["timing"]
    cap = int(g["holder_cap"])
    timing = R.Timing(*[int(x) for x in t4])
    dev = torch.device("cuda")
    ctx = R.Context(0, 8192, 64)
    max_scans = 8192
    wire_d = torch.from_numpy(wire.copy()).to(dev)
    rx = torch.from_numpy(rx_h.view(np.int64).copy()).to(dev)
    if ans == 0x81:
        n_slots = len(wire) // 5
        counts = torch.tensor([len(wire)], dtype=torch.int32, device=dev)
        nodes = torch.zeros((n_slots, 8), dtype=torch.uint8, device=dev)
        ncount = torch.zeros(1, dtype=torch.int32, device=dev)
        ends = torch.zeros(n_slots, dtype=torch.int32, device=dev)
        ts = torch.zeros(n_slots, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
        ctx.decode_normal_batch_dev(wire_d.data_ptr(), counts.data_ptr(), 1, len(wire), nodes.data_ptr(),
                                    ncount.data_ptr(), node_end=ends.data_ptr())
        torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
        ctx.normal_timestamps_dev(timing, ends.data_ptr(), ncount.data_ptr(), 1, n_slots, 64, rx.data_ptr(), len(rx_h),
                                  ts.data_ptr())
        kw = {}
        stride_nodes = n_slots
    else:
        cb, per = oracle.capsule_bytes(ans), oracle.capsule_nodes(ans)
        n_caps = len(wire) // cb
        n_slots = n_caps * per
        counts = torch.tensor([n_caps], dtype=torch.int32, device=dev)
        nodes = torch.zeros((n_slots, 8), dtype=torch.uint8, device=dev)
        ncount = torch.zeros(1, dtype=torch.int32, device=dev)
        status = torch.zeros(n_caps, dtype=torch.int32, device=dev)
        offs = torch.zeros(n_caps, dtype=torch.int32, device=dev)
        ts = torch.zeros(n_slots, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
        ctx.decode_capsules_batch_dev(ans, wire_d.data_ptr(), counts.data_ptr(), 1, n_caps, int(t4[0]),
                                      nodes.data_ptr(), ncount.data_ptr(), capsule_status=status.data_ptr(),
                                      capsule_node_offset=offs.data_ptr())
        torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
        ctx.node_timestamps_dev(ans, timing, rx.data_ptr(), status.data_ptr(), offs.data_ptr(), counts.data_ptr(), 1,
                                n_caps, ts.data_ptr())
        kw = dict(capsule_status=status.data_ptr(), capsule_node_offset=offs.data_ptr(),
                  capsule_counts=counts.data_ptr(), stride_capsules=n_caps)
        stride_nodes = n_slots
    scans = torch.zeros((max_scans, cap, 8), dtype=torch.uint8, device=dev)
    slen = torch.zeros(max_scans, dtype=torch.int32, device=dev)
    sps = torch.zeros(1, dtype=torch.int32, device=dev)
    sts = torch.zeros(max_scans, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
    ctx.assemble_scans_dev(nodes.data_ptr(), ncount.data_ptr(), 1, stride_nodes, cap, max_scans, cap, scans.data_ptr(),
                           slen.data_ptr(), sps.data_ptr(), node_ts_us=ts.data_ptr(), scan_begin_ts_us=sts.data_ptr(),
                           **kw)
    ctx.synchronize()
    torch.cuda.synchronize()
    n = int(ncount[0])
    k = int(sps[0])
    hn = nodes.cpu().numpy().view(oracle.NODE_DTYPE).reshape(-1)[:n]
    compare(g, ans, hn, None, ts.cpu().numpy().view(np.uint64)[:n], slen.cpu().numpy().astype(np.uint32)[:k],
            sts.cpu().numpy().view(np.uint64)[:k])
    ctx.close()
