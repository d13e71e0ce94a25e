"""The restatements of the wire-side steps (decoders, scan assembly, timestamps)
against tests/golden/wire_golden.npz, which tests/golden/make_golden_wire.py captured from the compiled
reference (the SDK's own unpacker on a settable clock and its ScanDataHolder).  CPU only; unlike the
*_vs_ref tests this one also runs where oracle/_ref is absent."""
import os

import numpy as np
import pytest

ANS = [0x81, 0x82, 0x83, 0x84, 0x85, 0x86]


def load(golden_dir):
    return np.load(os.path.join(golden_dir, "wire_golden.npz"))


def oracle_chain(O, g, ans):
    """(nodes, events, node timestamps, scan lengths, scan-begin timestamps) from the restatements."""
    t = f"{ans:02x}"
    wire, rx, t4 = g[f"wire_{t}"], g[f"rx_{t}"], g["timing"]
    cap = int(g["holder_cap"])
    if ans == 0x81:
        nodes, ends, _ = O.decode_normal(wire)
        ts = O.normal_timestamps(t4, ends, 64, rx)
        events = np.zeros((0, 3), np.uint32)
        resets = np.zeros(0, np.uint32)
    else:
        nodes, status, offs, _ = O.decode_capsules(ans, wire, int(t4[0]))
        ts = O.node_timestamps(ans, t4, rx, status, offs, len(nodes))
        ev = []
        for st, off in zip(status.tolist(), offs.tolist()):
            if st & O.CAPSULE_CHECKSUM_ERR:
                ev.append((2, off, 0x8002))
            if st & O.CAPSULE_SYNC:
                if st & O.CAPSULE_ENCODER_RESET_ERR:
                    ev.append((2, off, 0x8001))
                ev.append((1, off, 0))
        events = np.array(ev, np.uint32).reshape(-1, 3)
        resets = O.resets_from_capsules(status, offs)
    _, lens, k, sts = O.assemble_scans_ts(nodes, ts, resets, cap, 8192)
    return nodes, events, ts, lens[:k], sts[:k]


def compare(g, ans, nodes, events, ts, lens, sts):
    t = f"{ans:02x}"
    gn = g[f"nodes_{t}"].reshape(-1).view(np.uint64)
    assert len(nodes) == len(gn) and (nodes.view(np.uint64) == gn).all()
    if events is not None:
        assert events.shape == g[f"events_{t}"].shape and (events == g[f"events_{t}"]).all()
    assert (ts == g[f"ts_{t}"]).all()
    assert len(lens) == len(g[f"scan_len_{t}"]) and (lens == g[f"scan_len_{t}"]).all()
    assert (sts == g[f"scan_ts_{t}"]).all()


@pytest.mark.parametrize("ans", ANS)
def test_restatement_reproduces_the_captured_reference_outputs(oracle, golden_dir, ans):
    g = load(golden_dir)
    compare(g, ans, *oracle_chain(oracle, g, ans))
