"""Generates tests/golden/wire_golden.npz from the COMPILED REFERENCE (oracle/_ref): outputs of the SDK's own
LIDARSampleDataUnpacker (with the settable clock of libref_clock.so) and of its ScanDataHolder on small
wire streams of every measurement answer type.  Run in the authoring
container only:

    python tests/golden/make_golden_wire.py

Per answer type T in {0x81..0x86}:
  wire_T        the byte stream fed to the unpacker (capsule formats: whole capsules, with two checksum
                errors and scan-start capsules; 0x81: records with corrupted bytes)
  rx_T          receive time of every piece (one piece per capsule; 64-byte pieces for 0x81)
  nodes_T       the nodes the SDK decoded               ts_T     the timestamp it attached to each
  events_T      [n,3] (kind, nodes decoded so far, code) scan resets (1) and decoding errors (2)
  scan_len_T    lengths of the scans its ScanDataHolder (capacity 256) published from that node stream
  scan_ts_T     their scan-begin timestamps             timing   the SlamtecLidarTimingDesc used
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as O  # noqa: E402
from test_capsule_oracle_vs_ref import make_capsules  # noqa: E402
from test_decode_oracle_vs_ref import make_stream  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TIMING = (63, 256000, 17, 0)
HOLDER_CAP = 256  # ScanDataHolder capacity used for the scan_len / scan_ts vectors (exercises the overwrite rule)


def wire_for(ans: int) -> tuple[np.ndarray, int]:
    rng = np.random.default_rng(ans)
    if ans == 0x81:
        n = 1500
        rec = np.zeros((n, 5), np.uint8)
        s = (np.arange(n) % 360 == 0).astype(np.uint8)
        rec[:, 0] = (rng.integers(0, 64, n).astype(np.uint8) << 2) | ((1 - s) << 1) | s
        w = (((np.arange(n) % 360) * 64).astype(np.uint16) << 1) | 1
        rec[:, 1], rec[:, 2] = w & 0xFF, w >> 8
        rec[:, 3:] = rng.integers(0, 256, (n, 2))
        b = rec.reshape(-1).copy()
        b[rng.choice(len(b), 25, replace=False)] ^= 0xFF
        return b, 64
    n = 140
    per = O.capsule_nodes(ans)
    if ans == 0x85:
        caps = make_stream(O, n, 25.0, seed=ans, sync_every=60)
    else:
        caps = make_capsules(O, ans, n, 1000.0 / per, seed=ans, sync_every=60, near=True)
    caps[[33, 90], 20] ^= 0x10
    return caps.reshape(-1), O.capsule_bytes(ans)


def main():
    O.build(ref=True)
    assert O.have_ref() and O.have_ref_clock() and O.have_ref_holder(), "oracle/_ref missing"
    t4 = O.timing4(*TIMING)
    g = {"timing": t4, "holder_cap": np.uint32(HOLDER_CAP)}
    for ans in (0x81, 0x82, 0x83, 0x84, 0x85, 0x86):
        wire, piece = wire_for(ans)
        n_pieces = (len(wire) + piece - 1) // piece
        rx = (5_000_000 + np.cumsum(np.random.default_rng(1000 + ans).integers(300, 2500, n_pieces))).astype(np.uint64)
        nodes, ts = O.ref_unpack_ts(ans, wire, piece, rx, t4)
        nodes2, events = O.ref_unpack(ans, wire, TIMING[0], piece)
        assert (nodes.view(np.uint64) == nodes2.view(np.uint64)).all()
        resets = events[events[:, 0] == 1, 1].astype(np.uint32)
        _, lens, k, sts = O.ref_assemble_scans_ts(nodes, ts, resets, HOLDER_CAP, 8192)
        assert k <= 8192
        t = f"{ans:02x}"
        g[f"wire_{t}"], g[f"rx_{t}"] = wire, rx
        g[f"nodes_{t}"] = nodes.view(np.uint8).reshape(-1, 8)
        g[f"ts_{t}"], g[f"events_{t}"] = ts, events
        g[f"scan_len_{t}"], g[f"scan_ts_{t}"] = lens[:k], sts[:k]
        print(f"{ans:#x}: {len(wire)} bytes -> {len(nodes)} nodes, {len(events)} events, {k} scans")
    np.savez_compressed(os.path.join(OUT, "wire_golden.npz"), **g)


if __name__ == "__main__":
    main()
