"""bench.py builds the synthetic wire input of its decode / chain workloads itself (the oracle is only their
CPU baseline); those builders must produce exactly what the oracle-side builders produce.  CPU only."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("fmt", [0x82, 0x83, 0x84, 0x85, 0x86])
def test_capsule_builders_equal_the_oracle_side_builders(bench, oracle, fmt):
    rng = np.random.default_rng(fmt)
    cb, per = bench.WIRE_CAPSULE[fmt]
    assert (cb, per) == (oracle.capsule_bytes(fmt), oracle.capsule_nodes(fmt))
    n = 50
    payload = rng.integers(0, 256, (n, cb), dtype=np.uint8)
    q6 = rng.integers(0, 360 * 64, n).astype(np.uint32)
    sync = rng.random(n) < 0.1
    if fmt == 0x83:
        a, b = bench.wire_seal_capsules(fmt, payload), oracle.seal_capsules(fmt, payload)
    else:
        a, b = bench.wire_seal_capsules(fmt, payload, q6, sync), oracle.seal_capsules(fmt, payload, q6, sync)
    assert (a == b).all()
    nodes, status, _, _ = oracle.decode_capsules(fmt, a, 31)
    assert ((status & oracle.CAPSULE_OK) != 0).all()  # markers and checksums are what the decoders expect


def test_dense_builder(bench, oracle):
    rng = np.random.default_rng(1)
    q6 = (np.arange(64) * 288) % (360 * 64)
    sync = np.arange(64) % 20 == 0
    dist = rng.integers(0, 65536, (64, 40))
    assert (bench.wire_dense_capsules(q6, sync, dist) == oracle.make_dense_capsules(q6, sync, dist)).all()
