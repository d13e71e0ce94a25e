"""Driver for compute-sanitizer --tool racecheck: the shared-memory scan kernels WITHOUT their bulk-TMA staging (odd
stride -> ordinary loads; racecheck does not model async-proxy writes ordered by mbarriers and reports every ring slot
of the TMA kernels as a hazard), every LaserScan variant, the fused PointCloud2 chain, and the capsule decoders.
Small inputs: the tool slows kernels down by two orders of magnitude.  Results are also checked against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import rplidar_ros2_driver_b200 as R  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

O.build(ref=False)
n, stride, S = 1200, 1201, 6
base = O.synth_batch(99, S, n, 0)
base[1] = O.synth_batch(5, 1, n, 2)[0]            # duplicate keys
base[2] = O.synth_batch(6, 1, n, 3)[0]            # shuffled
nodes = np.zeros((S, stride), O.NODE_DTYPE)
nodes[:, :n] = base
counts = np.array([n, n, n, n - 1, 17, n], np.uint32)
ctx = R.Context(0, stride, S)
bad = 0
for newp, mode_a, inv in ((0, 0, 0), (1, 1, 1), (0, 1, 0), (1, 0, 1)):
    for emit in (False, True):
        got = ctx.scan_batch(nodes.view(R.NODE_DTYPE), counts, R.scan_params(newp, mode_a, inv, 1), emit_nodes=emit)
        buf = nodes.copy()
        exp = O.pipeline_batch(buf, counts, O.scan_params(newp, mode_a, inv, 1, 40.0, 0.1), stable=True, threads=2)
        for s in range(S):
            m = int(exp["beam_counts"][s])
            ok = int(got["beam_counts"][s]) == m and (got["ranges"][s, :m].view(np.uint32) == exp["ranges"][s, :m].view(np.uint32)).all() \
                and (got["intensities"][s, :m].view(np.uint32) == exp["intensities"][s, :m].view(np.uint32)).all()
            if emit:
                ok = ok and (got["nodes"][s, : counts[s]].view(np.uint64) == buf[s, : counts[s]].view(np.uint64)).all()
            bad += 0 if ok else 1
room = np.zeros((S, stride), O.NODE_DTYPE)
room[:, :n] = O.synth_batch(7, S, n, 4)
for sor in (0, 8):
    prm = R.cloud_params(range_min=0.15, range_max=40.0, voxel_size=0.05, sor_k=sor)
    xyzi, pc = ctx.cloud_batch(room.view(R.NODE_DTYPE), counts, prm)
    bad += 0 if int(pc.sum()) > 0 else 1
ctx.close()
ctx = R.Context(0, 4096, 4)
rng = np.random.default_rng(3)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_capsule_oracle_vs_ref import make_capsules  # noqa: E402

for ans in (0x82, 0x84, 0x86):
    caps = make_capsules(O, ans, 300, 40.0, seed=77 + ans, sync_every=90)
    gn = ctx.decode_capsules(ans, caps)
    en = O.decode_capsules(ans, caps, 31, (0, 0))[0]
    g = gn[0]
    bad += 0 if (len(g) == len(en) and (g.view(np.uint64) == en.view(np.uint64)).all()) else 1
ctx.close()
print("RACECHECK_DRIVER_DONE mismatches", bad)
sys.exit(1 if bad else 0)
