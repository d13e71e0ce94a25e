"""ros2_overlay/ as files: the reference's node and wrapper with ros2_overlay/patches applied (to temporary copies
made by ros2_overlay/apply.sh during `make -C oracle ref`; no reference source is stored in this repository),
compiled against the ROS 2 API stubs and linked with librplidar_b200.so -> oracle/_ref/node_overlay_check
(tests/cpp/node_overlay_check.cpp).  CPU: the patched node builds and its publish_scan refuses to run without a
B200.  GPU: every LaserScan the patched node publishes equals what the reference's own publish_scan publishes
(oracle/_ref/libref_node.so), and every PointCloud2 equals the cloud definition, for 16 scans x 8 configurations."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "node_overlay_check")


def _env():
    env = dict(os.environ)
    import torch

    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    env["LD_LIBRARY_PATH"] = ":".join(p for p in ("/usr/local/cuda/lib64", libdir, env.get("LD_LIBRARY_PATH", "")) if p)
    return env


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        if not os.path.isdir("/root/reference/src/sdk/src"):
            pytest.skip("oracle/_ref/node_overlay_check not built (reference tree absent on this box)")
        import rplidar_ros2_driver_b200 as R
        from oracle import pyoracle

        if not os.path.exists(R.capi.LIB_PATH):
            R.build()
        pyoracle.build(ref=True)
    assert os.path.exists(EXE)
    return EXE


def test_patches_apply_to_the_reference_and_touch_only_the_seam():
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree absent on this box")
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run(["bash", os.path.join(ROOT, "ros2_overlay", "apply.sh"), "/root/reference", tmp],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        changed = 0
        for rel in ("src/rplidar_node.cpp", "src/lidar_driver_wrapper.cpp", "include/rplidar_node.hpp",
                    "include/lidar_driver_wrapper.hpp"):
            a = open(os.path.join("/root/reference", rel)).read().splitlines()
            b = open(os.path.join(tmp, rel)).read().splitlines()
            removed = [l for l in a if l not in b]
            changed += len(b) - len(a) + len(removed)
            assert len(removed) <= 1, (rel, removed)  # the only line that goes: the SDK's ascendScanData call
        assert changed < 40


def test_patched_node_builds_and_has_no_cpu_fallback(exe):
    r = subprocess.run([exe, "cpu"], capture_output=True, text=True, env=_env(), timeout=120)
    assert r.returncode == 0 and "OK cpu" in r.stdout, r.stdout + r.stderr
    import torch

    if not torch.cuda.is_available():
        assert "refused: no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_patched_node_publishes_the_reference_laserscan_and_the_cloud(exe, oracle, golden_dir, tmp_path):
    raw_arr = np.load(f"{golden_dir}/dummy_scans.npz")["raw"]
    raw, out = str(tmp_path / "raw.bin"), str(tmp_path / "out.bin")
    raw_arr.tofile(raw)
    r = subprocess.run([exe, "gpu", raw, out], capture_output=True, text=True, env=_env(), timeout=300)
    assert r.returncode == 0 and "OK gpu: 128 message pairs" in r.stdout, r.stdout + r.stderr
    buf = open(out, "rb").read()
    scans = np.ascontiguousarray(raw_arr).view(oracle.NODE_DTYPE).reshape(16, 360)
    pos = 0
    for _ in range(128):
        s, cfg, beams = np.frombuffer(buf, "<i4", 2, pos).tolist() + [int(np.frombuffer(buf, "<u4", 1, pos + 8)[0])]
        pos += 12
        hdr = np.frombuffer(buf, "<f4", 7, pos); pos += 28
        ranges = np.frombuffer(buf, "<f4", beams, pos); pos += 4 * beams
        inten = np.frombuffer(buf, "<f4", beams, pos); pos += 4 * beams
        width = int(np.frombuffer(buf, "<u4", 1, pos)[0]); pos += 4
        cloud = np.frombuffer(buf, "<f4", 4 * width, pos).reshape(width, 4); pos += 16 * width
        nodes = scans[s].copy()
        if s % 3 == 1:
            nodes["dist_mm_q2"][::7] = 0
        newp, mode_a, inv = cfg & 1, (cfg >> 1) & 1, (cfg >> 2) & 1
        rmax = 12.0 + s
        pub, ehdr, er, ei = oracle.ref_publish(nodes, oracle.scan_params(newp, mode_a, inv, 1, rmax, 0.1 + 0.001 * s))
        assert pub and len(er) == beams
        assert (hdr.view(np.uint32) == ehdr.view(np.uint32)).all(), (s, cfg, hdr, ehdr)
        assert (ranges.view(np.uint32) == er.view(np.uint32)).all() and (inten.view(np.uint32) == ei.view(np.uint32)).all()
        ec = oracle.cloud(nodes, oracle.cloud_params(range_min=0.15, range_max=rmax, voxel_size=0.05, sor_k=8,
                                                     sor_alpha=1.0, is_new_protocol=newp))
        assert ec.shape[0] == width and (cloud.view(np.uint32) == ec.view(np.uint32)).all(), (s, cfg)
    assert pos == len(buf)
