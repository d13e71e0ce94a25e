"""The INTEGRATION.md replacement of RPlidarNode::publish_scan, built against the reference's OWN node class
(src/rplidar_node.cpp included in place, ROS 2 API stubs from oracle/ros_stubs/) and run next to the
reference's own publish_scan: tests/cpp/node_dropin_check.cpp.  The binary is built with the other checkers
(oracle/Makefile ref -> oracle/_ref/node_dropin_check, in the authoring container where /root/reference exists)
and travels to the GPU box.  Runs last (file name): it is the one test that exercises the whole seam at once."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "node_dropin_check")


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
            pytest.skip("oracle/_ref/node_dropin_check not built (reference tree absent on this box)")
        import rplidar_ros2_driver_b200 as R
        from oracle import pyoracle

        if not os.path.exists(R.capi.LIB_PATH):
            R.build()
        pyoracle.build(ref=True)
    assert os.path.exists(EXE)
    return EXE


def test_patched_publish_scan_compiles_against_the_reference_node(exe):
    r = subprocess.run([exe, "cpu"], capture_output=True, text=True, env=_env(), timeout=120)
    assert r.returncode == 0 and "OK cpu" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_patched_publish_scan_publishes_what_the_reference_node_publishes(exe, golden_dir, tmp_path):
    raw = str(tmp_path / "raw.bin")
    np.load(f"{golden_dir}/dummy_scans.npz")["raw"].tofile(raw)
    r = subprocess.run([exe, "gpu", raw], capture_output=True, text=True, env=_env(), timeout=300)
    assert r.returncode == 0 and "OK gpu: 128 messages" in r.stdout, r.stdout + r.stderr
