"""The C++17 host mirror (rplidar_ros2_driver_b200/host): compiled with g++ against the C-ABI
library and run -- on CPU for the interface + dummy generator, on the GPU for the whole seam."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rplidar_ros2_driver_b200")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    import rplidar_ros2_driver_b200 as R

    if not os.path.exists(R.capi.LIB_PATH):
        R.build()
    out = str(tmp_path_factory.mktemp("cpp") / "host_mirror_test")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I", os.path.join(PKG, "host"),
           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), os.path.join(PKG, "host", "lidar_driver_wrapper.cpp"),
           "-L", PKG, "-lrplidar_b200", f"-Wl,-rpath,{PKG}", "-Wl,--allow-shlib-undefined", "-lpthread", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


@pytest.fixture(scope="module")
def files(tmp_path_factory, golden_dir):
    d = tmp_path_factory.mktemp("golden_bin")
    g = np.load(f"{golden_dir}/dummy_scans.npz")
    ls = np.load(f"{golden_dir}/laserscan_golden.npz")
    paths = {k: str(d / f"{k}.bin") for k in ("raw", "asc", "ranges", "intens")}
    g["raw"].tofile(paths["raw"])
    g["variants_ascended"][0].tofile(paths["asc"])
    # golden config: variant 0, ascended input, old protocol, Mode A, not inverted
    for k in range(int(ls["n"])):
        if ls[f"cfg_{k}"].tolist() == [0, 1, 0, 1, 0]:
            ls[f"ranges_{k}"].tofile(paths["ranges"])
            ls[f"intens_{k}"].tofile(paths["intens"])
            break
    else:
        raise AssertionError("golden config not found")
    return paths


def _env():
    env = dict(os.environ)
    import torch

    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    env["LD_LIBRARY_PATH"] = ":".join(p for p in ("/usr/local/cuda/lib64", libdir, env.get("LD_LIBRARY_PATH", "")) if p)
    return env


def test_host_mirror_cpu(exe, files):
    r = subprocess.run([exe, "cpu", files["raw"]], capture_output=True, text=True, env=_env(), timeout=120)
    assert r.returncode == 0 and "OK cpu" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_mirror_gpu(exe, files):
    r = subprocess.run([exe, "gpu", files["raw"], files["asc"], files["ranges"], files["intens"]],
                       capture_output=True, text=True, env=_env(), timeout=300)
    assert r.returncode == 0 and "OK gpu" in r.stdout, r.stdout + r.stderr


# ---- the sample-data unpacker seam (SURVEY.md 8(f) rank 1 + 4) ------------------------------------------
@pytest.fixture(scope="module")
def unpacker_exe(tmp_path_factory):
    import rplidar_ros2_driver_b200 as R

    if not os.path.exists(R.capi.LIB_PATH):
        R.build()
    out = str(tmp_path_factory.mktemp("cpp") / "unpacker_mirror_test")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I", os.path.join(PKG, "host"),
           os.path.join(ROOT, "tests", "cpp", "unpacker_mirror_test.cpp"), "-L", PKG, "-lrplidar_b200",
           f"-Wl,-rpath,{PKG}", "-Wl,--allow-shlib-undefined", "-lpthread", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_unpacker_mirror_cpu(unpacker_exe):
    r = subprocess.run([unpacker_exe, "cpu"], capture_output=True, text=True, env=_env(), timeout=120)
    assert r.returncode == 0 and "OK cpu" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("ans", [0x82, 0x83, 0x84, 0x85, 0x86])
@pytest.mark.parametrize("batch", [1, 7, 256])
def test_unpacker_mirror_replays_the_sdk_callbacks(unpacker_exe, golden_dir, tmp_path, ans, batch):
    """Same callbacks, same order, same timestamps as the SDK's LIDARSampleDataUnpacker made on the captured
    streams (tests/golden/wire_golden.npz), whatever the batch size."""
    g = np.load(f"{golden_dir}/wire_golden.npz")
    t = f"{ans:02x}"
    paths = {}
    for k in ("wire", "rx", "nodes", "ts", "events"):
        paths[k] = str(tmp_path / f"{k}.bin")
        np.ascontiguousarray(g[f"{k}_{t}"]).tofile(paths[k])
    paths["timing"] = str(tmp_path / "timing.bin")
    g["timing"].astype(np.uint32).tofile(paths["timing"])
    r = subprocess.run([unpacker_exe, "gpu", hex(ans), str(batch), paths["wire"], paths["rx"], paths["nodes"], paths["ts"],
                        paths["events"], paths["timing"]], capture_output=True, text=True, env=_env(), timeout=300)
    assert r.returncode == 0 and "OK gpu" in r.stdout, r.stdout + r.stderr
