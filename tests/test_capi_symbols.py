"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every
symbol include/rpl_b200.h declares; without a CUDA device the product refuses to run (there
is no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def R():
    import rplidar_ros2_driver_b200 as R

    if not os.path.exists(R.capi.LIB_PATH):
        R.build()
    return R


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rpl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rpl_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(R):
    syms = declared_symbols()
    assert len(syms) >= 15
    assert sorted(R.capi.EXPORTS) == syms


def test_library_exports_every_declared_symbol(R):
    L = ctypes.CDLL(R.capi.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), f"librplidar_b200.so does not export {s}"
    assert R.lib().rpl_abi_version() == 1


def test_node_layout_matches_reference_struct(R):
    # reference src/sdk/include/sl_lidar_cmd.h:272-278: sizeof 8, offsets 0/2/6/7
    dt = R.NODE_DTYPE
    assert dt.itemsize == 8
    assert [dt.fields[n][1] for n in ("angle_z_q14", "dist_mm_q2", "quality", "flag")] == [0, 2, 6, 7]
    assert ctypes.sizeof(R.capi.ScanParams) == 8
    assert ctypes.sizeof(R.capi.CloudParams) == 28


def test_no_cpu_fallback(R):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(R.RplError) as e:
        R.Context(0, 1024, 1)
    assert e.value.code == 0x80008004  # OPERATION_NOT_SUPPORT


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "rplidar_ros2_driver_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".sh")):
                txt = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle.h" not in txt, f
