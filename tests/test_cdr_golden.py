"""The numpy CDR writer (oracle/cdr_oracle.py) against bytes laid out by hand from the XCDR1 rules
(tests/golden/make_golden_cdr.py -> cdr_*.bin): with the device writer compared against the same files in
tests/test_gpu_cdr.py, three independently written encoders have to agree byte for byte."""
import os

import numpy as np

from oracle import cdr_oracle as cdr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load():
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden_cdr", os.path.join(GOLD, "make_golden_cdr.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_committed_files_are_what_the_script_lays_out():
    m = _load()
    assert open(os.path.join(GOLD, "cdr_laserscan.bin"), "rb").read() == m.ls
    assert open(os.path.join(GOLD, "cdr_pointcloud2.bin"), "rb").read() == m.pc


def test_numpy_writer_matches_the_hand_derived_laserscan():
    m = _load()
    g = m.LASERSCAN
    got = cdr.laserscan_cdr(g["sec"], g["nanosec"], g["frame_id"], g["scalars"], np.array(g["ranges"], np.float32),
                            np.array(g["intensities"], np.float32))
    assert got == open(os.path.join(GOLD, "cdr_laserscan.bin"), "rb").read()
    back = cdr.parse_laserscan(got)
    assert back["frame_id"] == "laser" and len(back["ranges"]) == 3 and np.isinf(back["ranges"][1])


def test_numpy_writer_matches_the_hand_derived_pointcloud2():
    m = _load()
    g = m.PC2
    got = cdr.pointcloud2_cdr(g["sec"], g["nanosec"], g["frame_id"], np.array(g["points"], np.float32))
    assert got == open(os.path.join(GOLD, "cdr_pointcloud2.bin"), "rb").read()
    assert cdr.parse_pointcloud2(got)["width"] == 2
