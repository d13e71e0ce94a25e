"""CPU proofs of the arithmetic shortcuts the CUDA kernels take (rpl_device.cuh):

  dist_to_m        dist_mm_q2 / 4000.0f as one multiply + two FMAs  -> exhaustive C check
  deg_to_key       v * 16384 / 90 with the division replaced by a reciprocal refinement
                                                                     -> exhaustive C check
  mode_a_bin_fast  Mode A bin index by integer arithmetic wherever the exact ratio is further
                   than twice the float chain's error bound from a bin edge -> numpy sweep vs the
                   reference's float chain (rplidar_node.cpp:586-652) for every key
"""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32 = np.float32
TWO_PI = 2.0 * np.pi


def test_division_by_4000_is_exact_for_every_u32(tmp_path):
    src = os.path.join(ROOT, "oracle", "check_div4000.c")
    exe = str(tmp_path / "check_div4000")
    flags = ["-O2", "-ffp-contract=off"]
    cpu = open("/proc/cpuinfo").read()
    if " fma " in cpu or " fma\n" in cpu:
        flags.append("-mfma")
    subprocess.run(["gcc", *flags, src, "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    assert "mismatches 0" in r.stdout and "checked 83886082 values" in r.stdout


def test_division_by_90_is_exact_for_every_angle(tmp_path):
    """deg_to_key: v * 16384 / 90 through the reciprocal, for every float in [0, 1024] (1.15e9 values, ~12 s)."""
    src = os.path.join(ROOT, "oracle", "check_div90.c")
    exe = str(tmp_path / "check_div90")
    flags = ["-O2", "-ffp-contract=off"]
    cpu = open("/proc/cpuinfo").read()
    if " fma " in cpu or " fma\n" in cpu:
        flags.append("-mfma")
    subprocess.run(["gcc", *flags, src, "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout
    assert "mismatches 0, key mismatches 0" in r.stdout and "checked 1149239297 values" in r.stdout


def test_ultra_angle_table_index_by_float_division():
    """decode_formats.cu ultra_cabin: 98361 / dist_q2 as trunc(RN(98361.0f / (float)dist_q2)) for every distance the
    format can carry (dist_q2 < 2^22)."""
    d = np.arange(200, 1 << 22, dtype=np.uint32)
    q = (F32(98361.0) / d.astype(F32)).astype(np.uint32)
    assert (q == 98361 // d).all() and int(q.max()) == 491


def exact_bins(m: int, inverted: bool) -> np.ndarray:
    k = np.arange(65536, dtype=np.float32)
    deg = (k * F32(90.0)) / F32(16384.0)
    rad = (deg.astype(np.float64) * (np.pi / 180.0)).astype(F32)
    inc = F32(TWO_PI / float(m))
    a = rad
    if inverted:
        a = (TWO_PI - a.astype(np.float64)).astype(F32)
        wrap = a.astype(np.float64) >= TWO_PI
        a = np.where(wrap, (a.astype(np.float64) - TWO_PI).astype(F32), a)
    return ((a - F32(0.0)) / inc).astype(np.int64)


def fast_zone_bins(m: int, inverted: bool):
    """(bins, mask): integer-quotient bins and the keys for which the kernel trusts them."""
    k = np.arange(65536, dtype=np.uint64)
    kk = (65536 - k) if inverted else k
    t = kk * np.uint64(m)
    frac = (t & np.uint64(0xFFFF)).astype(np.int64)
    g = (m >> 5) + 2  # mode_a_bin_fast: twice the error bound M/64 of the float chain, in 2^-16 bin
    ok = (frac >= g) & (frac <= 65536 - g)
    if inverted:
        ok &= k != 0
    return (t >> np.uint64(16)).astype(np.int64), ok


def test_mode_a_integer_bins_match_the_float_chain():
    rng = np.random.default_rng(42)
    ms = set(range(1, 1537)) | {65536, 65535, 65534, 32768, 32767, 31117, 31130, 49152, 360, 3200, 8192, 8191, 4096, 4095}
    ms |= set(range(1537, 8193, 7)) | set(range(3100, 3300))  # the sizes the shared-memory kernels see
    ms |= set(int(x) for x in rng.integers(1537, 65537, size=1200))
    worst = 0
    for m in sorted(ms):
        for inverted in (False, True):
            ref = exact_bins(m, inverted)
            fast, ok = fast_zone_bins(m, inverted)
            assert (ref[ok] == fast[ok]).all(), (m, inverted)
            # the reference's bounds check (index < beam_count) never fires for u16 keys
            assert ref.max() < m and ref.min() >= 0, (m, inverted)
            worst = max(worst, int((~ok).sum()))
    assert worst <= 65536  # informational: keys that take the exact chain (at most all of them)


def test_mode_a_bins_are_monotonic_in_scan_order():
    """The head/tail/gap logic of the fast kernels needs bins to be non-decreasing along
    ascending keys (and, inverted, along key 0 followed by descending keys)."""
    for m in (1, 2, 7, 360, 3200, 31117, 65536):
        b = exact_bins(m, False)
        assert (np.diff(b) >= 0).all()
        bi = exact_bins(m, True)
        order = np.concatenate([[0], np.arange(65535, 0, -1)])
        assert (np.diff(bi[order]) >= 0).all()
        assert bi[0] == 0
