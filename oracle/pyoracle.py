"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module (see oracle/oracle.h).  The product package
rplidar_ros2_driver_b200 never does.

liboracle_scan.so  : the restatement (oracle/scan_oracle.cpp, oracle/cloud_oracle.cpp)
_ref/libref_rplidar.so : the UNMODIFIED reference SDK + wrapper compiled in place
                         (oracle/Makefile, oracle/ref_shim.cpp); optional.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle_scan.so")
REF_SO = os.path.join(HERE, "_ref", "libref_rplidar.so")
REF_HOLDER_SO = os.path.join(HERE, "_ref", "libref_holder.so")
REF_CLOCK_SO = os.path.join(HERE, "_ref", "libref_clock.so")
REF_NODE_SO = os.path.join(HERE, "_ref", "libref_node.so")

NODE_DTYPE = np.dtype(
    {
        "names": ["angle_z_q14", "dist_mm_q2", "quality", "flag"],
        "formats": ["<u2", "<u4", "u1", "u1"],
        "offsets": [0, 2, 6, 7],
        "itemsize": 8,
    }
)

RESULT_OK = 0
RESULT_OPERATION_FAIL = 0x80008001


class ScanHeader(C.Structure):
    _fields_ = [
        ("angle_min", C.c_float),
        ("angle_max", C.c_float),
        ("angle_increment", C.c_float),
        ("time_increment", C.c_float),
        ("scan_time", C.c_float),
        ("range_min", C.c_float),
        ("range_max", C.c_float),
        ("beam_count", C.c_uint32),
        ("published", C.c_uint32),
    ]


class ScanParams(C.Structure):
    _fields_ = [
        ("is_new_protocol", C.c_uint8),
        ("scan_processing", C.c_uint8),
        ("inverted", C.c_uint8),
        ("apply_ascend", C.c_uint8),
        ("range_max", C.c_float),
        ("scan_duration", C.c_double),
    ]


class CloudParams(C.Structure):
    _fields_ = [
        ("range_min", C.c_float),
        ("range_max", C.c_float),
        ("intensity_min", C.c_float),
        ("voxel_size", C.c_float),
        ("sor_k", C.c_uint32),
        ("sor_alpha", C.c_float),
        ("is_new_protocol", C.c_uint8),
        ("pad", C.c_uint8 * 3),
    ]


def build(ref: bool = True) -> None:
    """Compile the oracle (and, when /root/reference exists, oracle/_ref)."""
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if ref and os.path.isdir("/root/reference/src/sdk/src"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        vp, sz, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
        L.orc_ascend_scan.argtypes = [vp, sz, i32]
        L.orc_ascend_scan.restype = u32
        L.orc_publish_scan.argtypes = [vp, sz, C.POINTER(ScanParams), i32, vp, vp, C.POINTER(ScanHeader)]
        L.orc_publish_scan.restype = u32
        L.orc_pipeline_batch.argtypes = [vp, vp, u32, u32, C.POINTER(ScanParams), i32, vp, vp, vp, vp, vp, i32]
        L.orc_pipeline_batch.restype = C.c_double
        L.orc_dummy_scan.argtypes = [u32, vp]
        L.orc_dummy_scan.restype = None
        L.orc_synth_scan.argtypes = [C.c_uint64, u32, i32, vp]
        L.orc_synth_scan.restype = None
        L.orc_synth_batch.argtypes = [C.c_uint64, u32, u32, u32, i32, vp, i32]
        L.orc_synth_batch.restype = None
        L.orc_cloud_scan.argtypes = [vp, sz, C.POINTER(CloudParams), vp]
        L.orc_cloud_scan.restype = u32
        L.orc_dense_decode.argtypes = [vp, u32, u32, C.POINTER(u32), vp, vp, vp]
        L.orc_dense_decode.restype = u32
        L.orc_capsule_bytes.argtypes = [u32]
        L.orc_capsule_bytes.restype = u32
        L.orc_capsule_nodes.argtypes = [u32]
        L.orc_capsule_nodes.restype = u32
        L.orc_crc32_padded.argtypes = [vp, u32]
        L.orc_crc32_padded.restype = u32
        L.orc_decode_capsules.argtypes = [u32, vp, u32, u32, vp, vp, vp, vp]
        L.orc_decode_capsules.restype = u32
        L.orc_decode_normal.argtypes = [vp, u32, vp, vp, C.POINTER(u32)]
        L.orc_decode_normal.restype = u32
        L.orc_sample_delay_us.argtypes = [u32, vp, u32]
        L.orc_sample_delay_us.restype = C.c_uint64
        L.orc_node_timestamps.argtypes = [u32, vp, vp, vp, vp, u32, vp]
        L.orc_node_timestamps.restype = None
        L.orc_normal_timestamps.argtypes = [vp, vp, u32, u32, vp, vp]
        L.orc_normal_timestamps.restype = None
        L.orc_assemble_scans_ts.argtypes = [vp, u32, vp, u32, u32, vp, u32, vp, u32, vp, vp]
        L.orc_assemble_scans_ts.restype = u32
        L.orc_assemble_scans.argtypes = [vp, u32, vp, u32, u32, vp, u32, vp, u32]
        L.orc_assemble_scans.restype = u32
        _lib = L
    return _lib


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.ref_ascend_scan.argtypes = [C.c_void_p, C.c_size_t]
        L.ref_ascend_scan.restype = C.c_uint32
        L.ref_dummy_grab.argtypes = [C.c_void_p, C.c_size_t]
        L.ref_dummy_grab.restype = C.c_int
        L.ref_sizeof_node.restype = C.c_size_t
        L.ref_dense_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t,
                                       C.POINTER(C.c_uint32), C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
        L.ref_dense_decode.restype = C.c_int
        L.ref_unpack.argtypes = [C.c_uint32] + L.ref_dense_decode.argtypes
        L.ref_unpack.restype = C.c_int
        _ref = L
    return _ref


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def make_nodes(angle, dist, quality=None, flag=None) -> np.ndarray:
    angle = np.asarray(angle)
    n = np.zeros(angle.shape, dtype=NODE_DTYPE)
    n["angle_z_q14"] = angle
    n["dist_mm_q2"] = dist
    n["quality"] = 0 if quality is None else quality
    n["flag"] = 0 if flag is None else flag
    return n


def ascend(nodes: np.ndarray, stable: bool = False):
    """Returns (sl_result, ascended copy)."""
    buf = np.ascontiguousarray(nodes).copy()
    rc = lib().orc_ascend_scan(_ptr(buf), buf.shape[0], int(stable))
    return rc, buf


def ref_ascend(nodes: np.ndarray):
    buf = np.ascontiguousarray(nodes).copy()
    rc = ref().ref_ascend_scan(_ptr(buf), buf.shape[0])
    return rc, buf


def ref_dummy_grab() -> np.ndarray:
    buf = np.zeros(360, dtype=NODE_DTYPE)
    n = ref().ref_dummy_grab(_ptr(buf), 360)
    assert n == 360
    return buf


def dummy_scan(call_index: int) -> np.ndarray:
    buf = np.zeros(360, dtype=NODE_DTYPE)
    lib().orc_dummy_scan(call_index, _ptr(buf))
    return buf


def scan_params(is_new_protocol=0, scan_processing=1, inverted=0, apply_ascend=1, range_max=40.0,
                scan_duration=0.1) -> ScanParams:
    return ScanParams(int(is_new_protocol), int(scan_processing), int(inverted), int(apply_ascend),
                      float(range_max), float(scan_duration))


def publish(nodes: np.ndarray, params: ScanParams, stable: bool = False):
    """Returns (header, ranges[beam_count], intensities[beam_count])."""
    nodes = np.ascontiguousarray(nodes)
    n = nodes.shape[0]
    ranges = np.full(max(n, 1), np.nan, dtype=np.float32)
    inten = np.full(max(n, 1), np.nan, dtype=np.float32)
    hdr = ScanHeader()
    lib().orc_publish_scan(_ptr(nodes), n, C.byref(params), int(stable), _ptr(ranges), _ptr(inten),
                           C.byref(hdr))
    m = hdr.beam_count
    return hdr, ranges[:m].copy(), inten[:m].copy()


def pipeline_batch(nodes: np.ndarray, counts: np.ndarray, params: ScanParams, stable=False, threads=1):
    """nodes [n_scans, stride] (modified in place when apply_ascend).  Returns dict."""
    assert nodes.dtype == NODE_DTYPE and nodes.ndim == 2 and nodes.flags.c_contiguous
    n_scans, stride = nodes.shape
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    ranges = np.zeros((n_scans, stride), dtype=np.float32)
    inten = np.zeros((n_scans, stride), dtype=np.float32)
    beams = np.zeros(n_scans, dtype=np.uint32)
    inc = np.zeros(n_scans, dtype=np.float32)
    status = np.zeros(n_scans, dtype=np.uint32)
    secs = lib().orc_pipeline_batch(_ptr(nodes), _ptr(counts), n_scans, stride, C.byref(params),
                                    int(stable), _ptr(ranges), _ptr(inten), _ptr(beams), _ptr(inc),
                                    _ptr(status), int(threads))
    return dict(seconds=secs, ranges=ranges, intensities=inten, beam_counts=beams,
                angle_increment=inc, status=status)


def synth_batch(first_scan_id: int, n_scans: int, n: int, variant: int = 0, stride: int | None = None,
                threads: int = 0) -> np.ndarray:
    stride = n if stride is None else stride
    out = np.zeros((n_scans, stride), dtype=NODE_DTYPE)
    if threads <= 0:
        threads = min(os.cpu_count() or 1, 32)
    lib().orc_synth_batch(first_scan_id, n_scans, n, stride, variant, _ptr(out), threads)
    return out


def cloud_params(range_min=0.15, range_max=40.0, intensity_min=0.0, voxel_size=0.0, sor_k=0,
                 sor_alpha=1.0, is_new_protocol=0) -> CloudParams:
    return CloudParams(float(range_min), float(range_max), float(intensity_min), float(voxel_size),
                       int(sor_k), float(sor_alpha), int(is_new_protocol), (C.c_uint8 * 3)(0, 0, 0))


def cloud(nodes: np.ndarray, params: CloudParams) -> np.ndarray:
    nodes = np.ascontiguousarray(nodes)
    out = np.zeros((max(nodes.shape[0], 1), 4), dtype=np.float32)
    m = lib().orc_cloud_scan(_ptr(nodes), nodes.shape[0], C.byref(params), _ptr(out))
    return out[:m].copy()


# ---- dense-capsule decode ----------------------------------------------
CAPSULE_BYTES = 84
CAPSULE_OK, CAPSULE_SYNC, CAPSULE_EMIT, CAPSULE_DISCARD = 1, 2, 4, 8
CAPSULE_CHECKSUM_ERR, CAPSULE_ENCODER_RESET_ERR, CAPSULE_BAD_FRAME = 16, 32, 64


def make_dense_capsules(start_angle_q6, sync, dist) -> np.ndarray:
    """Builds framed dense capsules: start_angle_q6 [n] (15 bit), sync [n] bool, dist [n, 40] u16.
    Returns uint8 [n, 84] with correct sync nibbles and checksums."""
    start_angle_q6 = np.asarray(start_angle_q6, dtype=np.uint32)
    n = start_angle_q6.shape[0]
    caps = np.zeros((n, CAPSULE_BYTES), np.uint8)
    word = (start_angle_q6 & 0x7FFF) | (np.asarray(sync, dtype=np.uint32) << 15)
    caps[:, 2] = word & 0xFF
    caps[:, 3] = word >> 8
    d = np.asarray(dist, dtype=np.uint16).reshape(n, 40)
    caps[:, 4::2] = d & 0xFF
    caps[:, 5::2] = d >> 8
    chk = np.bitwise_xor.reduce(caps[:, 2:], axis=1)
    caps[:, 0] = 0xA0 | (chk & 0xF)
    caps[:, 1] = 0x50 | (chk >> 4)
    return caps


def dense_decode(capsules: np.ndarray, sample_duration_us: int = 31, sync_state: int = 0):
    """Returns (nodes, capsule_status, capsule_node_offset, sync_state_out)."""
    capsules = np.ascontiguousarray(capsules, dtype=np.uint8).reshape(-1, CAPSULE_BYTES)
    n = capsules.shape[0]
    nodes = np.zeros(max(40 * n, 1), NODE_DTYPE)
    status = np.zeros(max(n, 1), np.uint32)
    offs = np.zeros(max(n, 1), np.uint32)
    st = C.c_uint32(sync_state)
    m = lib().orc_dense_decode(_ptr(capsules), n, sample_duration_us, C.byref(st), _ptr(nodes), _ptr(status), _ptr(offs))
    return nodes[:m].copy(), status[:n].copy(), offs[:n].copy(), st.value


def ref_dense_decode(stream_bytes: np.ndarray, sample_duration_us: int = 31, chunk: int = 84):
    """The SDK's own unpacker on a raw byte stream.  Returns (nodes, events[n,3])."""
    b = np.ascontiguousarray(stream_bytes, dtype=np.uint8).reshape(-1)
    cap_nodes = 40 * (b.shape[0] // CAPSULE_BYTES + 2)
    nodes = np.zeros(cap_nodes, NODE_DTYPE)
    events = np.zeros((b.shape[0] // CAPSULE_BYTES * 2 + 16, 3), np.uint32)
    nn, ne = C.c_uint32(0), C.c_uint32(0)
    rc = ref().ref_dense_decode(_ptr(b), b.shape[0], chunk, sample_duration_us, _ptr(nodes), cap_nodes, C.byref(nn),
                                _ptr(events), events.shape[0], C.byref(ne))
    assert rc == 0, rc
    return nodes[: nn.value].copy(), events[: ne.value].copy()


# ---- the other measurement answer formats (capsule_oracle.cpp) ------------------------------------
ANS_NORMAL, ANS_EXPRESS, ANS_HQ, ANS_ULTRA, ANS_DENSE, ANS_ULTRA_DENSE = 0x81, 0x82, 0x83, 0x84, 0x85, 0x86


def capsule_bytes(ans: int) -> int:
    return int(lib().orc_capsule_bytes(ans))


def capsule_nodes(ans: int) -> int:
    return int(lib().orc_capsule_nodes(ans))


def seal_capsules(ans: int, payload: np.ndarray, start_q6=None, sync=None) -> np.ndarray:
    """Turns [n, capsule_bytes] payload bytes into well-formed capsules: optional start angle / scan-start
    bit, then the sync markers and the checksum (CRC32 for HQ capsules)."""
    cb = capsule_bytes(ans)
    caps = np.ascontiguousarray(payload, dtype=np.uint8).reshape(-1, cb).copy()
    n = caps.shape[0]
    if ans == ANS_HQ:
        caps[:, 0] = 0xA5
        for j in range(n):
            crc = lib().orc_crc32_padded(_ptr(caps[j]), cb - 4)
            caps[j, cb - 4:] = np.frombuffer(np.uint32(crc).tobytes(), np.uint8)
        return caps
    off = 8 if ans == ANS_ULTRA_DENSE else 2
    if start_q6 is not None:
        word = (np.asarray(start_q6, dtype=np.uint32) & 0x7FFF) | (np.asarray(sync, dtype=np.uint32) << 15)
        caps[:, off] = word & 0xFF
        caps[:, off + 1] = word >> 8
    chk = np.bitwise_xor.reduce(caps[:, 2:], axis=1)
    caps[:, 0] = 0xA0 | (chk & 0xF)
    caps[:, 1] = 0x50 | (chk >> 4)
    return caps


def decode_capsules(ans: int, capsules: np.ndarray, sample_duration_us: int = 31, state=(0, 0)):
    """Returns (nodes, capsule_status, capsule_node_offset, state_out)."""
    cb, per = capsule_bytes(ans), capsule_nodes(ans)
    capsules = np.ascontiguousarray(capsules, dtype=np.uint8).reshape(-1, cb)
    n = capsules.shape[0]
    nodes = np.zeros(max(per * n, 1), NODE_DTYPE)
    status = np.zeros(max(n, 1), np.uint32)
    offs = np.zeros(max(n, 1), np.uint32)
    st = np.array(state, np.uint32)
    m = lib().orc_decode_capsules(ans, _ptr(capsules), n, sample_duration_us, _ptr(st), _ptr(nodes), _ptr(status),
                                  _ptr(offs))
    return nodes[:m].copy(), status[:n].copy(), offs[:n].copy(), (int(st[0]), int(st[1]))


def frame_capsules(ans: int, stream_bytes: np.ndarray):
    """The SDK's byte-level framing (sync-nibble hunt) as a function: raw bytes -> (framed capsules with one all-zero
    capsule per skipped stretch, bytes left in an unfinished frame)."""
    b = np.ascontiguousarray(stream_bytes, dtype=np.uint8).reshape(-1)
    cb = capsule_bytes(ans)
    cap = 2 * (b.shape[0] // cb) + 2
    out = np.zeros((cap, cb), np.uint8)
    left = C.c_uint32(0)
    lib().orc_frame_capsules.restype = C.c_uint32
    m = lib().orc_frame_capsules(C.c_uint32(ans), _ptr(b), C.c_uint32(b.shape[0]), _ptr(out), C.c_uint32(cap), C.byref(left))
    assert m <= cap
    return out[:m].copy(), left.value


def decode_normal(stream_bytes: np.ndarray):
    """Returns (nodes, node_end_byte, fsm_pos)."""
    b = np.ascontiguousarray(stream_bytes, dtype=np.uint8).reshape(-1)
    nodes = np.zeros(max(b.shape[0] // 5 + 1, 1), NODE_DTYPE)
    ends = np.zeros(nodes.shape[0], np.uint32)
    pos = C.c_uint32(0)
    m = lib().orc_decode_normal(_ptr(b), b.shape[0], _ptr(nodes), _ptr(ends), C.byref(pos))
    return nodes[:m].copy(), ends[:m].copy(), pos.value


def ref_unpack(ans: int, stream_bytes: np.ndarray, sample_duration_us: int = 31, chunk: int = 0):
    """The SDK's own unpacker on a raw byte stream of any answer type.  Returns (nodes, events[n,3])."""
    b = np.ascontiguousarray(stream_bytes, dtype=np.uint8).reshape(-1)
    cap_nodes = b.shape[0] + 256
    nodes = np.zeros(cap_nodes, NODE_DTYPE)
    events = np.zeros((b.shape[0] // 40 + 16, 3), np.uint32)
    nn, ne = C.c_uint32(0), C.c_uint32(0)
    rc = ref().ref_unpack(ans, _ptr(b), b.shape[0], chunk, sample_duration_us, _ptr(nodes), cap_nodes, C.byref(nn),
                          _ptr(events), events.shape[0], C.byref(ne))
    assert rc == 0, rc
    return nodes[: nn.value].copy(), events[: ne.value].copy()


# ---- scan assembly -------------------------------------------------------
def resets_from_capsules(status: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """Node positions of the scan-reset requests: one per scan-start capsule."""
    return np.ascontiguousarray(offsets[(status & CAPSULE_SYNC) != 0], dtype=np.uint32)


def assemble_scans(nodes: np.ndarray, resets=None, max_nodes: int = 8192, max_scans: int = 64, scan_stride=None):
    """Returns (scans [n_scans_stored, stride], lengths, n_published)."""
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    resets = np.zeros(0, np.uint32) if resets is None else np.ascontiguousarray(resets, dtype=np.uint32)
    stride = max_nodes if scan_stride is None else scan_stride
    out = np.zeros((max_scans, stride), NODE_DTYPE)
    lens = np.zeros(max_scans, np.uint32)
    k = lib().orc_assemble_scans(_ptr(nodes), nodes.shape[0], _ptr(resets), resets.shape[0], max_nodes, _ptr(out),
                                 stride, _ptr(lens), max_scans)
    return out, lens, k


_holder = None


def have_ref_holder() -> bool:
    return os.path.exists(REF_HOLDER_SO)


def ref_assemble_scans(nodes: np.ndarray, resets=None, max_nodes: int = 8192, max_scans: int = 64, scan_stride=None):
    """The reference's own ScanDataHolder on the same stream."""
    global _holder
    if _holder is None:
        _holder = C.CDLL(REF_HOLDER_SO)
        _holder.ref_assemble_scans.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                               C.c_size_t, C.c_void_p, C.c_size_t]
        _holder.ref_assemble_scans.restype = C.c_int
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    resets = np.zeros(0, np.uint32) if resets is None else np.ascontiguousarray(resets, dtype=np.uint32)
    stride = max_nodes if scan_stride is None else scan_stride
    out = np.zeros((max_scans, stride), NODE_DTYPE)
    lens = np.zeros(max_scans, np.uint32)
    k = _holder.ref_assemble_scans(_ptr(nodes), nodes.shape[0], _ptr(resets), resets.shape[0], max_nodes, _ptr(out),
                                   stride, _ptr(lens), max_scans)
    return out, lens, k


# ---- per-sample timestamps ---------------------------------------------------
def timing4(sample_duration_us=31, native_baudrate=0, linkage_delay_us=0, native_interface_type=0) -> np.ndarray:
    return np.array([sample_duration_us, native_baudrate, linkage_delay_us, native_interface_type], np.uint32)


def node_timestamps(ans: int, timing: np.ndarray, capsule_rx_us, status, offsets, n_nodes: int) -> np.ndarray:
    rx = np.ascontiguousarray(capsule_rx_us, dtype=np.uint64)
    status = np.ascontiguousarray(status, dtype=np.uint32)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    ts = np.zeros(max(n_nodes, 1), np.uint64)
    lib().orc_node_timestamps(ans, _ptr(timing), _ptr(rx), _ptr(status), _ptr(offsets), status.shape[0], _ptr(ts))
    return ts[:n_nodes]


def normal_timestamps(timing: np.ndarray, node_end, chunk_bytes: int, chunk_rx_us) -> np.ndarray:
    node_end = np.ascontiguousarray(node_end, dtype=np.uint32)
    rx = np.ascontiguousarray(chunk_rx_us, dtype=np.uint64)
    ts = np.zeros(max(node_end.shape[0], 1), np.uint64)
    lib().orc_normal_timestamps(_ptr(timing), _ptr(node_end), node_end.shape[0], chunk_bytes, _ptr(rx), _ptr(ts))
    return ts[: node_end.shape[0]]


def assemble_scans_ts(nodes, node_ts, resets=None, max_nodes: int = 8192, max_scans: int = 64):
    """Returns (scans, lengths, n_published, scan_begin_ts)."""
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    node_ts = np.ascontiguousarray(node_ts, dtype=np.uint64)
    resets = np.zeros(0, np.uint32) if resets is None else np.ascontiguousarray(resets, dtype=np.uint32)
    out = np.zeros((max_scans, max_nodes), NODE_DTYPE)
    lens = np.zeros(max_scans, np.uint32)
    sts = np.zeros(max_scans, np.uint64)
    k = lib().orc_assemble_scans_ts(_ptr(nodes), nodes.shape[0], _ptr(resets), resets.shape[0], max_nodes, _ptr(out),
                                    max_nodes, _ptr(lens), max_scans, _ptr(node_ts), _ptr(sts))
    return out, lens, k, sts


_clock = None


def have_ref_clock() -> bool:
    return os.path.exists(REF_CLOCK_SO)


def ref_unpack_ts(ans: int, stream_bytes, chunk: int, rx_us, timing: np.ndarray):
    """The SDK's own unpacker, fed `chunk` bytes at a time with its clock set to rx_us[k] before piece k.
    Returns (nodes, timestamps)."""
    global _clock
    if _clock is None:
        _clock = C.CDLL(REF_CLOCK_SO)
        _clock.ref_unpack_ts.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
        _clock.ref_unpack_ts.restype = C.c_int
    b = np.ascontiguousarray(stream_bytes, dtype=np.uint8).reshape(-1)
    rx = np.ascontiguousarray(rx_us, dtype=np.uint64)
    assert rx.shape[0] >= (b.shape[0] + chunk - 1) // chunk
    cap_nodes = b.shape[0] + 256
    nodes = np.zeros(cap_nodes, NODE_DTYPE)
    ts = np.zeros(cap_nodes, np.uint64)
    nn = C.c_uint32(0)
    rc = _clock.ref_unpack_ts(ans, _ptr(b), b.shape[0], chunk, _ptr(rx), _ptr(timing), _ptr(nodes), _ptr(ts), cap_nodes,
                              C.byref(nn))
    assert rc == 0, rc
    return nodes[: nn.value].copy(), ts[: nn.value].copy()


def ref_assemble_scans_ts(nodes, node_ts, resets=None, max_nodes: int = 8192, max_scans: int = 64):
    """The reference's own ScanDataHolder with timestamps: (scans, lengths, n_published, scan_begin_ts)."""
    h = C.CDLL(REF_HOLDER_SO)
    h.ref_assemble_scans_ts.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                        C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    h.ref_assemble_scans_ts.restype = C.c_int
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    node_ts = np.ascontiguousarray(node_ts, dtype=np.uint64)
    resets = np.zeros(0, np.uint32) if resets is None else np.ascontiguousarray(resets, dtype=np.uint32)
    out = np.zeros((max_scans, max_nodes), NODE_DTYPE)
    lens = np.zeros(max_scans, np.uint32)
    sts = np.zeros(max_scans, np.uint64)
    k = h.ref_assemble_scans_ts(_ptr(nodes), nodes.shape[0], _ptr(resets), resets.shape[0], max_nodes, _ptr(out),
                                max_nodes, _ptr(lens), max_scans, _ptr(node_ts), _ptr(sts))
    return out, lens, k, sts


# ---- the reference's real RPlidarNode::publish_scan (compiled against the ROS API stubs) ----------------------
_node = None


def have_ref_node() -> bool:
    return os.path.exists(REF_NODE_SO)


def _node_lib():
    global _node
    if _node is None:
        _node = C.CDLL(REF_NODE_SO)
        _node.ref_publish_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double,
                                           C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_uint32)]
        _node.ref_publish_scan.restype = C.c_int
        _node.ref_pipeline_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_float, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _node.ref_pipeline_batch.restype = C.c_double
    return _node


def ref_pipeline_batch(nodes: np.ndarray, counts: np.ndarray, params: ScanParams, threads=1, outputs=True):
    """The reference's own per-scan path over a batch (grab_scan_data's ascend glue + RPlidarNode::publish_scan,
    compiled from the reference sources): dict like pipeline_batch; with outputs=False only the wall time."""
    assert nodes.dtype == NODE_DTYPE and nodes.ndim == 2 and nodes.flags.c_contiguous
    n_scans, stride = nodes.shape
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    ranges = inten = beams = None
    if outputs:
        ranges = np.zeros((n_scans, stride), np.float32)
        inten = np.zeros((n_scans, stride), np.float32)
        beams = np.zeros(n_scans, np.uint32)
    secs = _node_lib().ref_pipeline_batch(_ptr(nodes), _ptr(counts), n_scans, stride, int(params.is_new_protocol),
                                          int(params.scan_processing), int(params.inverted), int(params.apply_ascend),
                                          float(params.range_max), float(params.scan_duration),
                                          _ptr(ranges) if outputs else None, _ptr(inten) if outputs else None,
                                          _ptr(beams) if outputs else None, int(threads))
    return dict(seconds=secs, ranges=ranges, intensities=inten, beam_counts=beams)


def ref_publish(nodes: np.ndarray, params: ScanParams):
    """RPlidarNode::publish_scan itself (reference src/rplidar_node.cpp:556-680) on one scan.
    Returns (published, header7, ranges, intensities)."""
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    n = nodes.shape[0]
    ranges = np.full(max(n, 1), np.nan, np.float32)
    inten = np.full(max(n, 1), np.nan, np.float32)
    hdr = np.zeros(7, np.float32)
    beams = C.c_uint32(0)
    rc = _node_lib().ref_publish_scan(_ptr(nodes), n, int(params.is_new_protocol), int(params.scan_processing),
                                int(params.inverted), float(params.range_max), float(params.scan_duration),
                                _ptr(ranges), _ptr(inten), ranges.shape[0], _ptr(hdr), C.byref(beams))
    assert rc >= 0, rc
    return rc == 1, hdr, ranges[: beams.value].copy(), inten[: beams.value].copy()
