import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rplidar_ros2_driver_b200 as R
from oracle import pyoracle as O
ctx = R.Context(0, 8192, 1)
L = R.lib()
for n in (360, 3200, 8192):
    nodes = O.synth_batch(1, 1, n, 0)[0].view(R.NODE_DTYPE).copy()
    ranges = np.zeros(n, np.float32); inten = np.zeros(n, np.float32)
    beams, inc, st = C.c_uint32(0), C.c_float(0), C.c_uint32(0)
    def call(fn, reps=300):
        for _ in range(20): fn()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        return (time.perf_counter() - t0) / reps * 1e6
    res = {}
    for name, prm in (("scan+ascend B", R.scan_params(0, 0, 0, 1)), ("scan noascend B", R.scan_params(0, 0, 0, 0)),
                      ("scan+ascend A", R.scan_params(0, 1, 0, 1)), ("scan+ascend B general", R.scan_params(0, 0, 0, 1, 1))):
        buf = nodes.copy()
        f = lambda: L.rpl_scan(ctx._h, C.c_void_p(buf.ctypes.data), n, C.byref(prm), C.c_void_p(ranges.ctypes.data),
                               C.c_void_p(inten.ctypes.data), C.byref(beams), C.byref(inc), C.byref(st))
        res[name] = round(call(f), 1)
    buf = nodes.copy()
    res["ascend only"] = round(call(lambda: L.rpl_ascend_scan(ctx._h, C.c_void_p(buf.ctypes.data), n)), 1)
    prm = R.scan_params(0, 0, 0, 0)
    res["laserscan only"] = round(call(lambda: L.rpl_laserscan(ctx._h, C.c_void_p(nodes.ctypes.data), n, C.byref(prm), C.c_void_p(ranges.ctypes.data), C.c_void_p(inten.ctypes.data), C.byref(beams), C.byref(inc))), 1)
    print(n, res)
# CPU reference for comparison (1 thread)
for n in (360, 3200, 8192):
    nodes = O.synth_batch(1, 64, n, 0)
    cnt = np.full(64, n, np.uint32)
    r = O.pipeline_batch(nodes.copy(), cnt, O.scan_params(0, 0, 0, 1, 40.0, 0.1), threads=1)
    print("cpu oracle 1 thread us/scan", n, round(r["seconds"] / 64 * 1e6, 1))
