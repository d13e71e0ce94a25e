import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import rplidar_ros2_driver_b200 as R
S, N = 4096, 32768
ctx = R.Context(0, N, S)
dev = torch.device("cuda")
nodes = torch.empty((S, N, 8), dtype=torch.uint8, device=dev)
counts = torch.empty(S, dtype=torch.int32, device=dev)
ranges = torch.empty((S, N), dtype=torch.float32, device=dev)
intens = torch.empty((S, N), dtype=torch.float32, device=dev)
beams = torch.empty(S, dtype=torch.int32, device=dev)
path = torch.empty(S, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx.synth_batch_dev(0, S, N, N, 0, nodes.data_ptr(), counts.data_ptr(), stream=st.cuda_stream)
torch.cuda.synchronize()
bad = 0
for flags in [0] * 30 + [2]:
    ctx.scan_batch_dev(nodes.data_ptr(), counts.data_ptr(), S, N, R.scan_params(0, 0, 0, 1, flags), ranges=ranges.data_ptr(),
                       intensities=intens.data_ptr(), beam_counts=beams.data_ptr(), path=path.data_ptr(), stream=st.cuda_stream)
    torch.cuda.synchronize()
    idx = torch.nonzero(path).flatten().tolist()
    bad += len(idx)
print("total general-path scans over 31 runs:", bad, "grid info launches", ctx.launch_count)
# timing per kernel variant
for flags in (0, 2):
    ctx.profile(True)
    for _ in range(20):
        ctx.scan_batch_dev(nodes.data_ptr(), counts.data_ptr(), S, N, R.scan_params(0, 0, 0, 1, flags), ranges=ranges.data_ptr(),
                           intensities=intens.data_ptr(), beam_counts=beams.data_ptr(), path=path.data_ptr(), stream=st.cuda_stream)
    torch.cuda.synchronize()
    fm, fn, gm, gn = ctx.profile_read()
    print("flags", flags, "fast ms", fm / fn, "general ms", gm / gn)
