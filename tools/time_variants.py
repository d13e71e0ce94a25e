import sys, os, glob, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r)
import torch
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
ref = {}
for mode_a in (0, 1):
  for flags in (2, 0):
    for rep in range(2):
        ctx.profile(True)
        for _ in range(20):
            ctx.scan_batch_dev(nodes.data_ptr(), counts.data_ptr(), S, N, R.scan_params(0, mode_a, 0, 1, flags), ranges=ranges.data_ptr(),
                               intensities=intens.data_ptr(), beam_counts=beams.data_ptr(), path=path.data_ptr(), stream=st.cuda_stream)
        torch.cuda.synchronize()
        fm, fn, gm, gn = ctx.profile_read()
    live = torch.arange(N, device=dev)[None, :] < beams[:, None]
    snap = (beams.clone(), torch.where(live, ranges, torch.zeros_like(ranges)).view(torch.int32).clone(), torch.where(live, intens, torch.zeros_like(intens)).view(torch.int32).clone())
    if flags == 2: ref[mode_a] = snap
    else:
        ok = all(torch.equal(a, b) for a, b in zip(ref[mode_a], snap))
        print("   TMA == v1 outputs:", ok)
    print("   mode", "A" if mode_a else "B", "flags", flags, "kernel ms %%.4f  -> %%.0f GB/s  general-path scans %%d" %% (fm / fn, 16 * S * N / (fm / fn * 1e-3) / 1e9, int((path != 0).sum())))
''' % root
for lib in sorted(glob.glob(os.path.join(root, "tools/libs/*.so"))) + [""]:
    env = dict(os.environ)
    if lib: env["RPL_B200_LIB"] = lib
    print(lib or "default build", flush=True)
    subprocess.run([sys.executable, "-c", code], env=env)
