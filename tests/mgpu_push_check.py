"""Run under torchrun with >= 2 ranks (one per GPU): the fused pack + all-gather kernel over NVLink peer
memory (rpl_cloud_fuse_push_dev) and the C++ exchange object (rpl_exchange_*: NCCL all-gather and copy-engine
pushes, double-buffered, with a consumer stream) must leave on every rank exactly what rpl_cloud_fuse_dev + one
torch.distributed all-gather leave.  Started by tests/test_gpu_multi_push.py; prints PUSH_OK on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rplidar_ros2_driver_b200 as R  # noqa: E402
from rplidar_ros2_driver_b200.multi_gpu import FusedCloudGather, PeerCloudGather  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    S, N = 96 + 8 * rank, 3200  # ranks contribute different amounts
    ctx = R.Context(local, N, S)
    nodes = torch.empty((S, N, 8), dtype=torch.uint8, device=dev)
    counts = torch.empty(S, dtype=torch.int32, device=dev)
    xyzi = torch.empty((S, N, 4), dtype=torch.float32, device=dev)
    pc = torch.empty(S, dtype=torch.int32, device=dev)
    offs = torch.empty(S, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int32, device=dev)
    cap = S * N + 1024
    cap_t = torch.tensor([cap], device=dev)
    dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
    cap = int(cap_t.item())
    fused = torch.zeros((cap, 4), dtype=torch.float32, device=dev)
    ref_gather = FusedCloudGather(cap, dev)
    peer = PeerCloudGather(ctx, cap, dev)
    prm = R.cloud_params(range_min=0.15, range_max=40.0, voxel_size=0.05)
    ok = True
    for step in range(3):  # three steps: both buffers, and the first one again
        ctx.synth_batch_dev(1000 * rank + 17 * step, S, N, N, 4, nodes.data_ptr(), counts.data_ptr())
        ctx.cloud_batch_dev(nodes.data_ptr(), counts.data_ptr(), S, N, prm, xyzi.data_ptr(), pc.data_ptr())
        ctx.synchronize()
        torch.cuda.synchronize()
        ctx.cloud_fuse_dev(xyzi.data_ptr(), pc.data_ptr(), S, N, fused.data_ptr(), offs.data_ptr(), total.data_ptr(),
                           stream=torch.cuda.current_stream().cuda_stream)
        g_ref, c_ref = ref_gather(fused, total)
        half = peer.push(xyzi.data_ptr(), pc.data_ptr(), S, N, offs.data_ptr(), total.data_ptr(),
                         stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        c_push, g_push = peer.counts(half), peer.gathered(half)
        if not torch.equal(c_push, c_ref):
            ok = False
            print(f"[rank {rank}] step {step}: counts differ {c_push.tolist()} vs {c_ref.tolist()}", flush=True)
        for r in range(world):
            n = int(c_ref[r])
            if not torch.equal(g_push[r, :n].view(torch.int32), g_ref[r, :n].view(torch.int32)):
                ok = False
                print(f"[rank {rank}] step {step}: slot {r} differs", flush=True)
        assert int(c_ref.sum()) > 0
    # ---- the C++ exchange object (rpl_exchange_*): NCCL all-gather and copy-engine pushes, several steps so that
    # both buffers are reused, a consumer on its own stream with wait / release, against the reference gather ----
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(R.exchange_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    torch.cuda.synchronize()
    ex = R.Exchange(ctx, idt.cpu().numpy().tobytes(), world, rank, cap)
    consumer = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream().cuda_stream

    class View:
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}

    for step in range(6):
        mode = R.EXCHANGE_NCCL if step % 3 == 0 else R.EXCHANGE_COPY
        ctx.synth_batch_dev(5000 * rank + 31 * step, S, N, N, 4, nodes.data_ptr(), counts.data_ptr(), stream=main_stream)
        ctx.cloud_batch_dev(nodes.data_ptr(), counts.data_ptr(), S, N, prm, xyzi.data_ptr(), pc.data_ptr(), stream=main_stream)
        idx = ex.allgather(xyzi.data_ptr(), pc.data_ptr(), S, N, mode, stream=main_stream)
        # the reference for this step, from the same per-scan clouds
        ctx.cloud_fuse_dev(xyzi.data_ptr(), pc.data_ptr(), S, N, fused.data_ptr(), offs.data_ptr(), total.data_ptr(),
                           stream=main_stream)
        torch.cuda.synchronize()  # (the exchange's collectives and torch's must not be in flight together)
        ex.synchronize()
        g_ref, c_ref = ref_gather(fused, total)
        torch.cuda.synchronize()
        ex.wait(idx, stream=consumer.cuda_stream)
        with torch.cuda.stream(consumer):
            for r in range(world):
                p_pts, p_cnt = ex.slot(idx, r)
                c = int(torch.as_tensor(View(p_cnt, (1,), "<i4"), device=dev)[0].item())
                pts = torch.as_tensor(View(p_pts, (cap, 4), "<f4"), device=dev)
                if c != int(c_ref[r]) or not torch.equal(pts[:c].view(torch.int32), g_ref[r, :c].view(torch.int32)):
                    ok = False
                    print(f"[rank {rank}] exchange step {step} mode {mode}: slot {r} differs ({c} vs {int(c_ref[r])})", flush=True)
        ex.release(idx, stream=consumer.cuda_stream)
    torch.cuda.synchronize()
    ex.synchronize()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ex.close()
    peer.close()
    ctx.close()
    if rank == 0:
        print("PUSH_OK" if int(flag.item()) == 1 else "PUSH_MISMATCH", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
