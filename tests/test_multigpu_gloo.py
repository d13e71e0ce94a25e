"""World-size-2 gloo tests (CPU) of the N>1 host logic: stream sharding and the single
all-gather of the fused cloud.  The GPU run uses the same code with the nccl backend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_streams_partitions_every_stream_once():
    from rplidar_ros2_driver_b200.multi_gpu import shard_streams, stream_owner

    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                blk = shard_streams(n, world, r)
                seen += list(blk)
                for s in blk:
                    assert stream_owner(s, n, world) == r
            assert seen == list(range(n))
            sizes = [len(shard_streams(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    from rplidar_ros2_driver_b200.multi_gpu import FusedCloudGather, shard_streams

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cap = 1000
    mine = list(shard_streams(10, world, rank))
    # a deterministic "cloud" per stream: stream id in x, point index in y
    pts = [torch.tensor([[float(s), float(i), 0.0, float(s * 7 % 5)] for i in range(50 + 13 * s)]) for s in mine]
    cloud = torch.cat(pts)
    fused = torch.zeros((cap, 4))
    fused[: cloud.shape[0]] = cloud
    g = FusedCloudGather(cap, torch.device("cpu"))
    gathered, counts = g(fused, torch.tensor([cloud.shape[0]], dtype=torch.int32))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), g.compact().numpy())
    np.save(os.path.join(out_dir, f"counts{rank}.npy"), counts.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_fused_cloud_allgather_world2_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    expect = np.concatenate([[[float(s), float(i), 0.0, float(s * 7 % 5)] for i in range(50 + 13 * s)]
                             for s in range(10)]).astype(np.float32)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert got.shape == expect.shape and (got == expect).all()
        counts = np.load(tmp_path / f"counts{r}.npy")
        assert counts.sum() == expect.shape[0] and len(counts) == world
