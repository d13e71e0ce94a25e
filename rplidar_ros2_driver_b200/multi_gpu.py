"""Multi-GPU plumbing: stream sharding and the single all-gather of the fused cloud.

One process per GPU (torch.distributed; NCCL over NVLink on the GPUs, gloo in the CPU tests).
Streams are independent, so the LaserScan path needs no collective at all; the PointCloud2 path
has exactly one exchange per step: every rank contributes its dense fused cloud
(`rpl_cloud_fuse_dev`: [points][x, y, z, intensity]) in a fixed-capacity slot plus its point
count, and every rank ends up with the concatenation in rank order.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_streams(n_streams: int, world: int, rank: int) -> range:
    """Contiguous block of stream ids owned by `rank` (blocks differ by at most one stream)."""
    base, extra = divmod(n_streams, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def stream_owner(stream_id: int, n_streams: int, world: int) -> int:
    base, extra = divmod(n_streams, world)
    cut = extra * (base + 1)
    if stream_id < cut:
        return stream_id // (base + 1)
    return extra + (stream_id - cut) // base if base else world - 1


class FusedCloudGather:
    """Preallocated all-gather of per-rank fused clouds.

    capacity: points per rank slot (>= the largest per-rank cloud).  The gathered buffer is
    [world, capacity, 4] float32; `counts` is [world] int32.  One all_gather_into_tensor for the
    payload and one for the 4-byte counts (fused into the payload's tail would save a launch but
    not bandwidth: 4 B against megabytes).
    """

    def __init__(self, capacity: int, device: torch.device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.capacity = int(capacity)
        self.gathered = torch.empty((self.world, self.capacity, 4), dtype=torch.float32, device=device)
        self.counts = torch.zeros(self.world, dtype=torch.int32, device=device)

    def __call__(self, fused: torch.Tensor, total: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """fused: [capacity, 4] float32 (first total[0] points valid); total: [1] int32."""
        assert fused.shape == (self.capacity, 4) and fused.dtype == torch.float32
        if self.world == 1:
            self.gathered[0].copy_(fused)
            self.counts[:1].copy_(total)
        else:
            dist.all_gather_into_tensor(self.gathered.view(-1), fused.reshape(-1), group=self.group)
            dist.all_gather_into_tensor(self.counts, total.to(torch.int32).reshape(1), group=self.group)
        return self.gathered, self.counts

    def compact(self) -> torch.Tensor:
        """The concatenation of all ranks' clouds, in rank order ([sum(counts), 4])."""
        counts: List[int] = self.counts.tolist()
        return torch.cat([self.gathered[r, : counts[r]] for r in range(self.world)], dim=0)

    def payload_bytes(self) -> int:
        return self.world * self.capacity * 16


class _DevMem:
    """A raw device allocation seen through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class PeerCloudGather:
    """fuse + all-gather in ONE kernel over NVLink peer memory (include/rpl_b200.h: rpl_cloud_fuse_push_dev).

    Every rank owns two gather buffers [header | world x capacity x 16 B] allocated by the library
    (cudaMalloc + CUDA IPC handle); the handles are exchanged once with all_gather_object and every rank
    maps all peers' buffers.  push() stores this rank's points straight into its slot of every rank's
    buffer and then runs one tiny all-reduce on the same stream as the barrier that makes "all ranks have
    pushed" visible.  Two buffers alternate so that the push of step k+1 never overwrites what a consumer
    of step k is still reading (a buffer is reused only after the barrier of the step in between).
    """

    HEADER_BYTES = 256

    def __init__(self, ctx, capacity: int, device: torch.device, group=None):
        self.ctx, self.group, self.device = ctx, group, device
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.capacity = int(capacity)
        self.nbytes = self.HEADER_BYTES + self.world * self.capacity * 16
        self.own, self.bases, self._opened = [], [], []
        for _ in range(2):
            ptr, handle = ctx.peer_alloc(self.nbytes)
            handles = [handle]
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, handle, group=group)
            bases = []
            for r in range(self.world):
                if r == self.rank:
                    bases.append(ptr)
                else:
                    bases.append(ctx.peer_open(handles[r]))
                    self._opened.append(bases[-1])
            self.own.append(ptr)
            self.bases.append(bases)
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)
        self._step = 0

    def push(self, xyzi_ptr: int, point_counts_ptr: int, n_scans: int, stride: int, offsets_ptr: int, total_ptr: int,
             stream=None) -> int:
        """Returns the index (0/1) of the buffer that holds this step's gathered cloud."""
        half = self._step & 1
        self._step += 1
        self.ctx.cloud_fuse_push_dev(xyzi_ptr, point_counts_ptr, n_scans, stride, self.bases[half], self.rank,
                                     self.capacity, offsets_ptr, total_ptr, stream=stream)
        if self.world > 1:
            dist.all_reduce(self._flag, group=self.group)  # stream-ordered barrier: every rank has pushed
        return half

    def counts(self, half: int) -> torch.Tensor:
        return torch.as_tensor(_DevMem(self.own[half], (self.world,), "<i4"), device=self.device)

    def gathered(self, half: int) -> torch.Tensor:
        return torch.as_tensor(_DevMem(self.own[half] + self.HEADER_BYTES, (self.world, self.capacity, 4), "<f4"),
                               device=self.device)

    def payload_bytes(self) -> int:
        return self.world * self.capacity * 16

    def close(self):
        if dist.is_initialized() and self.world > 1:
            dist.barrier(group=self.group)  # nobody unmaps while a peer may still push
        for p in self._opened:
            self.ctx.peer_close(p)
        for p in self.own:
            self.ctx.peer_free(p)
        self._opened, self.own = [], []
