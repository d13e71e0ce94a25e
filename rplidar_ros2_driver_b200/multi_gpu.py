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
