"""Data-parallel plumbing for the backbone forward.

The path shards by image: rank r of R runs the whole backbone on images
[r*B/R, (r+1)*B/R) with a full weight replica; there is no cross-image operation anywhere
on the path (LayerNorm is per token, attention per image), hence no data-path collective.
The only exchange is the OUTPUT gather.  Gathering the full pyramids would move 6.4 MB per
image (33 MB at 512x512) into every GPU and cap scaling, so the gathered payload is the
coarsest tap (p5) -- p2..p4 stay rank-local, where a data-parallel detection head consumes
them.  One process per GPU (torchrun); NCCL over NVLink on the GPUs, gloo in the CPU tests.
The reference has no distributed code at all (R:README.md:59): this module is new.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def rank_slice(global_batch: int, rank: int, world: int) -> slice:
    """Images of rank ``rank``: contiguous, sizes differ by at most one (first ranks get the extra)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if global_batch < 0:
        raise ValueError("negative batch")
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def shard_pages(pages: torch.Tensor, rank: int | None = None, world: int | None = None) -> torch.Tensor:
    """This rank's images of a [B, 3, H, W] page batch (a view, no copy)."""
    if rank is None or world is None:
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    return pages[rank_slice(pages.shape[0], rank, world)]


def gather_tap(tap: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather one tap along the batch dimension: [b, D, h, w] per rank -> [B, D, h, w] on every rank.

    Ranks may hold different numbers of images (``rank_slice``).  The tap is exchanged in the
    channels-last memory order the kernels wrote it in, so no transposition happens on either side."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tap
    world = dist.get_world_size(group)
    nhwc = tap.permute(0, 2, 3, 1).contiguous()          # no copy for the library's channels-last outputs
    counts = torch.tensor([nhwc.shape[0]], dtype=torch.int64, device=tap.device)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    sizes = [int(c.item()) for c in all_counts]
    if len(set(sizes)) == 1:                              # the common case: one fused collective
        out = torch.empty((world * sizes[0],) + nhwc.shape[1:], dtype=nhwc.dtype, device=nhwc.device)
        dist.all_gather_into_tensor(out, nhwc, group=group)
    else:                                                 # ragged split: pad to the largest shard, trim after
        m = max(sizes)
        padded = torch.zeros((m,) + nhwc.shape[1:], dtype=nhwc.dtype, device=nhwc.device)
        padded[: nhwc.shape[0]] = nhwc
        buf = torch.empty((world * m,) + nhwc.shape[1:], dtype=nhwc.dtype, device=nhwc.device)
        dist.all_gather_into_tensor(buf, padded, group=group)
        out = torch.cat([buf[r * m: r * m + n] for r, n in enumerate(sizes)], dim=0)
    return out.permute(0, 3, 1, 2)


class PeerGather:
    """All-gather of one fixed-shape tap by PEER COPIES over NVLink instead of a NCCL kernel.

    Every rank stages its shard in a symmetric-memory buffer (``torch.distributed._symmetric_memory``: the same
    allocation mapped into every rank of the node), the ranks meet in a signal-pad barrier (one tiny kernel), each rank
    PULLS the other shards with plain device-to-device copies -- executed by the copy engines, no SM is taken -- and a
    second barrier releases the staging buffers.  This matters because the forward is a sequence of persistent
    kernels that own all 148 SMs: a NCCL kernel running beside them on a side stream displaces some of their CTAs and a
    persistent kernel that loses an SM runs its static tile schedule in two waves (measured: the overlapped NCCL gather
    cost +14 % per step on 8 GPUs where the same gather on the compute stream cost +2.5 %).

    One node only (NVLink / NVSwitch peers), equal shard shapes on every rank.  Construction is collective; it raises if
    symmetric memory is unavailable and the caller falls back to :func:`gather_tap`."""

    def __init__(self, shard_shape, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group or dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.shape, self.dtype = tuple(shard_shape), dtype
        self.stage = symm_mem.empty(self.shape, dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.stage, group.group_name)
        self.out = torch.empty((self.world * self.shape[0],) + self.shape[1:], dtype=dtype, device=device)
        self.peers = [self.handle.get_buffer(r, self.shape, dtype) for r in range(self.world)]

    def stage_in(self, shard: torch.Tensor) -> None:
        """Copy this rank's [b, ...] block into its symmetric staging buffer (current stream)."""
        self.stage.copy_(shard)

    def exchange(self) -> torch.Tensor:
        """Barrier, pull every rank's staged block, barrier (current stream); returns [world*b, ...], valid until the
        next call."""
        self.handle.barrier()                        # every rank's shard is staged
        chunks = self.out.chunk(self.world)
        for step in range(self.world):
            r = (self.rank - step) % self.world      # start with the local shard, spread the pulls over the peers
            chunks[r].copy_(self.peers[r])
        self.handle.barrier()                        # every rank has read every staging buffer: they may be rewritten
        return self.out

    def __call__(self, shard: torch.Tensor) -> torch.Tensor:
        self.stage_in(shard)
        return self.exchange()
