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
