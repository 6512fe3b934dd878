"""Host logic of the data-parallel path on CPU: world_size-2 gloo processes (no GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from layoutdit_b200.sharding import gather_tap, rank_slice, shard_pages


def test_rank_slices_partition_the_batch():
    for B in (0, 1, 7, 32, 64, 65):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s = rank_slice(B, r, world)
                seen += list(range(B))[s]
            assert seen == list(range(B))
            sizes = [rank_slice(B, r, world).stop - rank_slice(B, r, world).start for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        rank_slice(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        pages = torch.rand(B, 3, 32, 32, generator=g)
        mine = shard_pages(pages)
        assert mine.data_ptr() == pages[rank_slice(B, rank, world)].data_ptr()
        # stand-in for the per-rank backbone: a per-image function, stored channels-last like the library's taps
        tap = mine.mean(dim=1, keepdim=True).repeat(1, 8, 1, 1)[:, :, ::2, ::2].contiguous(memory_format=torch.channels_last)
        full = gather_tap(tap)
        ref = pages.mean(dim=1, keepdim=True).repeat(1, 8, 1, 1)[:, :, ::2, ::2]
        ok = full.shape == ref.shape and torch.equal(full, ref)
        out.put((rank, bool(ok), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5])
def test_gather_over_two_gloo_ranks(B):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True], res
    assert res[0][2] == (B, 8, 16, 16)
