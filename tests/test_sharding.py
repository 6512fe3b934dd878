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


# ------------------------------------------------------------ bucketed gradient all-reduce (BASELINE config 5)
def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from layoutdit_b200.train import GradientBuckets
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
        frozen = model[3].bias
        frozen.requires_grad_(False)
        buckets = GradientBuckets(model.parameters(), bucket_bytes=1024)     # forces several buckets
        assert len(buckets.buckets) > 2
        # every rank sees its own shard of the batch
        g = torch.Generator().manual_seed(100)
        xs, ys = torch.randn(8, 16, generator=g), torch.randn(8, 4, generator=g)
        sl = rank_slice(8, rank, world)
        loss = ((model(xs[sl]) - ys[sl]) ** 2).sum() / 8              # global mean: per-rank partial sums
        loss.backward()
        model[2].weight.grad = None                                    # a parameter that got no gradient on this rank
        buckets.all_reduce()
        out[rank] = {n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad}
    finally:
        dist.destroy_process_group()


def test_bucketed_gradient_all_reduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    # reference: the same model on the full batch in one process; DDP averages, so compare against grad / world... of the
    # SUM of the per-rank losses, i.e. the mean of the per-rank gradients
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
    g = torch.Generator().manual_seed(100)
    xs, ys = torch.randn(8, 16, generator=g), torch.randn(8, 4, generator=g)
    per_rank = []
    for r in range(world):
        model.zero_grad()
        sl = rank_slice(8, r, world)
        (((model(xs[sl]) - ys[sl]) ** 2).sum() / 8).backward()
        per_rank.append({n: p.grad.clone() for n, p in model.named_parameters()})
    for n in out[0]:
        want = sum((torch.zeros_like(pr[n]) if n == "2.weight" else pr[n]) for pr in per_rank) / world
        for r in range(world):
            torch.testing.assert_close(out[r][n], want, rtol=1e-6, atol=1e-7)
    assert "3.bias" not in out[0]


def _peer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from layoutdit_b200.sharding import PeerGather
        g = torch.Generator().manual_seed(11)
        full = torch.randn(world * 3, 7, 7, 64, generator=g).to(torch.bfloat16)
        mine = full[rank * 3: (rank + 1) * 3].to(dev)
        pg = PeerGather(tuple(mine.shape), mine.dtype, dev)
        ok = True
        for it in range(3):                                  # repeated use: the closing barrier frees the staging buffers
            got = pg(mine * (it + 1))
            torch.cuda.synchronize(dev)
            ok &= torch.equal(got.cpu(), (full.to(dev) * (it + 1)).cpu())
        ref = gather_tap(mine.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)     # the NCCL path agrees
        ok &= torch.equal(pg(mine).cpu(), ref.cpu())
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_gather_matches_nccl_on_two_gpus():
    """sharding.PeerGather (copy-engine pulls out of symmetric memory) against the NCCL all-gather; needs 2 GPUs."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _ddp_overlap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from layoutdit_b200.train import GradientBuckets
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
        buckets = GradientBuckets(model.parameters(), bucket_bytes=1024, overlap=True)   # hooks launch the buckets during backward
        g = torch.Generator().manual_seed(100)
        xs, ys = torch.randn(8, 16, generator=g), torch.randn(8, 4, generator=g)
        sl = rank_slice(8, rank, world)
        for step in range(2):                                          # the second step checks that the counters re-arm
            model.zero_grad(set_to_none=True)
            (((model(xs[sl]) - ys[sl]) ** 2).sum() / 8).backward()
            launched = sum(w is not None for w in buckets._works)
            buckets.all_reduce()
        out[rank] = ({n: p.grad.clone() for n, p in model.named_parameters()}, launched, len(buckets.buckets))
    finally:
        dist.destroy_process_group()


def test_bucketed_gradient_all_reduce_overlapped_with_backward_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ddp_overlap_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
    g = torch.Generator().manual_seed(100)
    xs, ys = torch.randn(8, 16, generator=g), torch.randn(8, 4, generator=g)
    (((model(xs) - ys) ** 2).sum() / 8 / world).backward()             # mean over ranks of the per-rank gradients
    for r in range(world):
        grads, launched, nb = out[r]
        assert launched == nb                                          # every bucket went out from a hook, before all_reduce()
        for n, p in model.named_parameters():
            torch.testing.assert_close(grads[n], p.grad, rtol=1e-5, atol=1e-6)
