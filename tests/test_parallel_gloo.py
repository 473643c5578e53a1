"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in deblur4dgs_b200/parallel.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deblur4dgs_b200.parallel import _DistCombine, allreduce_sum_, shard_indices
        N, H, W, D = 5, 6, 7, 17
        g = torch.Generator().manual_seed(0)
        imgs = torch.randn(N, 1, H, W, D, generator=g)
        alphas = torch.rand(N, 1, H, W, 1, generator=g)
        vi, va = torch.randn(1, H, W, D, generator=g), torch.randn(1, H, W, 1, generator=g)
        mine = shard_indices(N, rank, world)
        li = imgs[mine].clone().requires_grad_(True)
        la = alphas[mine].clone().requires_grad_(True)
        out, oa = _DistCombine.apply(li, la, N, 3, 16, None)
        # single-process reference: mean, max on ch 3, min on ch 16 over all N
        ri = imgs.clone().requires_grad_(True)
        ra = alphas.clone().requires_grad_(True)
        ref = ri.mean(0).clone()
        ref[..., 3] = ri[..., 3].max(0)[0]
        ref[..., 16] = ri[..., 16].min(0)[0]
        ok = torch.allclose(out, ref, atol=1e-6) and torch.allclose(oa, ra.mean(0), atol=1e-6)
        ((out * vi).sum() + (oa * va).sum()).backward()
        ((ref * vi).sum() + (ra.mean(0) * va).sum()).backward()
        ok = ok and torch.allclose(li.grad, ri.grad[mine], atol=1e-6) and torch.allclose(la.grad, ra.grad[mine], atol=1e-6)
        # bucketed gradient all-reduce
        ts = [torch.full((3, 4), float(rank + 1)), None, torch.arange(5.0) * (rank + 1), torch.ones(1000) * rank]
        allreduce_sum_(ts, bucket_bytes=64)
        tot = sum(r + 1 for r in range(world))
        ok = ok and torch.equal(ts[0], torch.full((3, 4), float(tot))) and torch.equal(ts[2], torch.arange(5.0) * tot)
        ok = ok and torch.equal(ts[3], torch.ones(1000) * sum(range(world)))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharding_helpers():
    from deblur4dgs_b200.parallel import shard_counts, shard_indices
    assert shard_counts(9, 8) == [2, 1, 1, 1, 1, 1, 1, 1] and shard_counts(13, 8)[:5] == [2, 2, 2, 2, 2]
    assert sorted(sum((shard_indices(9, r, 4) for r in range(4)), [])) == list(range(9))


@pytest.mark.timeout(120)
def test_dist_combine_and_grad_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]
