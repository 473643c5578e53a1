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
        ok = ok and _band_check(rank, world)
        ok = ok and _rows_check(rank, world)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _reference_combine(imgs, alphas):
    """scene_model.py:386-397 restated literally (incl. the in-place alias of the last render)."""
    N, D = imgs.shape[0], imgs.shape[-1]
    allc = [imgs[i].clone() for i in range(N)]
    render_colors = allc[-1]
    avg = torch.stack(allc, 0).mean(0)
    render_colors[:, :, :, 0:D] = avg[:, :, :, 0:D]
    render_colors[:, :, :, 3:4] = torch.stack(allc, 0).max(0)[0][:, :, :, 3:4]
    render_colors[:, :, :, 16:17] = torch.stack(allc, 0).min(0)[0][:, :, :, 16:17]
    return render_colors, torch.stack([alphas[i] for i in range(N)], 0).mean(0)


def _rows_check(rank, world):
    """Band-major partition: rank r holds row band r of ALL sub-exposures, combines locally and all-gathers -- against
    the literal reference combine of the whole image: values, and the gradient every rank's band receives."""
    from deblur4dgs_b200.parallel import band_layout, render_frame_rows
    N, H, W, D = 5, 40, 7, 17
    band_h, n_bands = band_layout(H, world)
    Hp = band_h * n_bands
    g = torch.Generator().manual_seed(2)
    full = torch.randn(N, 1, Hp, W, D, generator=g)
    full[..., 3] = (torch.rand(N, 1, Hp, W, generator=g) > 0.6).float()
    falpha = torch.rand(N, 1, Hp, W, 1, generator=g)
    vi, va = torch.randn(1, H, W, D, generator=g), torch.randn(1, H, W, 1, generator=g)
    leaves = {}

    def render_units(times, RTs, camera_of, row0, bh):
        assert bh == band_h and camera_of.tolist() == list(range(N)) and row0.tolist() == [rank * band_h] * N
        leaves["i"] = full[:, :, rank * band_h:(rank + 1) * band_h].clone().requires_grad_(True)
        leaves["a"] = falpha[:, :, rank * band_h:(rank + 1) * band_h].clone().requires_grad_(True)
        return leaves["i"], leaves["a"]

    out, oa = render_frame_rows(torch.zeros(N), None, H, render_units, combine=_reference_combine)
    ri = full[:, :, :H].clone().requires_grad_(True)
    ra = falpha[:, :, :H].clone().requires_grad_(True)
    ref, refa = _reference_combine(ri, ra)
    ok = out.shape == ref.shape and torch.allclose(out, ref, atol=1e-6) and torch.allclose(oa, refa, atol=1e-6)
    ((out * vi).sum() + (oa * va).sum()).backward()
    ((ref * vi).sum() + (refa * va).sum()).backward()
    r0, r1 = rank * band_h, min((rank + 1) * band_h, H)
    want = torch.zeros(N, 1, band_h, W, D)
    want[:, :, :r1 - r0] = ri.grad[:, :, r0:r1]
    wa = torch.zeros(N, 1, band_h, W, 1)
    wa[:, :, :r1 - r0] = ra.grad[:, :, r0:r1]
    return bool(ok and torch.allclose(leaves["i"].grad, want, atol=1e-6) and torch.allclose(leaves["a"].grad, wa, atol=1e-6))


def _band_check(rank, world):
    """2-D partition: (sub-exposure, row band) units over the ranks, against the literal reference combine --
    values, gradients, exact ties on the max / min channels (mask is 0 / 1 over large areas), rows beyond the image."""
    from deblur4dgs_b200.parallel import band_layout, band_units, combine_band_units
    N, H, W, D = 5, 40, 7, 17
    band_h, n_bands = band_layout(H, world)
    Hp = band_h * n_bands
    g = torch.Generator().manual_seed(1)
    full = torch.randn(N, 1, Hp, W, D, generator=g)
    full[..., 3] = (torch.rand(N, 1, Hp, W, generator=g) > 0.6).float()  # mask channel: exact ties everywhere
    full[:, :, :, :3, 16] = 0.0                                           # depth channel: a tied region
    full[:, :, :10, :, 3] = 0.0                                           # a region where the MEAN wins the max (quirk)
    falpha = torch.rand(N, 1, Hp, W, 1, generator=g)
    vi, va = torch.randn(1, H, W, D, generator=g), torch.randn(1, H, W, 1, generator=g)
    units = band_units(N, rank, world)
    subs, bands = [u[0] for u in units], [u[1] for u in units]
    li = torch.stack([full[s, :, b * band_h:(b + 1) * band_h] for s, b in units]).clone().requires_grad_(True)
    la = torch.stack([falpha[s, :, b * band_h:(b + 1) * band_h] for s, b in units]).clone().requires_grad_(True)
    ok = True
    for quirk in (True, False):
        li.grad = la.grad = None
        out, oa = combine_band_units(li, la, subs, bands, N, n_bands, H, ref_quirk=quirk)
        ri = full[:, :, :H].clone().requires_grad_(True)
        ra = falpha[:, :, :H].clone().requires_grad_(True)
        if quirk:
            ref, refa = _reference_combine(ri, ra)
        else:
            ref = ri.mean(0).clone()
            ref[..., 3] = ri[..., 3].max(0)[0]
            ref[..., 16] = ri[..., 16].min(0)[0]
            refa = ra.mean(0)
        ok = ok and torch.allclose(out, ref, atol=1e-6) and torch.allclose(oa, refa, atol=1e-6)
        ((out * vi).sum() + (oa * va).sum()).backward()
        ((ref * vi).sum() + (refa * va).sum()).backward()
        for k, (s, b) in enumerate(units):
            r0, r1 = b * band_h, min((b + 1) * band_h, H)
            want = torch.zeros(1, band_h, W, D)
            want[:, :r1 - r0] = ri.grad[s, :, r0:r1]
            wa = torch.zeros(1, band_h, W, 1)
            wa[:, :r1 - r0] = ra.grad[s, :, r0:r1]
            ok = ok and torch.allclose(li.grad[k], want, atol=1e-6) and torch.allclose(la.grad[k], wa, atol=1e-6)
    return ok


def test_band_combine_single_process():
    """The same check without a process group (world 1: one band = the whole image)."""
    assert _band_check(0, 1)


def test_sharding_helpers():
    from deblur4dgs_b200.parallel import band_layout, band_units, shard_counts, shard_indices
    assert band_layout(720, 8) == (96, 8) and band_layout(288, 2) == (144, 2)
    units = sum((band_units(9, r, 8) for r in range(8)), [])
    assert len(units) == 72 and len(set(units)) == 72 and all(len(band_units(9, r, 8)) == 9 for r in range(8))
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
