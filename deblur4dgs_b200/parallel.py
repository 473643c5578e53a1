"""Multi-GPU sharding of the render path.  One process per GPU,
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) for the exchange steps.

The path shards two ways, both with replicated parameters (~20 MB at config c3):

* ``"frames"``  (weak scaling, the default of bench.py): every rank renders its own
  blurry frame (its own camera / timestamps, all N sub-exposures).  No data-path
  collective in the forward; the only exchange is the SUM all-reduce of the parameter
  gradients, as in data-parallel training over ``batch1/2/3`` (trainer.py:211-222).
* ``"subexposures"`` (strong scaling, BASELINE.json configs[3]): the N sub-exposures of
  ONE frame are dealt round-robin to the ranks (``ii -> rank ii % R``,
  scene_model.py:323 loop index).  Forward exchange: SUM all-reduce of the locally
  pre-averaged image (+ MAX on the mask channel, MIN on the depth channel), i.e. the
  N-way combine of scene_model.py:386-397 distributed; backward exchange: SUM all-reduce
  of the parameter gradients.

Everything here is host logic over ``torch.distributed``; the kernels are unchanged.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: item ii belongs to rank ii % world."""
    return [i for i in range(n_items) if i % world == rank]


def shard_counts(n_items: int, world: int) -> List[int]:
    return [len(shard_indices(n_items, r, world)) for r in range(world)]


def allreduce_sum_(tensors: Iterable[Optional[Tensor]], group=None, bucket_bytes: int = 64 << 20) -> None:
    """In-place SUM all-reduce of a list of tensors, coalesced into flat buckets so that the
    number of collectives is bounded by launch latency, not by the parameter count."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    todo = [t for t in tensors if t is not None]
    bucket: List[Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([t.reshape(-1) for t in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for t in bucket:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
        bucket, size = [], 0

    for t in todo:
        nbytes = t.numel() * t.element_size()
        if size + nbytes > bucket_bytes and bucket:
            flush()
        bucket.append(t)
        size += nbytes
    flush()


class _DistCombine(torch.autograd.Function):
    """Distributed N-way combine: local (sum, max, min) over the rank's sub-exposures, then
    all-reduce.  Backward routes the gradient exactly as the single-GPU combine does with
    ``ref_quirk=False`` (extrema over all N)."""

    @staticmethod
    def forward(ctx, imgs, alphas, n_total, max_ch, min_ch, group):
        # imgs [n_local, ..., D] (n_local may be 0 on some ranks when N < world)
        D = imgs.shape[-1]
        ssum = imgs.sum(0) if imgs.shape[0] else torch.zeros(imgs.shape[1:], dtype=imgs.dtype, device=imgs.device)
        asum = alphas.sum(0) if alphas.shape[0] else torch.zeros(alphas.shape[1:], dtype=alphas.dtype, device=alphas.device)
        dist.all_reduce(ssum, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(asum, op=dist.ReduceOp.SUM, group=group)
        out = ssum / n_total
        gmax = gmin = None
        if 0 <= max_ch < D:
            lmax = imgs[..., max_ch].max(0)[0] if imgs.shape[0] else torch.full(imgs.shape[1:-1], float("-inf"), device=imgs.device)
            gmax = lmax.clone()
            dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
            out[..., max_ch] = gmax
        if 0 <= min_ch < D:
            lmin = imgs[..., min_ch].min(0)[0] if imgs.shape[0] else torch.full(imgs.shape[1:-1], float("inf"), device=imgs.device)
            gmin = lmin.clone()
            dist.all_reduce(gmin, op=dist.ReduceOp.MIN, group=group)
            out[..., min_ch] = gmin
        ctx.save_for_backward(imgs, gmax, gmin)
        ctx.cfg = (n_total, max_ch, min_ch, alphas.shape)
        return out, asum / n_total

    @staticmethod
    def backward(ctx, v_out, v_alpha):
        imgs, gmax, gmin = ctx.saved_tensors
        n_total, max_ch, min_ch, ashape = ctx.cfg
        v_imgs = (v_out / n_total).unsqueeze(0).expand(imgs.shape).clone()
        for ch, ext in ((max_ch, gmax), (min_ch, gmin)):
            if ext is None:
                continue
            hit = imgs[..., ch] == ext.unsqueeze(0)
            # first arg-extremum within the rank; ties across ranks are measure-zero for float renders
            first = hit & (torch.cumsum(hit.int(), 0) == 1)
            v_imgs[..., ch] = torch.where(first, v_out[..., ch].unsqueeze(0), torch.zeros_like(v_imgs[..., ch]))
        v_alphas = (v_alpha / n_total).unsqueeze(0).expand(ashape).clone()
        return v_imgs, v_alphas, None, None, None, None


def render_frame_sharded(scene_args: Dict, times: Tensor, RTs: Optional[Tensor], render_local, group=None):
    """Strong-scaling mode: this rank renders sub-exposures ``ii % world == rank`` of one frame.

    ``render_local(times_local, RTs_local) -> (imgs [n_local,1,H,W,D], alphas [n_local,1,H,W,1])`` is the
    single-GPU path (scene.render_subexposures with combine=False).  Returns the combined image and
    alpha (replicated on every rank) -- the distributed equivalent of scene_model.py:386-397."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    N = times.shape[0]
    mine = shard_indices(N, rank, world)
    idx = torch.as_tensor(mine, dtype=torch.long, device=times.device)
    imgs, alphas = render_local(times[idx], None if RTs is None else RTs[idx])
    D = imgs.shape[-1]
    if world == 1:
        from .scene import combine_subexposures
        return combine_subexposures(imgs, alphas, 3 if D > 3 else -1, 16 if D > 16 else -1, ref_quirk=False)
    return _DistCombine.apply(imgs, alphas, N, 3 if D > 3 else -1, 16 if D > 16 else -1, group)
