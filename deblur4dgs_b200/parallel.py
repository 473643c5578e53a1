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

* ``"bands"`` (strong scaling, balanced for any N: the 2-D partition of SURVEY 8e): every sub-exposure is
  cut into R tile-row bands, the N x R (sub-exposure, band) units are dealt N to a rank in sub-exposure-major
  order, and a rank renders its N units as one "C = N cameras" launch with per-camera row windows
  (``rendering.rasterization(row_windows=...)``: full-image projection, binning / blend on the band only).
  Forward exchange: one SUM all-reduce of the pre-scaled partial image (+ alpha plane), one MAX all-reduce of the
  (mask, -depth) extrema and one MIN all-reduce of the winning sub-exposure index (so that ties -- the mask
  channel is exactly 0 or 1 over large areas -- route their gradient once, to the first sub-exposure, as
  torch.max / min do); backward exchange: the SUM all-reduce of the parameter gradients.  Reproduces the
  reference's in-place alias quirk (``ref_quirk``), i.e. it is training-equivalent to the single-GPU path.

* ``"rows"`` (strong scaling, band-major; what bench.py's strong leg runs): rank r renders the tile-row band r of ALL N
  sub-exposures (one "C = N cameras" launch with a common row window).  Every pixel's N samples then live on ONE
  rank, so the N-way combine (scene_model.py:386-397, alias quirk included) is the single-GPU kernel on the band --
  no image reduction across ranks at all.  Forward exchange: ONE all-gather of the combined band (image | alpha),
  only because the callers expect the whole image on every rank; backward: a slice of the (replicated) image
  cotangent, then the SUM all-reduce of the parameter gradients that every mode has.

Everything here is host logic over ``torch.distributed``; the kernels are unchanged.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


PROFILE = None  # bench.py sets this to a dict: tag -> [(start, end) CUDA events] around every collective


def _all_reduce(t: Tensor, op, group, tag: str):
    prof = PROFILE
    if prof is not None and t.is_cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_reduce(t, op=op, group=group)
        e1.record()
        prof.setdefault(tag, []).append((e0, e1))
    else:
        dist.all_reduce(t, op=op, group=group)


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
        _all_reduce(flat, dist.ReduceOp.SUM, group, "grad_sum")
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


# ------------------------------------------------------------------------------------------------ #
# 2-D partition: (sub-exposure, tile-row band) units
# ------------------------------------------------------------------------------------------------ #
def band_layout(height: int, world: int, tile_size: int = 16):
    """R = world bands of equal height (a multiple of the tile size) that cover the image: (band_rows_px, n_bands)."""
    tile_h = (height + tile_size - 1) // tile_size
    rows = (tile_h + world - 1) // world
    return rows * tile_size, world


def band_units(n_sub: int, rank: int, world: int) -> List[tuple]:
    """The (sub-exposure, band) units of ``rank``: N x R units in sub-exposure-major order, N consecutive ones each."""
    units = [(s, b) for s in range(n_sub) for b in range(world)]
    return units[rank * n_sub:(rank + 1) * n_sub]


def _i32_array(xs):
    import ctypes
    return (ctypes.c_int32 * max(len(xs), 1))(*xs)


class _BandCombine(torch.autograd.Function):
    """N-way combine (scene_model.py:386-397) over units spread across ranks.  imgs [U,1,band_h,W,D], alphas
    [U,1,band_h,W,1] are this rank's units, ``subs`` / ``bands`` their coordinates.  CUDA tensors go through the
    d4_band_* kernels (csrc/band_combine.cu); the torch expressions of the same steps below serve the world-size-2 gloo
    tests of this host logic on CPU tensors (tests/test_parallel_gloo.py) and are the kernels' reference."""

    @staticmethod
    def forward(ctx, imgs, alphas, subs, bands, n_sub, n_bands, height, max_ch, min_ch, ref_quirk, group):
        U, _, bh, W, D = imgs.shape
        dev, dt = imgs.device, imgs.dtype
        Hp = n_bands * bh
        n_ext = n_sub - 1 if ref_quirk else n_sub  # extrema over r_0 .. r_{N-2} (+ the mean) in quirk mode
        distributed = dist.is_initialized() and dist.get_world_size(group) > 1
        if imgs.is_cuda:
            from ._cabi import call, check_tensors, ptr, stream_ptr
            check_tensors(imgs, alphas, what="band combine")
            imgs, alphas = imgs.float().contiguous(), alphas.float().contiguous()
            c_subs, c_bands = _i32_array(subs), _i32_array(bands)
            part = torch.empty((Hp, W, D + 1), dtype=torch.float32, device=dev)
            ext = torch.empty((Hp, W, 2), dtype=torch.float32, device=dev)
            call("d4_band_partial", ptr(imgs), ptr(alphas), c_subs, c_bands, U, n_sub, n_ext, n_bands, bh, W, D, max_ch,
                 min_ch, ptr(part), ptr(ext), stream_ptr())
            if distributed:
                _all_reduce(part, dist.ReduceOp.SUM, group, "image_sum")
                _all_reduce(ext, dist.ReduceOp.MAX, group, "extrema_max")
            winner = torch.empty((Hp, W, 2), dtype=torch.int32, device=dev)
            call("d4_band_winner", ptr(imgs), c_subs, c_bands, U, n_ext, n_bands, bh, W, D, max_ch, min_ch, ptr(ext),
                 ptr(winner), stream_ptr())
            if distributed:
                _all_reduce(winner, dist.ReduceOp.MIN, group, "winner_min")
            out = torch.empty((1, height, W, D), dtype=torch.float32, device=dev)
            out_alpha = torch.empty((1, height, W, 1), dtype=torch.float32, device=dev)
            call("d4_band_finalize", ptr(part), ptr(ext), ptr(winner), height, W, D, max_ch, min_ch, int(ref_quirk), n_ext,
                 ptr(out), ptr(out_alpha), stream_ptr())
            ctx.save_for_backward(winner)
            ctx.cfg = (subs, bands, n_sub, n_bands, bh, height, (max_ch, min_ch), imgs.shape, alphas.shape)
            return out, out_alpha
        part = torch.zeros((Hp, W, D + 1), dtype=dt, device=dev)
        chans = [(0, max_ch, 1.0), (1, min_ch, -1.0)]
        chans = [c for c in chans if 0 <= c[1] < D]
        ext = torch.full((Hp, W, 2), float("-inf"), dtype=dt, device=dev)
        for u in range(U):
            rows = slice(bands[u] * bh, (bands[u] + 1) * bh)
            part[rows, :, :D] += imgs[u, 0] / n_sub
            part[rows, :, D:] += alphas[u, 0] / n_sub
            if subs[u] < n_ext:
                for j, ch, sg in chans:
                    ext[rows, :, j] = torch.maximum(ext[rows, :, j], sg * imgs[u, 0, :, :, ch])
        if distributed:
            _all_reduce(part, dist.ReduceOp.SUM, group, "image_sum")
            _all_reduce(ext, dist.ReduceOp.MAX, group, "extrema_max")
        # first sub-exposure attaining the extremum (torch.max / min(dim) semantics), agreed on globally
        winner = torch.full(ext.shape, 2 ** 30, dtype=torch.int32, device=dev)
        for u in range(U):
            if subs[u] < n_ext:
                rows = slice(bands[u] * bh, (bands[u] + 1) * bh)
                for j, ch, sg in chans:
                    hit = (sg * imgs[u, 0, :, :, ch]) == ext[rows, :, j]
                    winner[rows, :, j] = torch.where(hit, torch.clamp(winner[rows, :, j], max=subs[u]), winner[rows, :, j])
        if distributed:
            _all_reduce(winner, dist.ReduceOp.MIN, group, "winner_min")
        out = part[:height, :, :D].clone()
        out_alpha = part[:height, :, D:].clone()
        for j, ch, sg in chans:
            best = sg * ext[:height, :, j]
            mean = out[:, :, ch]
            if ref_quirk:
                mean_wins = (mean > best) if j == 0 else (mean < best)
                if n_ext == 0:
                    mean_wins = torch.ones_like(mean_wins)
                winner[:height, :, j] = torch.where(mean_wins, torch.full_like(winner[:height, :, j], -1), winner[:height, :, j])
                out[:, :, ch] = torch.where(mean_wins, mean, best)
            else:
                out[:, :, ch] = best
        ctx.save_for_backward(winner)
        ctx.cfg = (subs, bands, n_sub, n_bands, bh, height, (max_ch, min_ch), imgs.shape, alphas.shape)
        return out[None], out_alpha[None]

    @staticmethod
    def backward(ctx, v_out, v_alpha):
        (winner,) = ctx.saved_tensors
        subs, bands, n_sub, n_bands, bh, height, (max_ch, min_ch), ishape, ashape = ctx.cfg
        U, _, _, W, D = ishape
        dev = winner.device
        Hp = winner.shape[0]
        if winner.is_cuda:
            from ._cabi import call, ptr, stream_ptr
            v_out = v_out.float().contiguous() if v_out is not None else torch.zeros((1, height, W, D), device=dev)
            v_alpha = v_alpha.float().contiguous() if v_alpha is not None else torch.zeros((1, height, W, 1), device=dev)
            v_imgs = torch.empty(ishape, dtype=torch.float32, device=dev)
            v_alphas = torch.empty(ashape, dtype=torch.float32, device=dev)
            call("d4_band_bwd", ptr(winner), _i32_array(subs), _i32_array(bands), U, n_sub, n_bands, bh, W, D, height,
                 max_ch, min_ch, ptr(v_out), ptr(v_alpha), ptr(v_imgs), ptr(v_alphas), stream_ptr())
            return v_imgs, v_alphas, None, None, None, None, None, None, None, None, None
        chans = [c for c in ((0, max_ch), (1, min_ch)) if 0 <= c[1] < D]
        vo = torch.zeros((Hp, W, D), dtype=torch.float32, device=dev)
        va = torch.zeros((Hp, W, 1), dtype=torch.float32, device=dev)
        if v_out is not None:
            vo[:height] = v_out[0]
        if v_alpha is not None:
            va[:height] = v_alpha[0]
        v_imgs = torch.empty(ishape, dtype=torch.float32, device=dev)
        v_alphas = torch.empty(ashape, dtype=torch.float32, device=dev)
        for u in range(U):
            rows = slice(bands[u] * bh, (bands[u] + 1) * bh)
            v_imgs[u, 0] = vo[rows] / n_sub
            v_alphas[u, 0] = va[rows] / n_sub
            for j, ch in chans:
                w = winner[rows, :, j]
                routed = torch.where(w == subs[u], vo[rows, :, ch], torch.zeros_like(vo[rows, :, ch]))
                v_imgs[u, 0, :, :, ch] = torch.where(w < 0, vo[rows, :, ch] / n_sub, routed)  # -1: the mean itself won
        return v_imgs, v_alphas, None, None, None, None, None, None, None, None, None


def combine_band_units(imgs: Tensor, alphas: Tensor, subs: Sequence[int], bands: Sequence[int], n_sub: int, n_bands: int,
                       height: int, ref_quirk: bool = True, group=None):
    D = imgs.shape[-1]
    return _BandCombine.apply(imgs, alphas, list(subs), list(bands), int(n_sub), int(n_bands), int(height),
                              3 if D > 3 else -1, 16 if D > 16 else -1, bool(ref_quirk), group)


def render_frame_banded(times: Tensor, RTs: Optional[Tensor], height: int, render_units, ref_quirk: bool = True, group=None):
    """Strong scaling, 2-D partition: this rank renders N (sub-exposure, row band) units of ONE frame.

    ``render_units(times_d [S], RTs_d [S,3,4] | None, camera_of i64 [U], row0 i32 [U], band_h) -> (imgs
    [U,1,band_h,W,D], alphas [U,1,band_h,W,1])`` is the single-GPU path with row windows: S distinct sub-exposures
    are deformed, unit u shows sub-exposure ``camera_of[u]`` of them through the rows ``[row0[u], row0[u] + band_h)``
    (``scene.render_subexposures(..., camera_of=..., row_windows=(row0, band_h), combine=False)``).  Returns the
    combined image [1,H,W,D] and alpha [1,H,W,1], replicated on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    N = times.shape[0]
    band_h, n_bands = band_layout(height, world)
    units = band_units(N, rank, world)
    subs, bands = [u[0] for u in units], [u[1] for u in units]
    distinct = sorted(set(subs))
    idx = torch.as_tensor(distinct, dtype=torch.long, device=times.device)
    camera_of = torch.as_tensor([distinct.index(s) for s in subs], dtype=torch.long, device=times.device)
    row0 = torch.as_tensor([b * band_h for b in bands], dtype=torch.int32, device=times.device)
    imgs, alphas = render_units(times[idx], None if RTs is None else RTs[idx], camera_of, row0, band_h)
    return combine_band_units(imgs, alphas, subs, bands, N, n_bands, height, ref_quirk, group)


# ------------------------------------------------------------------------------------------------ #
# band-major partition: rank r owns row band r of every sub-exposure; the combine is local
# ------------------------------------------------------------------------------------------------ #
class _GatherBands(torch.autograd.Function):
    """band [1, band_h, W, D] of this rank -> the whole image [1, H, W, D] on every rank (all-gather along the rows).
    Backward: the rows of this rank out of the image cotangent, which every rank holds in full (each computes the
    loss on the whole replicated image), so no reduction is needed."""

    @staticmethod
    def forward(ctx, band, height, group):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        _, bh, W, D = band.shape
        band = band.contiguous()
        if world > 1:
            full = torch.empty((world, bh, W, D), dtype=band.dtype, device=band.device)
            prof = PROFILE
            if prof is not None and band.is_cuda:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dist.all_gather_into_tensor(full, band, group=group)
                e1.record()
                prof.setdefault("band_gather", []).append((e0, e1))
            else:
                dist.all_gather_into_tensor(full, band, group=group)
        else:
            full = band
        ctx.cfg = (rank, bh, height)
        return full.reshape(1, world * bh, W, D)[:, :height]

    @staticmethod
    def backward(ctx, v_full):
        rank, bh, height = ctx.cfg
        lo, hi = rank * bh, min((rank + 1) * bh, height)
        v_band = v_full.new_zeros((1, bh) + tuple(v_full.shape[2:]))
        if hi > lo:
            v_band[:, :hi - lo] = v_full[:, lo:hi]
        return v_band, None, None


def render_frame_rows(times: Tensor, RTs: Optional[Tensor], height: int, render_units, ref_quirk: bool = True, group=None,
                      combine=None):
    """Strong scaling, band-major partition: this rank renders rows ``[rank * band_h, (rank + 1) * band_h)`` of all N
    sub-exposures of ONE frame, combines them locally and all-gathers the combined band.

    ``render_units`` as in ``render_frame_banded``.  Returns the combined image [1,H,W,D] and alpha [1,H,W,1],
    replicated on every rank -- the same values as the single-GPU path (the combine sees exactly the N samples of a
    pixel in sub-exposure order, ``ref_quirk`` included).  ``combine(imgs, alphas) -> (img, alpha)`` replaces the
    CUDA combine kernel (the CPU tests of this host logic pass the reference expression)."""
    from .scene import combine_subexposures
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    N = times.shape[0]
    band_h, _ = band_layout(height, world)
    camera_of = torch.arange(N, dtype=torch.long, device=times.device)
    row0 = torch.full((N,), rank * band_h, dtype=torch.int32, device=times.device)
    imgs, alphas = render_units(times, RTs, camera_of, row0, band_h)  # [N,1,band_h,W,D], [N,1,band_h,W,1]
    D = imgs.shape[-1]
    if combine is not None:
        img_b, acc_b = combine(imgs, alphas)
    else:
        img_b, acc_b = combine_subexposures(imgs, alphas, 3 if D > 3 else -1, 16 if D > 16 else -1, ref_quirk=ref_quirk)
    both = _GatherBands.apply(torch.cat([img_b, acc_b], dim=-1), int(height), group)  # one collective for image | alpha
    return both[..., :D], both[..., D:]
