"""gsplat-1.1.1-compatible operator surface on top of libd4gs.so.

``rasterization(...)`` is the drop-in for the call at
``flow3d/scene_model.py:360-373`` (``from gsplat.rendering import
rasterization``, scene_model.py:5): same keyword names, same returns
``(render_colors [C,H,W,D], render_alphas [C,H,W,1], meta)``, same autograd
contract (two autograd Functions with ``meta["means2d"]`` the non-leaf tensor in
between, so ``meta["means2d"].retain_grad()`` -- scene_model.py:456-459 -- yields
the screen-space gradient the densifier reads at trainer.py:975).

The lower-level functions mirror gsplat's own op names
(``fully_fused_projection``, ``isect_tiles``, ``isect_offset_encode``,
``rasterize_to_pixels``).  Everything runs on the current CUDA stream through
the C ABI; there is no CPU / eager fallback.

Extension used by the fused sub-exposure path: ``means`` / ``quats`` may be
``[C,G,3]`` / ``[C,G,4]`` (one deformed copy of the scene per "camera" = per
sub-exposure) while ``viewmats`` / ``Ks`` hold a single camera.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _cabi
from ._cabi import call, ptr, stream_ptr

SUPPORTED_D = (1, 2, 3, 4, 5, 6, 7, 8, 9, 16, 17, 32, 33)
CHANNEL_CHUNK = 32
HIT_MASKS = True  # forward records which 8x4 blocks pass the alpha test per intersection; the backward skips the rest
HIT_MASK_TAP = None  # tests set this to a list: every forward appends its hit-mask tensor (parity check against the oracle)
BINNING_METHOD = "auto"  # "auto" | "bucket" | "radix" (see bin_tiles)
# "slab": packed per-tile record slabs streamed through an mbarrier ring by a producer warp (csrc/slab.cuh; default);
# "direct": the gather-staged kernels of csrc/blend.cu / blend_bwd_gp.cu (kept as the second implementation the
# parity tests cross-check)
BLEND_PATH = "slab"
BWD_MODE = 0  # direct path only: 0 = grouped backward kernel, 1 = warp-shuffle backward kernel
SLAB_D0 = (4, 8, 16, 32)  # colour widths the slab kernels are built for (rasterization() pads)
# slab backward formulation: None = the library default, 0 = fp32 pipe, 1 / 2 = tensor cores (d4_blend_bwd_slab_variant)
SLAB_BWD_VARIANT = None
# slab forward formulation: None = the library default, 0 = fp32 pipe, 1 = queued + tensor cores (d4_blend_fwd_slab_variant)
SLAB_FWD_VARIANT = None


class RenderCapacity:
    """Buffer capacities for the sync-free ("capacity") mode of the tile binning.

    gsplat sizes its intersection buffers by reading the intersection count back to the host inside every call
    (``isect_tiles``); that read-back is the one device -> host sync of ``rasterization`` and it keeps the step from
    being captured in a CUDA graph.  With a ``RenderCapacity`` the buffers are sized by a capacity instead
    (``headroom`` x the largest count seen so far), the count stays on the device, and an overflow only raises a
    device-side flag.  Counts and flag are copied to pinned host memory asynchronously after every binning;
    ``poll()`` -- called at the start of the next render, never blocking -- adapts the capacity, and raises
    ``D4Error`` if the previous render overflowed (its tiles beyond the capacity were rendered empty).  Use
    ``check()`` after a ``torch.cuda.synchronize()`` for a definite answer, e.g. once per K steps.

    The first render through a fresh object runs in the ordinary synchronising mode to learn the sizes."""

    RING = 16  # pinned result slots: one per binning in flight

    def __init__(self, headroom: float = 1.3):
        self.headroom = float(headroom)
        self.n_isects = 0   # capacity of the per-intersection buffers (0: not known yet)
        self.sort_cap = 0   # capacity of the per-tile shared-memory sort
        self.seen_isects = self.seen_tile = 0
        self.overflowed = False
        self._host = None   # pinned int64 [RING, 3]: n_isects, max per tile, overflow flag
        self._pending = []  # (event | None, slot) in issue order
        self._slot = 0
        self._stream = None
        self._last_n = 0

    @property
    def ready(self) -> bool:
        return self.n_isects > 0

    def learn(self, n_isects: int, max_count: int):
        self.seen_isects, self.seen_tile = max(self.seen_isects, n_isects), max(self.seen_tile, max_count)
        cap_max = _cabi.lib().d4_tile_sort_capacity_max()
        self.n_isects = max(self.n_isects, min(int(self.seen_isects * self.headroom) + 1024, 2 ** 31 - 1))
        want = max(1024, int(self.seen_tile * 1.5))
        sort_cap = 1
        while sort_cap < want:
            sort_cap <<= 1
        self.sort_cap = max(self.sort_cap, min(sort_cap, cap_max))
        self._last_n = n_isects

    def record(self, stats: Tensor):
        """Enqueue the async copy of (n_isects, max per tile, overflow) to pinned memory.  Outside graph capture the copy
        runs on a side stream: on the compute stream it would queue behind whatever large device -> host download the
        caller has in flight on the copy engine, and stall the kernels behind it."""
        if self._host is None:
            self._host = torch.zeros((self.RING, 3), dtype=torch.int64).pin_memory()
        if len(self._pending) >= self.RING:
            self._drain(block=True)
        slot = self._slot
        self._slot = (slot + 1) % self.RING
        if torch.cuda.is_current_stream_capturing():
            self._host[slot].copy_(stats[:3], non_blocking=True)  # replayed with the graph: read by check()
            self._pending = [(None, slot)]
            return
        cur = torch.cuda.current_stream()
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=stats.device)
        binned = torch.cuda.Event()
        binned.record(cur)
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(binned)
            self._host[slot].copy_(stats[:3], non_blocking=True)
            stats.record_stream(self._stream)
            ev = torch.cuda.Event()
            ev.record(self._stream)
        self._pending.append((ev, slot))

    def _consume(self, slot: int):
        n, mx, ovf = (int(x) for x in self._host[slot].tolist())
        self.learn(n, mx)
        if ovf:
            self.overflowed = True
            self._host[slot, 2] = 0
            raise _cabi.D4Error(f"tile binning overflowed its capacity in a previous render ({n} intersections, {mx} in one "
                                f"tile; that render dropped the tiles beyond it) -- capacity raised to {self.n_isects} / "
                                f"{self.sort_cap}, render again")

    def _drain(self, block: bool):
        while self._pending:
            ev, slot = self._pending[0]
            if ev is not None:
                if block:
                    ev.synchronize()
                elif not ev.query():
                    return
            elif not block:
                return  # recorded under graph capture: only check() (after a synchronize) may read it
            self._pending.pop(0)
            self._consume(slot)

    @property
    def last_n_isects(self) -> int:
        """Intersection count of the last binning that poll() / check() has looked at (or of the learning render)."""
        return self._last_n

    def poll(self):
        """Non-blocking: look at the binnings that have finished, if any."""
        if not torch.cuda.is_current_stream_capturing():
            self._drain(block=False)

    def check(self):
        """Blocking: validate every binning issued so far (call after a synchronize when replaying a CUDA graph)."""
        pend = list(self._pending)
        try:
            self._drain(block=True)
        finally:
            if pend and pend[-1][0] is None:  # a captured graph keeps writing its slot on every replay
                self._pending = [pend[-1]]


def _check_cuda(*ts):
    _cabi.check_tensors(*ts, what="deblur4dgs_b200 render ops")


def _f32c(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _cam_layout(means: Tensor, quats: Tensor, viewmats: Tensor, Ks: Tensor, G: int):
    """Number of cameras and element strides between cameras (0 = shared)."""
    C = max(viewmats.shape[0], means.shape[0] if means.dim() == 3 else 1, quats.shape[0] if quats.dim() == 3 else 1)
    def stride(t, per, batched_dim):
        if t.dim() == batched_dim:
            if t.shape[0] == C:
                return per
            if t.shape[0] == 1:
                return 0
            raise ValueError("camera dimension mismatch")
        return 0
    ms = stride(means, G * 3, 3)
    qs = stride(quats, G * 4, 3)
    vs = 16 if viewmats.shape[0] == C else (0 if viewmats.shape[0] == 1 else None)
    ks = 9 if Ks.shape[0] == C else (0 if Ks.shape[0] == 1 else None)
    if vs is None or ks is None:
        raise ValueError("viewmats / Ks must have 1 or C cameras")
    return C, ms, qs, vs, ks


# --------------------------------------------------------------------------- #
# projection (SURVEY rows a8 + a12)
# --------------------------------------------------------------------------- #
class _Projection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
                tile_size, row0=None, window_height=0):
        _check_cuda(means, quats, scales, viewmats, Ks, row0)
        means, quats, scales, viewmats, Ks = map(_f32c, (means, quats, scales, viewmats, Ks))
        G = scales.shape[0]
        C, ms, qs, vs, ks = _cam_layout(means, quats, viewmats, Ks, G)
        dev = means.device
        tile_w, tile_h = math.ceil(width / tile_size), math.ceil((window_height or height) / tile_size)
        if row0 is not None:
            row0 = row0.to(torch.int32).contiguous()
            assert row0.shape == (C,) and window_height > 0
        radii = torch.empty((C, G), dtype=torch.int32, device=dev)
        means2d = torch.empty((C, G, 2), dtype=torch.float32, device=dev)
        depths = torch.empty((C, G), dtype=torch.float32, device=dev)
        conics = torch.empty((C, G, 3), dtype=torch.float32, device=dev)
        tiles_per_gauss = torch.empty((C, G), dtype=torch.int32, device=dev)
        call("d4_project_fwd", ptr(means), ms, ptr(quats), qs, ptr(scales), ptr(viewmats), vs, ptr(Ks), ks, C, G,
             width, height, eps2d, near_plane, far_plane, radius_clip, tile_size, tile_w, tile_h, ptr(radii),
             ptr(means2d), ptr(depths), ptr(conics), ptr(tiles_per_gauss), ptr(row0), int(window_height), stream_ptr())
        ctx.save_for_backward(means, quats, scales, viewmats, Ks, radii, conics)
        ctx.cfg = (C, G, ms, qs, vs, ks, width, height, eps2d)
        ctx.mark_non_differentiable(radii, tiles_per_gauss)
        ctx.set_materialize_grads(False)  # no zero-filled [C,G] "gradients" for the integer outputs
        return radii, means2d, depths, conics, tiles_per_gauss

    @staticmethod
    def backward(ctx, _v_radii, v_means2d, v_depths, v_conics, _v_tpg):
        means, quats, scales, viewmats, Ks, radii, conics = ctx.saved_tensors
        C, G, ms, qs, vs, ks, width, height, eps2d = ctx.cfg
        dev = means.device
        z3 = 0
        if v_means2d is None or v_depths is None or v_conics is None:  # an output nothing downstream used
            v_means2d = _f32c(v_means2d) if v_means2d is not None else torch.zeros((C, G, 2), device=dev)
            v_depths = _f32c(v_depths) if v_depths is not None else torch.zeros((C, G), device=dev)
            v_conics = _f32c(v_conics) if v_conics is not None else torch.zeros((C, G, 3), device=dev)
            z3 = 1
        else:
            v_means2d, v_depths, v_conics = _f32c(v_means2d), _f32c(v_depths), _f32c(v_conics)
        want_vm = ctx.needs_input_grad[3]
        # one zero-filled workspace for the accumulated outputs (a single memset instead of four)
        n_m, n_q, n_s, n_v = means.numel(), quats.numel(), scales.numel(), (C * 16 if want_vm else 0)
        ws = torch.zeros((n_m + n_q + n_s + n_v,), dtype=torch.float32, device=dev)
        _cabi.count_fill(1 + z3)
        v_means = ws[:n_m].view(means.shape)
        v_quats = ws[n_m:n_m + n_q].view(quats.shape)
        v_scales = ws[n_m + n_q:n_m + n_q + n_s].view(scales.shape)
        v_vm_full = ws[n_m + n_q + n_s:].view(C, 4, 4) if want_vm else None
        call("d4_project_bwd", ptr(means), ms, ptr(quats), qs, ptr(scales), ptr(viewmats), vs, ptr(Ks), ks, C, G,
             width, height, eps2d, ptr(radii), ptr(conics), ptr(v_means2d), ptr(v_depths), ptr(v_conics),
             ptr(v_means), ptr(v_quats), ptr(v_scales), ptr(v_vm_full), stream_ptr())
        v_viewmats = None
        if want_vm:
            v_viewmats = v_vm_full if viewmats.shape[0] == C else v_vm_full.sum(0, keepdim=True)
        return (v_means, v_quats, v_scales, v_viewmats, None, None, None, None, None, None, None, None, None, None)


def fully_fused_projection(means, quats, scales, viewmats, Ks, width, height, eps2d=0.3, near_plane=0.01,
                           far_plane=1e10, radius_clip=0.0, tile_size=16, row_windows=None):
    """gsplat.fully_fused_projection (packed=False): returns (radii i32 [C,G], means2d [C,G,2],
    depths [C,G], conics [C,G,3], tiles_per_gauss i32 [C,G]).

    row_windows = (row0 i32 [C], window_height): camera c covers rows [row0[c], row0[c] + window_height) of the
    image only (multi-GPU tile-row bands); means2d.y comes back relative to the window."""
    row0, wh = (None, 0) if row_windows is None else row_windows
    return _Projection.apply(means, quats, scales, viewmats, Ks, int(width), int(height), float(eps2d),
                             float(near_plane), float(far_plane), float(radius_clip), int(tile_size), row0, int(wh))


# --------------------------------------------------------------------------- #
# tile binning -- integer, not differentiable
# --------------------------------------------------------------------------- #
@torch.no_grad()
def isect_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_width: int, tile_height: int,
                tiles_per_gauss: Optional[Tensor] = None, sort: bool = True):
    """gsplat.isect_tiles: returns (tiles_per_gauss i32 [C,G], isect_ids i64 [I], flatten_ids i32 [I]),
    sorted by (camera, tile, depth bits) with a stable radix sort."""
    _check_cuda(means2d, radii, depths)
    C, G = radii.shape
    dev = means2d.device
    st = stream_ptr()
    if tiles_per_gauss is None:
        raise ValueError("tiles_per_gauss comes from fully_fused_projection (fused first pass)")
    n = C * G
    ws_bytes = _cabi.lib().d4_scan_workspace_bytes(n)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    cum = torch.empty((C, G), dtype=torch.int32, device=dev)
    total = torch.empty((1,), dtype=torch.int64, device=dev)
    call("d4_exclusive_scan_i32", ptr(tiles_per_gauss), n, ptr(cum), ptr(total), ptr(ws), ws_bytes, st)
    n_isects = int(total.item())  # the one device->host sync of the op (as in gsplat)
    if n_isects >= 2 ** 31:
        raise _cabi.D4Error("more than 2^31 tile intersections")
    isect_ids = torch.empty((n_isects,), dtype=torch.int64, device=dev)
    flatten_ids = torch.empty((n_isects,), dtype=torch.int32, device=dev)
    if n_isects == 0:
        return tiles_per_gauss, isect_ids, flatten_ids
    call("d4_isect_emit", ptr(means2d), ptr(radii), ptr(depths), ptr(cum), C, G, tile_size, tile_width,
         tile_height, ptr(isect_ids), ptr(flatten_ids), st)
    if sort:
        tile_n_bits = _cabi.lib().d4_tile_n_bits(tile_width * tile_height)
        cam_n_bits = int(math.floor(math.log2(C))) + 1
        isect_ids, flatten_ids = sort_pairs(isect_ids, flatten_ids, 0, 32 + tile_n_bits + cam_n_bits)
    return tiles_per_gauss, isect_ids, flatten_ids


@torch.no_grad()
def sort_pairs(keys: Tensor, vals: Tensor, begin_bit: int = 0, end_bit: int = 64):
    """Stable LSD radix sort of (int64 key >= 0, int32 value) pairs on key bits [begin_bit, end_bit)."""
    _check_cuda(keys, vals)
    n = keys.shape[0]
    dev = keys.device
    keys_b, vals_b = torch.empty_like(keys), torch.empty_like(vals)
    ws_bytes = _cabi.lib().d4_sort_workspace_bytes(n)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    in_b = ctypes.c_int(0)
    call("d4_sort_pairs_u64", ptr(keys), ptr(vals), ptr(keys_b), ptr(vals_b), n, begin_bit, end_bit, ptr(ws),
         ws_bytes, ctypes.byref(in_b), stream_ptr())
    return (keys_b, vals_b) if in_b.value else (keys, vals)


@torch.no_grad()
def isect_offset_encode(isect_ids: Tensor, C: int, tile_width: int, tile_height: int) -> Tensor:
    """gsplat.isect_offset_encode: offsets i32 [C, tile_height, tile_width]."""
    offsets = torch.empty((C, tile_height, tile_width), dtype=torch.int32, device=isect_ids.device)
    call("d4_tile_offsets", ptr(isect_ids), isect_ids.shape[0], C, tile_width, tile_height, ptr(offsets), stream_ptr())
    return offsets


@torch.no_grad()
def bin_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_width: int, tile_height: int,
              tiles_per_gauss: Optional[Tensor] = None, method: str = "auto", pack=None,
              capacity: Optional["RenderCapacity"] = None):
    """Tile binning, both halves of SURVEY row a9 in one call: returns
    (isect_ids i64 [I], flatten_ids i32 [I], isect_offsets i32 [C,th,tw]) -- exactly what
    gsplat.isect_tiles + gsplat.isect_offset_encode produce.

    method "bucket": count per (camera, tile) -> scan (== offsets) -> emit into tile segments ->
    per-tile shared-memory sort; "radix": gsplat's structure (emit, global LSD radix sort, offset
    encode); "auto": bucket unless a tile holds more entries than the shared-memory sort can take.

    pack = (conics [C,G,3], opacities [G], blend_depths [C,G] | None): additionally build the packed record
    slabs of the slab blend kernels (include/d4gs.h: d4_isect_pack) and return (..., recs, rec_counts).

    capacity: a RenderCapacity -> sync-free mode once it has learnt the sizes (the first call synchronises):
    isect_ids / flatten_ids / recs then hold ``capacity.n_isects`` entries, of which the first n_isects (a device
    value) are meaningful."""
    _check_cuda(means2d, radii, depths)
    C, G = radii.shape
    dev = means2d.device
    st = stream_ptr()
    n_seg = C * tile_width * tile_height

    def packed(isect_ids, flatten_ids, offsets, recs=None, rec_counts=None):
        if pack is None:
            return isect_ids, flatten_ids, offsets
        if recs is None:
            conics, opacities, bdepths = pack
            n = flatten_ids.shape[0]
            recs = torch.empty((max(n, 1), 8), dtype=torch.float32, device=dev)
            rec_counts = torch.empty((n_seg,), dtype=torch.int32, device=dev)
            call("d4_isect_pack", ptr(means2d), ptr(conics), ptr(opacities), ptr(bdepths), C, G, tile_size, tile_width,
                 tile_height, ptr(offsets), ptr(flatten_ids), n, ptr(recs), ptr(rec_counts), st)
        return isect_ids, flatten_ids, offsets, recs, rec_counts

    if (capacity is not None and capacity.ready and pack is not None and method != "radix"
            and capacity.seen_tile <= _cabi.lib().d4_tile_sort_capacity_max()):
        # ---- capacity mode: no device -> host read-back; one zero-filled int workspace (counts | cursors | stats)
        capacity.poll()
        lib = _cabi.lib()
        cap = capacity.n_isects
        # fixed-stride buckets (tile t owns keys[t * stride ...]): the emit counts while it writes, so the separate
        # count pass is gone -- emit -> scan of the counts (== isect_offsets) -> per-tile sort into the compact lists
        stride = capacity.sort_cap
        izero = torch.zeros((n_seg + 8,), dtype=torch.int32, device=dev)
        _cabi.count_fill()
        counts = izero[:n_seg]
        stats = izero[n_seg + (n_seg & 1):n_seg + (n_seg & 1) + 6].view(torch.int64)  # n_isects, max per tile, overflow
        keys = torch.empty((n_seg * stride,), dtype=torch.int64, device=dev)
        call("d4_bucket_emit", ptr(means2d), ptr(radii), ptr(depths), C, G, tile_size, tile_width, tile_height,
             None, ptr(counts), ptr(keys), cap, stride, st)
        offsets = torch.empty((C, tile_height, tile_width), dtype=torch.int32, device=dev)
        ws_bytes = lib.d4_scan_workspace_bytes(n_seg)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        call("d4_scan_counts", ptr(counts), n_seg, ptr(offsets), ptr(stats), ptr(ws), ws_bytes, st)
        isect_ids = torch.empty((cap,), dtype=torch.int64, device=dev)
        flatten_ids = torch.empty((cap,), dtype=torch.int32, device=dev)
        conics, opacities, bdepths = pack
        recs = torch.empty((cap, 8), dtype=torch.float32, device=dev)
        rec_counts = torch.empty((n_seg,), dtype=torch.int32, device=dev)
        call("d4_tile_sort_pack_cap", ptr(keys), ptr(offsets), ptr(stats), cap, capacity.sort_cap, C, tile_width,
             tile_height, ptr(isect_ids), ptr(flatten_ids), ptr(means2d), ptr(conics), ptr(opacities), ptr(bdepths), G,
             tile_size, ptr(recs), ptr(rec_counts), ptr(stats[2:]), stride, st)
        capacity.record(stats)
        return isect_ids, flatten_ids, offsets, recs, rec_counts

    if method == "radix":
        _, isect_ids, flatten_ids = isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height,
                                                tiles_per_gauss=tiles_per_gauss)
        return packed(isect_ids, flatten_ids, isect_offset_encode(isect_ids, C, tile_width, tile_height))
    izero = torch.zeros((2 * n_seg + 8,), dtype=torch.int32, device=dev)  # counts | cursors | stats, one fill
    _cabi.count_fill()
    counts, cursors = izero[:n_seg], izero[n_seg:2 * n_seg]
    stats = izero[2 * n_seg:2 * n_seg + 8].view(torch.int64)  # n_isects, max per tile
    call("d4_tile_count", ptr(means2d), ptr(radii), C, G, tile_size, tile_width, tile_height, ptr(counts), st)
    offsets = torch.empty((C, tile_height, tile_width), dtype=torch.int32, device=dev)
    ws_bytes = _cabi.lib().d4_scan_workspace_bytes(n_seg)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    call("d4_scan_counts", ptr(counts), n_seg, ptr(offsets), ptr(stats), ptr(ws), ws_bytes, st)
    n_isects, max_count = (int(x) for x in stats[:2].tolist())  # the one device->host sync of the op
    if n_isects >= 2 ** 31:
        raise _cabi.D4Error("more than 2^31 tile intersections")
    if capacity is not None:
        capacity.learn(n_isects, max_count)
    if max_count > _cabi.lib().d4_tile_sort_capacity():
        if method == "bucket":
            raise _cabi.D4Error(f"a tile holds {max_count} intersections: beyond the shared-memory sort capacity")
        return bin_tiles(means2d, radii, depths, tile_size, tile_width, tile_height, tiles_per_gauss, method="radix",
                         pack=pack)
    isect_ids = torch.empty((n_isects,), dtype=torch.int64, device=dev)
    flatten_ids = torch.empty((n_isects,), dtype=torch.int32, device=dev)
    if n_isects == 0:
        return packed(isect_ids, flatten_ids, offsets)
    keys = torch.empty((n_isects,), dtype=torch.int64, device=dev)
    call("d4_bucket_emit", ptr(means2d), ptr(radii), ptr(depths), C, G, tile_size, tile_width, tile_height,
         ptr(offsets), ptr(cursors), ptr(keys), n_isects, 0, st)
    if pack is None:
        call("d4_tile_sort", ptr(keys), ptr(offsets), n_isects, C, tile_width, tile_height, max_count, ptr(isect_ids),
             ptr(flatten_ids), st)
        return isect_ids, flatten_ids, offsets
    conics, opacities, bdepths = pack
    recs = torch.empty((n_isects, 8), dtype=torch.float32, device=dev)
    rec_counts = torch.empty((n_seg,), dtype=torch.int32, device=dev)
    call("d4_tile_sort_pack", ptr(keys), ptr(offsets), n_isects, C, tile_width, tile_height, max_count, ptr(isect_ids),
         ptr(flatten_ids), ptr(means2d), ptr(conics), ptr(opacities), ptr(bdepths), G, tile_size, ptr(recs),
         ptr(rec_counts), st)
    return isect_ids, flatten_ids, offsets, recs, rec_counts


# --------------------------------------------------------------------------- #
# blend (SURVEY rows a10 + a11)
# --------------------------------------------------------------------------- #
class _Blend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means2d, conics, opacities, colors, depths, backgrounds, isect_offsets, flatten_ids, width,
                height, tile_size, normalize_depth):
        _check_cuda(means2d, conics, opacities, colors)
        means2d, conics, opacities, colors = map(_f32c, (means2d, conics, opacities, colors))
        depths, backgrounds = _f32c(depths), _f32c(backgrounds)
        C, G = means2d.shape[:2]
        D0 = colors.shape[-1]
        ccs = 0 if colors.dim() == 2 else G * D0
        D = D0 + (1 if depths is not None else 0)
        dev = means2d.device
        tile_h, tile_w = isect_offsets.shape[1:]
        render_colors = torch.empty((C, height, width, D), dtype=torch.float32, device=dev)
        render_alphas = torch.empty((C, height, width, 1), dtype=torch.float32, device=dev)
        last_ids = torch.empty((C, height, width), dtype=torch.int32, device=dev)
        acc_depth = torch.empty((C, height, width), dtype=torch.float32, device=dev) if normalize_depth else None
        n_isects = flatten_ids.shape[0]
        # per-intersection hit masks (which 8x4 pixel blocks passed the alpha test): the backward visits only those
        needs_bwd = any(ctx.needs_input_grad[:5])
        use_masks = needs_bwd and HIT_MASKS
        hit_masks = torch.zeros((n_isects,), dtype=torch.uint8, device=dev) if use_masks else None
        call("d4_blend_fwd", ptr(means2d), ptr(conics), ptr(opacities), ptr(colors), ccs, ptr(depths),
             ptr(backgrounds), C, G, D0, width, height, tile_size, tile_w, tile_h, ptr(isect_offsets),
             ptr(flatten_ids), n_isects, int(normalize_depth), ptr(render_colors), ptr(render_alphas), ptr(last_ids),
             ptr(acc_depth), ptr(hit_masks), stream_ptr())
        if HIT_MASK_TAP is not None:
            HIT_MASK_TAP.append(hit_masks)
        # NOTE: render_colors is NOT saved -- the reference edits it in place (scene_model.py:391-393)
        ctx.save_for_backward(means2d, conics, opacities, colors, depths, backgrounds, isect_offsets, flatten_ids,
                              render_alphas, last_ids, acc_depth, hit_masks)
        ctx.cfg = (C, G, D0, ccs, width, height, tile_size, tile_w, tile_h, n_isects, int(normalize_depth))
        return render_colors, render_alphas

    @staticmethod
    def backward(ctx, v_render_colors, v_render_alphas):
        (means2d, conics, opacities, colors, depths, backgrounds, isect_offsets, flatten_ids, render_alphas, last_ids,
         acc_depth, hit_masks) = ctx.saved_tensors
        C, G, D0, ccs, width, height, tile_size, tile_w, tile_h, n_isects, normalize_depth = ctx.cfg
        dev = means2d.device
        D = D0 + (1 if depths is not None else 0)
        v_rc = _f32c(v_render_colors) if v_render_colors is not None else torch.zeros((C, height, width, D), device=dev)
        v_ra = _f32c(v_render_alphas) if v_render_alphas is not None else torch.zeros((C, height, width, 1), device=dev)
        v_means2d = torch.zeros_like(means2d)
        v_conics = torch.zeros_like(conics)
        v_colors = torch.zeros_like(colors)
        v_opacities = torch.zeros_like(opacities)
        v_depths = torch.zeros_like(depths) if depths is not None else None
        call("d4_blend_bwd", ptr(means2d), ptr(conics), ptr(opacities), ptr(colors), ccs, ptr(depths),
             ptr(backgrounds), C, G, D0, width, height, tile_size, tile_w, tile_h, ptr(isect_offsets),
             ptr(flatten_ids), n_isects, normalize_depth, ptr(render_alphas), ptr(last_ids), ptr(acc_depth),
             ptr(v_rc), ptr(v_ra), ptr(v_means2d), ptr(v_conics), ptr(v_colors), ptr(v_opacities), ptr(v_depths),
             ptr(hit_masks), int(BWD_MODE), stream_ptr())
        v_backgrounds = None
        if backgrounds is not None and ctx.needs_input_grad[5]:
            # as gsplat: sum over pixels of v_colors * (1 - alpha); the depth channel has no background
            v_backgrounds = (v_rc[..., :D0] * (1.0 - render_alphas)).sum(dim=(1, 2))
        return (v_means2d, v_conics, v_opacities, v_colors, v_depths, v_backgrounds, None, None, None, None, None, None)


class _BlendSlab(torch.autograd.Function):
    """Blend over the packed record slabs (csrc/slab.cuh).  means2d / conics / opacities / depths are inputs of the
    autograd node only -- the kernels read their values from ``recs`` -- so that the gradients land where
    gsplat's rasterize_to_pixels puts them."""

    @staticmethod
    def forward(ctx, means2d, conics, opacities, colors, depths, backgrounds, isect_offsets, recs, rec_counts, width,
                height, tile_size, normalize_depth):
        _check_cuda(means2d, conics, opacities, colors, recs)
        if backgrounds is not None and not backgrounds.is_contiguous():
            _cabi.count_fill()  # the [C,D0] background rows of an expand()ed view are materialised by a torch copy kernel
        colors, backgrounds = _f32c(colors), _f32c(backgrounds)
        C, G = means2d.shape[:2]
        D0 = colors.shape[-1]
        ccs = 0 if colors.dim() == 2 else G * D0
        with_depth = depths is not None
        D = D0 + int(with_depth)
        dev = means2d.device
        tile_h, tile_w = isect_offsets.shape[1:]
        n_seg = C * tile_h * tile_w
        render_colors = torch.empty((C, height, width, D), dtype=torch.float32, device=dev)
        render_alphas = torch.empty((C, height, width, 1), dtype=torch.float32, device=dev)
        last_ids = torch.empty((C, height, width), dtype=torch.int32, device=dev)
        acc_depth = torch.empty((C, height, width), dtype=torch.float32, device=dev) if normalize_depth else None
        needs_bwd = any(ctx.needs_input_grad[:5])
        hit_bits = None
        if needs_bwd or HIT_MASK_TAP is not None:
            words = _cabi.lib().d4_slab_hit_words(recs.shape[0], n_seg)
            alloc = torch.zeros if HIT_MASK_TAP is not None else torch.empty  # every word the backward reads is written
            hit_bits = alloc((words,), dtype=torch.int32, device=dev)
        fargs = (ptr(recs), ptr(isect_offsets), ptr(rec_counts), ptr(colors), ccs, ptr(backgrounds), C, G,
                 D0, int(with_depth), width, height, tile_size, tile_w, tile_h, int(normalize_depth), ptr(render_colors),
                 ptr(render_alphas), ptr(last_ids), ptr(acc_depth), ptr(hit_bits))
        if SLAB_FWD_VARIANT is None:
            call("d4_blend_fwd_slab", *fargs, stream_ptr())
        else:
            call("d4_blend_fwd_slab_variant", *fargs, int(SLAB_FWD_VARIANT), stream_ptr())
        if HIT_MASK_TAP is not None:
            HIT_MASK_TAP.append({"hit_bits": hit_bits, "recs": recs, "rec_counts": rec_counts,
                                 "isect_offsets": isect_offsets, "last_ids": last_ids})
        # NOTE: render_colors is NOT saved -- the reference edits it in place (scene_model.py:391-393)
        ctx.save_for_backward(colors, backgrounds, isect_offsets, recs, rec_counts, render_alphas, last_ids, acc_depth,
                              hit_bits)
        ctx.cfg = (C, G, D0, ccs, width, height, tile_size, tile_w, tile_h, int(normalize_depth), with_depth,
                   opacities.shape)
        return render_colors, render_alphas

    @staticmethod
    def backward(ctx, v_render_colors, v_render_alphas):
        colors, backgrounds, isect_offsets, recs, rec_counts, render_alphas, last_ids, acc_depth, hit_bits = ctx.saved_tensors
        C, G, D0, ccs, width, height, tile_size, tile_w, tile_h, normalize_depth, with_depth, oshape = ctx.cfg
        dev = colors.device
        D = D0 + int(with_depth)
        v_rc = _f32c(v_render_colors) if v_render_colors is not None else torch.zeros((C, height, width, D), device=dev)
        v_ra = _f32c(v_render_alphas) if v_render_alphas is not None else torch.zeros((C, height, width, 1), device=dev)
        # one zero-filled workspace for every accumulated gradient (a single memset instead of five)
        n_m, n_c, n_col, n_o, n_d = C * G * 2, C * G * 3, colors.numel(), G, (C * G if with_depth else 0)
        ws = torch.zeros((n_m + n_c + n_col + n_o + n_d,), dtype=torch.float32, device=dev)
        _cabi.count_fill()
        # colour rows first, then the xy pairs: both start 8-byte aligned (vectorised reductions in the kernel)
        v_colors = ws[:n_col].view(colors.shape)
        v_means2d = ws[n_col:n_col + n_m].view(C, G, 2)
        v_conics = ws[n_col + n_m:n_col + n_m + n_c].view(C, G, 3)
        v_opacities = ws[n_col + n_m + n_c:n_col + n_m + n_c + n_o].view(oshape)
        v_depths = ws[n_col + n_m + n_c + n_o:].view(C, G) if with_depth else None
        args = (ptr(recs), ptr(isect_offsets), ptr(rec_counts), ptr(colors), ccs, ptr(backgrounds), C, G,
                D0, int(with_depth), width, height, tile_size, tile_w, tile_h, normalize_depth, ptr(render_alphas),
                ptr(last_ids), ptr(acc_depth), ptr(v_rc), ptr(v_ra), ptr(hit_bits), ptr(v_means2d), ptr(v_conics),
                ptr(v_colors), ptr(v_opacities), ptr(v_depths))
        if SLAB_BWD_VARIANT is None:
            call("d4_blend_bwd_slab", *args, stream_ptr())
        else:
            call("d4_blend_bwd_slab_variant", *args, int(SLAB_BWD_VARIANT), stream_ptr())
        v_backgrounds = None
        if backgrounds is not None and ctx.needs_input_grad[5]:
            v_backgrounds = (v_rc[..., :D0] * (1.0 - render_alphas)).sum(dim=(1, 2))
        return (v_means2d, v_conics, v_opacities, v_colors, v_depths, v_backgrounds, None, None, None, None, None, None,
                None)


def rasterize_slabs(means2d, conics, colors, opacities, image_width, image_height, tile_size, isect_offsets, recs,
                    rec_counts, backgrounds=None, depths=None, normalize_depth=False):
    """Blend over packed record slabs (``bin_tiles(..., pack=...)``): the slab counterpart of
    ``rasterize_to_pixels``.  ``colors`` [G,D0] / [C,G,D0] with D0 in SLAB_D0."""
    if colors.shape[-1] not in SLAB_D0:
        raise _cabi.D4Error(f"colour width {colors.shape[-1]} is not built for the slab path; rasterization() pads")
    return _BlendSlab.apply(means2d, conics, opacities, colors, depths, backgrounds, isect_offsets, recs, rec_counts,
                            int(image_width), int(image_height), int(tile_size), bool(normalize_depth))


def rasterize_to_pixels(means2d, conics, colors, opacities, image_width, image_height, tile_size, isect_offsets,
                        flatten_ids, backgrounds=None, depths=None, normalize_depth=False):
    """gsplat.rasterize_to_pixels (packed=False).  ``colors`` is [G,D0] (shared by all
    cameras) or [C,G,D0]; ``opacities`` [G]; optional ``depths`` [C,G] is blended as one
    extra (last) channel, normalised to expected depth when ``normalize_depth``."""
    D = colors.shape[-1] + (1 if depths is not None else 0)
    if D not in SUPPORTED_D:
        raise _cabi.D4Error(f"channel count {D} is not built; rasterization() pads/chunks automatically")
    return _Blend.apply(means2d, conics, opacities, colors, depths, backgrounds, isect_offsets, flatten_ids,
                        int(image_width), int(image_height), int(tile_size), bool(normalize_depth))


# --------------------------------------------------------------------------- #
# the operator
# --------------------------------------------------------------------------- #
def _pad_channels(colors: Tensor, backgrounds: Optional[Tensor], has_depth: bool):
    D0 = colors.shape[-1]
    D = D0 + int(has_depth)
    target = next(d for d in SUPPORTED_D if d >= D)
    pad = target - D
    if pad:
        colors = torch.cat([colors, colors.new_zeros(colors.shape[:-1] + (pad,))], dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, backgrounds.new_zeros(backgrounds.shape[:-1] + (pad,))], dim=-1)
    return colors, backgrounds, pad


def rasterization(
    means: Tensor,  # [G,3]  (or [C,G,3]: per-camera centres, sub-exposure batch)
    quats: Tensor,  # [G,4] wxyz, un-normalised allowed (or [C,G,4])
    scales: Tensor,  # [G,3]
    opacities: Tensor,  # [G]
    colors: Tensor,  # [G,D0] or [C,G,D0]
    viewmats: Tensor,  # [C,4,4]
    Ks: Tensor,  # [C,3,3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = False,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: str = "classic",
    channel_chunk: int = CHANNEL_CHUNK,
    capacity: Optional[RenderCapacity] = None,  # extension: sync-free binning (see RenderCapacity)
    row_windows=None,  # extension: (row0 i32 [C], window_height) -- camera c renders that row band only (parallel.py)
    **unsupported,
) -> Tuple[Tensor, Tensor, Dict]:
    """Drop-in for ``gsplat.rendering.rasterization`` (gsplat==1.1.1) as called at
    flow3d/scene_model.py:360-373.  Supports the argument combinations the reference uses
    (Appendix A of SURVEY.md): packed=False, sh_degree=None, classic mode, RGB / RGB+D / RGB+ED."""
    if unsupported:
        raise NotImplementedError(f"unsupported rasterization arguments: {sorted(unsupported)}")
    if packed or sparse_grad or absgrad or sh_degree is not None or rasterize_mode != "classic":
        raise NotImplementedError("only packed=False, sparse_grad=False, absgrad=False, sh_degree=None, "
                                  "rasterize_mode='classic' are built (what Deblur4DGS uses)")
    if render_mode not in ("RGB", "RGB+D", "RGB+ED", "D", "ED"):
        raise ValueError(f"render_mode {render_mode!r}")
    if tile_size != 16:
        raise NotImplementedError("tile_size 16 only")
    G = scales.shape[0]
    assert means.shape[-2:] == (G, 3) and quats.shape[-2:] == (G, 4) and scales.shape == (G, 3)
    assert opacities.shape == (G,), "opacities must be [G] (flow3d/params.py:76-77)"
    assert viewmats.shape[-2:] == (4, 4) and Ks.shape[-2:] == (3, 3)

    radii, means2d, depths, conics, tiles_per_gauss = fully_fused_projection(
        means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip, tile_size,
        row_windows=row_windows)
    C = radii.shape[0]
    if row_windows is not None:
        height = int(row_windows[1])  # binning and blend work on C images of width x window_height
    tile_width, tile_height = math.ceil(width / tile_size), math.ceil(height / tile_size)
    with_depth = render_mode in ("RGB+D", "RGB+ED", "D", "ED")
    normalize = render_mode in ("RGB+ED", "ED")
    blend_depths = depths if with_depth else None
    slab = BLEND_PATH == "slab" and G < 2 ** 24
    opac_c = _f32c(opacities)
    if slab:
        isect_ids, flatten_ids, isect_offsets, recs, rec_counts = bin_tiles(
            means2d, radii, depths, tile_size, tile_width, tile_height, tiles_per_gauss=tiles_per_gauss,
            method=BINNING_METHOD, pack=(conics.detach(), opac_c.detach(), None if blend_depths is None else blend_depths.detach()),
            capacity=capacity)
    else:
        isect_ids, flatten_ids, isect_offsets = bin_tiles(means2d, radii, depths, tile_size, tile_width, tile_height,
                                                          tiles_per_gauss=tiles_per_gauss, method=BINNING_METHOD)

    meta = {
        "camera_ids": None, "gaussian_ids": None, "radii": radii, "means2d": means2d, "depths": depths,
        "conics": conics, "opacities": opacities[None].expand(C, -1), "tile_width": tile_width,
        "tile_height": tile_height, "tiles_per_gauss": tiles_per_gauss, "isect_ids": isect_ids,
        "flatten_ids": flatten_ids, "isect_offsets": isect_offsets, "width": width, "height": height,
        "tile_size": tile_size, "n_cameras": C,
    }

    if render_mode in ("D", "ED"):
        colors = colors.new_zeros(colors.shape[:-1] + (0,))
        backgrounds = None if backgrounds is None else backgrounds.new_zeros(backgrounds.shape[:-1] + (0,))
    D0 = colors.shape[-1]

    def blend(col, bg, dd, norm):
        """One blend call on <= channel_chunk colour channels (+ the depth channel when dd is given)."""
        d0 = col.shape[-1]
        if slab:
            target = next(d for d in SLAB_D0 if d >= max(d0, 1))
            pad = target - d0
            if pad:
                col = torch.cat([col, col.new_zeros(col.shape[:-1] + (pad,))], dim=-1)
                bg = None if bg is None else torch.cat([bg, bg.new_zeros(bg.shape[:-1] + (pad,))], dim=-1)
            rc, ra = rasterize_slabs(means2d, conics, col, opac_c, width, height, tile_size, isect_offsets, recs,
                                     rec_counts, backgrounds=bg, depths=dd, normalize_depth=norm)
        else:
            col, bg, pad = _pad_channels(col, bg, dd is not None)
            rc, ra = rasterize_to_pixels(means2d, conics, col, opacities, width, height, tile_size, isect_offsets,
                                         flatten_ids, backgrounds=bg, depths=dd, normalize_depth=norm)
        if pad:
            rc = torch.cat([rc[..., :d0], rc[..., d0 + pad:]], dim=-1)
        return rc, ra

    max_d0 = max(SLAB_D0) if slab else max(SUPPORTED_D) - int(with_depth)
    if D0 <= min(max_d0, channel_chunk + (0 if slab else 1)):
        render_colors, render_alphas = blend(colors, backgrounds, blend_depths, normalize)
    else:
        # channel chunking, as gsplat does for > channel_chunk channels
        chunks, render_alphas = [], None
        n_chunks = (D0 + channel_chunk - 1) // channel_chunk
        for i in range(n_chunks):
            last = i == n_chunks - 1
            col = colors[..., i * channel_chunk:(i + 1) * channel_chunk]
            bg = None if backgrounds is None else backgrounds[..., i * channel_chunk:(i + 1) * channel_chunk]
            rc, ra = blend(col, bg, blend_depths if last else None, normalize and last)
            chunks.append(rc)
            render_alphas = ra if render_alphas is None else render_alphas
        render_colors = torch.cat(chunks, dim=-1)
    return render_colors, render_alphas, meta
