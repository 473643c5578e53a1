"""Adaptive-density-control statistics on top of libd4gs.so.

``accumulate_densify_stats`` replaces the per-render loop of ``Trainer._prepare_control_step``
(flow3d/trainer.py:967-989) for the N sub-exposure renders of one frame: it consumes the screen-space
gradient ``meta["means2d"].grad`` ([N,G,2]) and ``meta["radii"]`` ([N,G]) that
``scene.render_subexposures`` / ``rendering.rasterization`` expose, and updates the trainer's
``running_stats`` tensors in place with one kernel (no atomics, one pass over N*G*12 bytes).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from ._cabi import D4Error, call, check_tensors, ptr, stream_ptr


@torch.no_grad()
def accumulate_densify_stats(running_stats: Dict[str, Tensor], means2d_grad: Tensor, radii: Tensor,
                             img_wh: Tuple[int, int], batch_size: int = 1, n_renders: Optional[int] = None,
                             update_max_radii: bool = False) -> None:
    """running_stats: {"xys_grad_norm_acc" f32 [G], "vis_count" i64 [G], "max_radii" f32 [G]} (trainer.py
    running_stats); means2d_grad [N,G,2]; radii int32 [N,G].  ``update_max_radii=False`` reproduces the
    reference literally: its ``index_put`` (no underscore, trainer.py:989) discards the maximum."""
    check_tensors(means2d_grad, radii, what="control ops")
    g = means2d_grad.float().contiguous()
    r = radii.to(torch.int32).contiguous()
    N, G = r.shape
    n_renders = N if n_renders is None else n_renders
    W, H = img_wh
    acc, cnt, mr = running_stats["xys_grad_norm_acc"], running_stats["vis_count"], running_stats["max_radii"]
    assert acc.dtype == torch.float32 and cnt.dtype == torch.int64 and acc.is_contiguous() and cnt.is_contiguous()
    call("d4_densify_stats", ptr(g), ptr(r), N, G, W / 2.0 * batch_size * n_renders, H / 2.0 * batch_size * n_renders,
         1.0 / max(W, H), ptr(acc), ptr(cnt), ptr(mr) if update_max_radii else None, stream_ptr())
