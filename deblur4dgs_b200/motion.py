"""Motion-basis deformation ops on top of libd4gs.so.

``deform_subexposures`` replaces, for all N sub-exposures of a frame at once,
what the reference computes per loop iteration at flow3d/scene_model.py:323-353:
``GaussianParams`` activations (params.py:39-43), ``MotionBases.compute_transforms``
(params.py:142-180), ``cont_6d_to_rmat`` (transforms.py:41-53),
``SceneModel.compute_poses_fg/all`` (scene_model.py:76-120) and the camera
sub-exposure transform (scene_model.py:352-353).

``compute_poses_all`` / ``compute_poses_fg`` keep the reference's method names
and ``[G,B,*]`` output layout for its other callers (trainer.py:478,485).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from ._cabi import D4Error, call, check_tensors, count_fill, ptr, stream_ptr


def _c(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _Deform(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fg_means, fg_quats, motion_coefs, bg_means, bg_quats, rots, transls, times, RTs):
        check_tensors(fg_means, fg_quats, motion_coefs, bg_means, bg_quats, rots, transls, times, RTs, what="deform ops")
        fg_means, fg_quats, motion_coefs, bg_means, bg_quats, rots, transls, times, RTs = map(
            _c, (fg_means, fg_quats, motion_coefs, bg_means, bg_quats, rots, transls, times, RTs))
        Gf = fg_means.shape[0]
        Gb = 0 if bg_means is None else bg_means.shape[0]
        K, T = rots.shape[0], rots.shape[1]
        N = times.shape[0]
        G = Gf + Gb
        dev = fg_means.device
        out_means = torch.empty((N, G, 3), dtype=torch.float32, device=dev)
        out_quats = torch.empty((N, G, 4), dtype=torch.float32, device=dev)
        call("d4_deform_fwd", ptr(fg_means), ptr(fg_quats), ptr(motion_coefs), ptr(bg_means), ptr(bg_quats),
             ptr(rots), ptr(transls), ptr(times), ptr(RTs), Gf, Gb, K, T, N, ptr(out_means), ptr(out_quats),
             stream_ptr())
        ctx.save_for_backward(fg_means, fg_quats, motion_coefs, bg_means, bg_quats, rots, transls, times, RTs)
        ctx.cfg = (Gf, Gb, K, T, N)
        return out_means, out_quats

    @staticmethod
    def backward(ctx, v_means, v_quats):
        fg_means, fg_quats, motion_coefs, bg_means, bg_quats, rots, transls, times, RTs = ctx.saved_tensors
        Gf, Gb, K, T, N = ctx.cfg
        dev = fg_means.device
        G = Gf + Gb
        v_means = _c(v_means) if v_means is not None else torch.zeros((N, G, 3), device=dev)
        v_quats = _c(v_quats) if v_quats is not None else torch.zeros((N, G, 4), device=dev)
        v_fg_means = torch.empty_like(fg_means)
        v_fg_quats = torch.empty_like(fg_quats)
        v_coefs = torch.empty_like(motion_coefs)
        v_bg_means = torch.empty_like(bg_means) if bg_means is not None else None
        v_bg_quats = torch.empty_like(bg_quats) if bg_quats is not None else None
        # one zero-filled workspace for the accumulated outputs (a single memset instead of four)
        n_r, n_t, n_ti, n_rt = rots.numel(), transls.numel(), times.numel(), (RTs.numel() if RTs is not None else 0)
        ws = torch.zeros((n_r + n_t + n_ti + n_rt,), dtype=torch.float32, device=dev)
        count_fill()
        v_rots = ws[:n_r].view(rots.shape)
        v_transls = ws[n_r:n_r + n_t].view(transls.shape)
        v_times = ws[n_r + n_t:n_r + n_t + n_ti].view(times.shape)
        v_RTs = ws[n_r + n_t + n_ti:].view(RTs.shape) if RTs is not None else None
        call("d4_deform_bwd", ptr(fg_means), ptr(fg_quats), ptr(motion_coefs), ptr(bg_means), ptr(bg_quats),
             ptr(rots), ptr(transls), ptr(times), ptr(RTs), Gf, Gb, K, T, N, ptr(v_means), ptr(v_quats),
             ptr(v_fg_means), ptr(v_fg_quats), ptr(v_coefs), ptr(v_bg_means), ptr(v_bg_quats), ptr(v_rots),
             ptr(v_transls), ptr(v_times), ptr(v_RTs), stream_ptr())
        return v_fg_means, v_fg_quats, v_coefs, v_bg_means, v_bg_quats, v_rots, v_transls, v_times, v_RTs


def deform_subexposures(fg_means: Tensor, fg_quats: Tensor, motion_coefs: Tensor, bg_means: Optional[Tensor],
                        bg_quats: Optional[Tensor], rots: Tensor, transls: Tensor, times: Tensor,
                        RTs: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """All N sub-exposures in one launch.

    fg_means [Gf,3], fg_quats [Gf,4] (raw wxyz), motion_coefs [Gf,K] (raw logits),
    bg_means [Gb,3] / bg_quats [Gb,4] (raw) or None, rots [K,T,6], transls [K,T,3],
    times [N] (float frame coordinates), RTs [N,3,4] camera deltas or None.
    Returns means [N,G,3], quats [N,G,4] (unit, wxyz), fg first then bg."""
    times = times.reshape(-1).to(torch.float32)
    return _Deform.apply(fg_means, fg_quats, motion_coefs, bg_means, bg_quats, rots, transls, times, RTs)


def compute_poses_fg(fg_means, fg_quats, motion_coefs, rots, transls, ts) -> Tuple[Tensor, Tensor]:
    """SceneModel.compute_poses_fg (scene_model.py:76-106): ts [B] -> means [Gf,B,3], quats [Gf,B,4]."""
    m, q = deform_subexposures(fg_means, fg_quats, motion_coefs, None, None, rots, transls, ts, None)
    return m.permute(1, 0, 2), q.permute(1, 0, 2)


def compute_poses_all(fg_means, fg_quats, motion_coefs, bg_means, bg_quats, rots, transls, ts):
    """SceneModel.compute_poses_all (scene_model.py:108-120): fg (deformed) first, bg (static) after."""
    m, q = deform_subexposures(fg_means, fg_quats, motion_coefs, bg_means, bg_quats, rots, transls, ts, None)
    return m.permute(1, 0, 2), q.permute(1, 0, 2)


class _ComputeTransforms(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ts, coefs, rots, transls):
        check_tensors(ts, coefs, rots, transls, what="deform ops")
        ts, coefs, rots, transls = map(_c, (ts, coefs, rots, transls))
        G, K = coefs.shape
        T, B = rots.shape[1], ts.shape[0]
        out = torch.empty((G, B, 3, 4), dtype=torch.float32, device=coefs.device)
        call("d4_compute_transforms_fwd", ptr(coefs), ptr(rots), ptr(transls), ptr(ts), G, K, T, B, ptr(out),
             stream_ptr())
        ctx.save_for_backward(ts, coefs, rots, transls)
        return out

    @staticmethod
    def backward(ctx, v_out):
        ts, coefs, rots, transls = ctx.saved_tensors
        G, K = coefs.shape
        T, B = rots.shape[1], ts.shape[0]
        v_coefs = torch.empty_like(coefs)
        v_rots, v_transls, v_ts = torch.zeros_like(rots), torch.zeros_like(transls), torch.zeros_like(ts)
        v_out = _c(v_out)  # bound to a name: must outlive the launch
        call("d4_compute_transforms_bwd", ptr(coefs), ptr(rots), ptr(transls), ptr(ts), G, K, T, B, ptr(v_out),
             ptr(v_coefs), ptr(v_rots), ptr(v_transls), ptr(v_ts), stream_ptr())
        return v_ts, v_coefs, v_rots, v_transls


def compute_transforms(ts: Tensor, coefs: Tensor, rots: Tensor, transls: Tensor) -> Tensor:
    """MotionBases.compute_transforms (params.py:142-180): ts [B] or [1,B] (float or integer frame
    coordinates), coefs [G,K] already softmaxed -> transforms [G,B,3,4] = [R | t]."""
    ts_f = ts.reshape(-1).to(torch.float32)
    return _ComputeTransforms.apply(ts_f, coefs, rots, transls)
