"""Camera sub-exposure pose interpolation on top of libd4gs.so.

``interpolate_camera_deltas(start6, end6, N)`` replaces the pose part of
``MoveModel.forward_start_end_mid`` (flow3d/models/move_model.py:143-147):
``pp.se3(RT_start).Exp()``, ``pp.se3(RT_end).Exp()``, ``_interpolate`` (spline_utils.py:371-408),
``.Log()`` and ``postprocessPose``/``se3_to_SE3`` (spline_utils.py:204-215) -- ~100 tiny torch/pypose
launches in the reference, one kernel here, with gradients to both 6-vectors (the MLP heads).
The MLP itself (11-row GEMMs) stays the reference's torch code.
"""
from __future__ import annotations

import torch
from torch import Tensor

from ._cabi import D4Error, call, check_tensors, ptr, stream_ptr


class _CameraInterp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, start6, end6, N):
        check_tensors(start6, end6, what="camera ops")
        s = start6.reshape(6).float().contiguous()
        e = end6.reshape(6).float().contiguous()
        RTs = torch.empty((N, 3, 4), dtype=torch.float32, device=s.device)
        call("d4_camera_interp_fwd", ptr(s), ptr(e), N, ptr(RTs), stream_ptr())
        ctx.save_for_backward(s, e)
        ctx.N = N
        ctx.shapes = (start6.shape, end6.shape)
        return RTs

    @staticmethod
    def backward(ctx, v_RTs):
        s, e = ctx.saved_tensors
        vs, ve = torch.zeros_like(s), torch.zeros_like(e)
        v = v_RTs.float().contiguous()  # bound to a name: must outlive the launch
        call("d4_camera_interp_bwd", ptr(s), ptr(e), ctx.N, ptr(v), ptr(vs), ptr(ve), stream_ptr())
        return vs.reshape(ctx.shapes[0]), ve.reshape(ctx.shapes[1]), None


def interpolate_camera_deltas(start6: Tensor, end6: Tensor, num_cameras: int) -> Tensor:
    """start6 / end6: [6] or [1,6] se(3) vectors (pypose order [rho, phi]) -> RTs [N,3,4]."""
    return _CameraInterp.apply(start6, end6, int(num_cameras))


def subexposure_times(t: float, delta0: Tensor, delta1: Tensor, num_cameras: int) -> Tensor:
    """move_model.py:150-156 (plain torch: N elements)."""
    w = (torch.arange(num_cameras, device=delta0.device) / (num_cameras - 1)).to(delta0.dtype)
    return ((delta0 + t) * (1.0 - w) + (delta1 + t) * w).reshape(-1)
