"""PWC-Net cost-volume layer on top of libd4gs.so.

Drop-in for ``flow3d/models/external/pwcnet/correlation/correlation.py``: ``FunctionCorrelation(tenFirst, tenSecond)``
and ``ModuleCorrelation`` keep the reference's names and call signature (pwcnet.py:179, 187), so
``flow3d/models/pwcnet.py`` can import this module in place of the CuPy one.  tenFirst / tenSecond
[B, C, H, W] float32 contiguous -> [B, 81, H, W]: channel ``(dy + 4) * 9 + (dx + 4)`` holds the mean over C of
``tenFirst[:, :, y, x] * tenSecond[:, :, y + dy, x + dx]`` (zero outside the image).

Differentiable like the reference's ``_FunctionCorrelation`` (correlation.py:281-385): the backward runs
``d4_correlation_bwd`` (one launch for the batch; the reference launches two kernels per sample).  The reference's own
caller evaluates PWC-Net under ``torch.no_grad()`` (loss_utils.py:171-172) and never reaches it.  There is no CPU path.
"""
from __future__ import annotations

import torch
from torch import Tensor

from ._cabi import call, check_tensors, ptr, stream_ptr


class _Correlation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, first, second):
        B, C, H, W = first.shape
        out = torch.empty((B, 81, H, W), dtype=torch.float32, device=first.device)
        call("d4_correlation_fwd", ptr(first), ptr(second), B, C, H, W, ptr(out), stream_ptr())
        ctx.save_for_backward(first, second)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        first, second = ctx.saved_tensors
        B, C, H, W = first.shape
        grad_out = grad_out.float().contiguous()
        g1 = torch.empty_like(first) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(second) if ctx.needs_input_grad[1] else None
        call("d4_correlation_bwd", ptr(first), ptr(second), ptr(grad_out), B, C, H, W, ptr(g1), ptr(g2), stream_ptr())
        return g1, g2


def FunctionCorrelation(tenFirst: Tensor, tenSecond: Tensor) -> Tensor:
    check_tensors(tenFirst, tenSecond, what="correlation")
    assert tenFirst.shape == tenSecond.shape and tenFirst.dim() == 4
    return _Correlation.apply(tenFirst.float().contiguous(), tenSecond.float().contiguous())


class ModuleCorrelation(torch.nn.Module):
    def forward(self, tenFirst: Tensor, tenSecond: Tensor) -> Tensor:
        return FunctionCorrelation(tenFirst, tenSecond)
