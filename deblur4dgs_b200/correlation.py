"""PWC-Net cost-volume layer on top of libd4gs.so.

Drop-in for ``flow3d/models/external/pwcnet/correlation/correlation.py``: ``FunctionCorrelation(tenFirst, tenSecond)``
and ``ModuleCorrelation`` keep the reference's names and call signature (pwcnet.py:179, 187), so
``flow3d/models/pwcnet.py`` can import this module in place of the CuPy one.  tenFirst / tenSecond
[B, C, H, W] float32 contiguous -> [B, 81, H, W]: channel ``(dy + 4) * 9 + (dx + 4)`` holds the mean over C of
``tenFirst[:, :, y, x] * tenSecond[:, :, y + dy, x + dx]`` (zero outside the image).

Forward only.  The reference evaluates PWC-Net under ``torch.no_grad()`` (loss_utils.py:171-172), so its backward
kernels (correlation.py:105-233) never run; asking for a gradient here raises instead of silently returning zeros.
There is no CPU path.
"""
from __future__ import annotations

import torch
from torch import Tensor

from ._cabi import D4Error, call, check_tensors, ptr, stream_ptr


def FunctionCorrelation(tenFirst: Tensor, tenSecond: Tensor) -> Tensor:
    check_tensors(tenFirst, tenSecond, what="correlation")
    if torch.is_grad_enabled() and (tenFirst.requires_grad or tenSecond.requires_grad):
        raise D4Error("correlation is forward-only: the reference runs PWC-Net under torch.no_grad() "
                      "(flow3d/loss_utils.py:171-172)")
    assert tenFirst.shape == tenSecond.shape and tenFirst.dim() == 4
    first, second = tenFirst.float().contiguous(), tenSecond.float().contiguous()
    B, C, H, W = first.shape
    out = torch.empty((B, 81, H, W), dtype=torch.float32, device=first.device)
    call("d4_correlation_fwd", ptr(first), ptr(second), B, C, H, W, ptr(out), stream_ptr())
    return out


class ModuleCorrelation(torch.nn.Module):
    def forward(self, tenFirst: Tensor, tenSecond: Tensor) -> Tensor:
        return FunctionCorrelation(tenFirst, tenSecond)
