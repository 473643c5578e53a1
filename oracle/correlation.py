"""oracle/correlation.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (numpy) restatement of the PWC-Net cost-volume layer the reference ships as CuPy RawKernel source strings:
``kernel_Correlation_rearrange`` (flow3d/models/external/pwcnet/correlation/correlation.py:8-33: NCHW -> zero-padded
(+4 on every side) NHWC) and ``kernel_Correlation_updateOutput`` (:35-103: 81 displacements (dx, dy) in [-4, 4]^2,
``top_channel = (dy + 4) * 9 + (dx + 4)``, dot product over the C channels of first[y, x] and second[y + dy, x + dx],
divided by C), launched from ``_FunctionCorrelation.forward`` (:281-331), and of the two gradient kernels
``kernel_Correlation_updateGradFirst`` (:105-167) / ``kernel_Correlation_updateGradSecond`` (:169-233) launched per
sample from ``backward`` (:336-385).  SURVEY.md row f4.

Unlike the rasterizer, this source IS under /root/reference, so the restatement follows it line by line, including its
summation order: lane ``ch_off`` of the 32-thread block accumulates channels ch_off, ch_off + 32, ... with fused
multiply-adds (nvrtc contracts ``sum += a * b``), lane 0 then adds the 32 partial sums in lane order and divides by C.
The kernels cannot be executed here (CuPy and a GPU are both absent): parity is pinned to the reference SOURCE, not
to its output.  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def rearrange(x: np.ndarray) -> np.ndarray:
    """correlation.py:8-33 -- [B, C, H, W] -> [B, H + 8, W + 8, C], zero border of 4."""
    B, C, H, W = x.shape
    out = np.zeros((B, H + 8, W + 8, C), np.float32)
    out[:, 4:4 + H, 4:4 + W, :] = np.transpose(x, (0, 2, 3, 1))
    return out


def _fma32(a, b, c):
    # fp32 fused multiply-add: the product of two fp32 values is exact in fp64; one rounding of the sum to fp32
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def correlation(first: np.ndarray, second: np.ndarray) -> np.ndarray:
    """correlation.py:35-103 -- returns top [B, 81, H, W] float32."""
    first, second = np.ascontiguousarray(first, np.float32), np.ascontiguousarray(second, np.float32)
    B, C, H, W = first.shape
    rbot0, rbot1 = rearrange(first), rearrange(second)
    top = np.zeros((B, 81, H, W), np.float32)
    patch = rbot0[:, 4:4 + H, 4:4 + W, :]  # patch_data: the C channels at (y1, x1) = (y + 4, x + 4)
    for tc in range(81):
        s2o, s2p = tc % 9 - 4, tc // 9 - 4
        nb = rbot1[:, 4 + s2p:4 + s2p + H, 4 + s2o:4 + s2o + W, :]
        lanes = np.zeros((B, H, W, 32), np.float32)  # __shared__ float sum[32]
        for ch0 in range(0, C, 32):  # for (ch = ch_off; ch < C; ch += 32)
            n = min(32, C - ch0)
            lanes[..., :n] = _fma32(patch[..., ch0:ch0 + n], nb[..., ch0:ch0 + n], lanes[..., :n])
        total = np.zeros((B, H, W), np.float32)
        for idx in range(32):  # lane 0: total_sum += sum[idx]
            total = (total + lanes[..., idx]).astype(np.float32)
        top[:, tc] = (total / np.float32(C)).astype(np.float32)
    return top


def correlation_backward(first: np.ndarray, second: np.ndarray, grad_out: np.ndarray):
    """correlation.py:105-233 -- (gradFirst, gradSecond), both [B, C, H, W] float32.

    One thread of the reference owns (n = channel, l = x + 4, m = y + 4); with stride 1 and kernel size 1 its
    ``xmin == xmax == l - 4`` (resp. ``l - 4 - s2o``), so the two inner loops visit one gradOutput element.  Sums run
    over p (dy) outer, o (dx) inner with fused multiply-adds; one division by C (``sumelems``) at the end."""
    first, second = np.ascontiguousarray(first, np.float32), np.ascontiguousarray(second, np.float32)
    grad_out = np.ascontiguousarray(grad_out, np.float32)
    B, C, H, W = first.shape
    rbot0, rbot1 = rearrange(first), rearrange(second)  # [B, H + 8, W + 8, C]
    s1 = np.zeros((B, H, W, C), np.float32)
    s2 = np.zeros((B, H, W, C), np.float32)
    ys, xs = np.arange(H)[:, None], np.arange(W)[None, :]
    for p in range(-4, 5):
        for o in range(-4, 5):
            op = (p + 4) * 9 + (o + 4)
            # updateGradFirst: bot1tmp = rbot1[m + p, l + o, n] (zero padding), gradOutput[op, y, x]
            bot1 = rbot1[:, 4 + p:4 + p + H, 4 + o:4 + o + W, :]
            s1 = _fma32(grad_out[:, op, :, :, None], bot1, s1)
            # updateGradSecond: only where (y - p, x - o) lies inside gradOutput; bot0tmp = rbot0[m - p, l - o, n]
            yy, xx = ys - p, xs - o
            ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            yc, xc = np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)
            g = grad_out[:, op][:, yc, xc]                      # [B, H, W]
            bot0 = rbot0[:, 4 + yc, 4 + xc, :]                  # [B, H, W, C]
            upd = _fma32(g[..., None], bot0, s2)
            s2 = np.where(ok[None, :, :, None], upd, s2)
    gf = (s1 / np.float32(C)).astype(np.float32)
    gs = (s2 / np.float32(C)).astype(np.float32)
    return np.transpose(gf, (0, 3, 1, 2)).copy(), np.transpose(gs, (0, 3, 1, 2)).copy()
