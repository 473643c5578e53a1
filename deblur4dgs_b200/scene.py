"""Fused per-frame render: all N sub-exposures of ``SceneModel.render`` in one pass.

``render_subexposures`` replaces the serial Python loop of the reference
(flow3d/scene_model.py:323-385: per sub-exposure deformation -> camera delta ->
``rasterization``) and the N-way combine after it (scene_model.py:386-397) by

    1 deformation launch  (all N timestamps; motion.deform_subexposures)
    1 projection launch   (N x G Gaussians, "C = N cameras" with per-camera centres)
    1 binning + sort      (keys carry the sub-exposure index in the camera bits)
    1 blend launch        (N x tiles CTAs)
    1 combine launch      (mean / max / min over N, each input read once)

with ONE device->host sync (the intersection count) per blurry frame instead of
N, and identical per-sub-exposure results: every stage is the same arithmetic
``rasterization`` runs for a single sub-exposure.

Outputs keep what the reference's callers read: the combined image / alpha, the
per-sub-exposure stack (``exposure_imgs``, trainer.py:601-617), the sharp middle
image, and per-sub-exposure ``means2d`` / ``radii`` for the densifier
(scene_model.py:456-461, trainer.py:953-990).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

from ._cabi import call, ptr, stream_ptr
from .motion import deform_subexposures
from .rendering import rasterization


class _Combine(torch.autograd.Function):
    """scene_model.py:386-397 -- mean over N, max on the mask channel, min on the depth channel."""

    @staticmethod
    def forward(ctx, imgs, alphas, max_ch, min_ch, ref_quirk):
        imgs, alphas = imgs.contiguous(), alphas.contiguous()
        N = imgs.shape[0]
        D = imgs.shape[-1]
        P = imgs[0].numel() // D
        out_img = torch.empty_like(imgs[0])
        out_alpha = torch.empty_like(alphas[0])
        # winner maps of the max / min channel: the backward routes from them, the stack is not kept
        arg_max = torch.empty((P,), dtype=torch.uint8, device=imgs.device) if 0 <= max_ch < D else None
        arg_min = torch.empty((P,), dtype=torch.uint8, device=imgs.device) if 0 <= min_ch < D else None
        call("d4_combine_fwd", ptr(imgs), ptr(alphas), N, P, D, max_ch, min_ch, int(ref_quirk), ptr(out_img),
             ptr(out_alpha), ptr(arg_max), ptr(arg_min), stream_ptr())
        ctx.save_for_backward(arg_max, arg_min)
        ctx.cfg = (N, P, D, max_ch, min_ch, imgs.shape, alphas.shape)
        return out_img, out_alpha

    @staticmethod
    def backward(ctx, v_img, v_alpha):
        arg_max, arg_min = ctx.saved_tensors
        N, P, D, max_ch, min_ch, ishape, ashape = ctx.cfg
        dev = v_img.device if v_img is not None else v_alpha.device
        v_img = v_img.contiguous() if v_img is not None else torch.zeros(ishape[1:], device=dev)
        v_alpha = v_alpha.contiguous() if v_alpha is not None else torch.zeros(ashape[1:], device=dev)
        v_imgs = torch.empty(ishape, dtype=torch.float32, device=dev)
        v_alphas = torch.empty(ashape, dtype=torch.float32, device=dev)
        call("d4_combine_bwd", ptr(arg_max), ptr(arg_min), N, P, D, max_ch, min_ch, ptr(v_img), ptr(v_alpha),
             ptr(v_imgs), ptr(v_alphas), stream_ptr())
        return v_imgs, v_alphas, None, None, None


def combine_subexposures(imgs: Tensor, alphas: Tensor, max_ch: int = -1, min_ch: int = -1, ref_quirk: bool = True):
    """imgs [N,...,D], alphas [N,...,1] -> (combined image [...,D], mean alpha [...,1])."""
    return _Combine.apply(imgs, alphas, int(max_ch), int(min_ch), bool(ref_quirk))


def render_subexposures(
    fg_means: Tensor, fg_quats: Tensor, motion_coefs: Tensor,
    bg_means: Optional[Tensor], bg_quats: Optional[Tensor],
    rots: Tensor, transls: Tensor,
    times: Tensor,  # [N]
    RTs: Optional[Tensor],  # [N,3,4]
    scales: Tensor,  # [G,3] activated (exp), fg first
    opacities: Tensor,  # [G] activated (sigmoid)
    colors: Tensor,  # [G,D0] per-Gaussian feature vector (rgb | mask | tracks), scene_model.py:205-289
    w2c: Tensor,  # [1,4,4]
    K: Tensor,  # [1,3,3]
    width: int, height: int,
    backgrounds: Optional[Tensor] = None,  # [1,D0]
    render_mode: str = "RGB+ED",
    combine: bool = True,
    ref_quirk: bool = True,
    capacity=None,  # rendering.RenderCapacity: sync-free tile binning (no device -> host read-back in the step)
    row_windows=None,  # (row0 i32 [C], window_height): camera c renders that row band only (parallel.py)
    camera_of=None,  # LongTensor [C]: camera c shows sub-exposure camera_of[c] of times / RTs (several row bands of one)
) -> Dict[str, Tensor]:
    means, quats = deform_subexposures(fg_means, fg_quats, motion_coefs, bg_means, bg_quats, rots, transls, times, RTs)
    if camera_of is not None:  # the deformation runs once per DISTINCT sub-exposure
        means, quats = means.index_select(0, camera_of), quats.index_select(0, camera_of)
    N = means.shape[0]
    bg = None if backgrounds is None else backgrounds.expand(N, -1)
    imgs, alphas, meta = rasterization(means=means, quats=quats, scales=scales, opacities=opacities, colors=colors,
                                       backgrounds=bg, viewmats=w2c, Ks=K, width=width, height=height, packed=False,
                                       render_mode=render_mode, capacity=capacity, row_windows=row_windows)
    out = {"exposure_imgs": imgs[:, None], "exposure_alphas": alphas[:, None], "means2d": meta["means2d"],
           "radii": meta["radii"], "meta": meta, "means": means, "quats": quats,
           "pred_sharp_img": imgs[N // 2][None, ..., 0:3]}
    if combine:
        D = imgs.shape[-1]
        if N > 1:
            img, alpha = combine_subexposures(imgs, alphas, 3 if D > 3 else -1, 16 if D > 16 else -1, ref_quirk)
            out["img"], out["acc"] = img[None], alpha[None]
        else:
            out["img"], out["acc"] = imgs, alphas
    return out


class _Assemble(torch.autograd.Function):
    """Row f1: activations + fg|bg concat + [rgb | mask | extra] feature vector in one pass."""

    @staticmethod
    def forward(ctx, fg_scales, bg_scales, fg_opac, bg_opac, fg_colors, bg_colors, extra, with_mask):
        c = lambda t: None if t is None else t.float().contiguous()
        fg_scales, bg_scales, fg_opac, bg_opac, fg_colors, bg_colors, extra = map(
            c, (fg_scales, bg_scales, fg_opac, bg_opac, fg_colors, bg_colors, extra))
        Gf = fg_scales.shape[0]
        Gb = 0 if bg_scales is None else bg_scales.shape[0]
        E = 0 if extra is None else extra.shape[1]
        G, dev = Gf + Gb, fg_scales.device
        D0 = 3 + int(with_mask) + E
        scales = torch.empty((G, 3), dtype=torch.float32, device=dev)
        opac = torch.empty((G,), dtype=torch.float32, device=dev)
        colors = torch.empty((G, D0), dtype=torch.float32, device=dev)
        call("d4_assemble_fwd", ptr(fg_scales), ptr(bg_scales), ptr(fg_opac), ptr(bg_opac), ptr(fg_colors),
             ptr(bg_colors), ptr(extra) if E else None, Gf, Gb, E, int(with_mask), ptr(scales), ptr(opac), ptr(colors),
             stream_ptr())
        ctx.save_for_backward(scales, opac, colors)
        ctx.cfg = (Gf, Gb, E, int(with_mask), extra is not None and E > 0)
        return scales, opac, colors

    @staticmethod
    def backward(ctx, v_scales, v_opac, v_colors):
        scales, opac, colors = ctx.saved_tensors
        Gf, Gb, E, with_mask, has_extra = ctx.cfg
        dev = scales.device
        z = lambda t, ref: torch.zeros_like(ref) if t is None else t.float().contiguous()
        v_scales, v_opac, v_colors = z(v_scales, scales), z(v_opac, opac), z(v_colors, colors)
        mk = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        vfs, vfo, vfc = mk(Gf, 3), mk(Gf), mk(Gf, 3)
        vbs, vbo, vbc = (mk(Gb, 3), mk(Gb), mk(Gb, 3)) if Gb else (None, None, None)
        vex = mk(Gf + Gb, E) if has_extra else None
        call("d4_assemble_bwd", ptr(scales), ptr(opac), ptr(colors), ptr(v_scales), ptr(v_opac), ptr(v_colors), Gf, Gb,
             E, with_mask, ptr(vfs), ptr(vbs), ptr(vfo), ptr(vbo), ptr(vfc), ptr(vbc), ptr(vex), stream_ptr())
        return vfs, vbs, vfo, vbo, vfc, vbc, vex, None


def assemble_gaussians(fg_scales, bg_scales, fg_opacities, bg_opacities, fg_colors, bg_colors,
                       extra: Optional[Tensor] = None, with_mask: bool = True):
    """Raw (pre-activation) fg / bg parameters -> (scales [G,3], opacities [G], colors [G, 3(+1)+E]) as
    SceneModel.render assembles them (params.py:70-84, scene_model.py:122-143, 205-289).  bg_* may be None."""
    return _Assemble.apply(fg_scales, bg_scales, fg_opacities, bg_opacities, fg_colors, bg_colors, extra, bool(with_mask))
