"""Whole-frame renderer with the call surface of the reference's ``SceneModel.render``.

``FrameRenderer.render(t, w2cs, Ks, img_wh, ...)`` takes the same arguments and returns the same
``out_dict`` keys (``img, mask, tracks_3d, depth, acc, deltaT, RTs, pred_sharp_img, exposure_imgs``)
and maintains the same densifier side channel (``_current_xys / _current_radii / _current_img_wh``)
as ``flow3d/scene_model.py:162-487``, but runs the N sub-exposures as ONE fused pass
(``scene.render_subexposures``) instead of the reference's serial loop (scene_model.py:323-385).

``CameraMotionModel`` is the host-side counterpart of ``flow3d/models/move_model.py`` (same
parameter names, so ``ckpt["move_model"]`` loads with ``load_state_dict``): the tiny pose MLP stays in
torch, the SE(3) interpolation of its two outputs runs in ``camera.interpolate_camera_deltas``.

Reference behaviours kept on purpose (each is cited):
  * 11 sub-exposures, hard-coded (scene_model.py:248); modes "mid" / "start" / "end" render one of them;
  * the camera delta moves the Gaussian centres only, not their orientation (scene_model.py:352-353);
  * the combined image is written over the LAST sub-exposure's tensor before the max / min over the
    stack are taken (scene_model.py:391-393) -> ``ref_quirk=True`` in the combine;
  * ``target_w2cs`` are used un-refined for the track channels (the refined ones are computed and
    dropped, scene_model.py:266-281);
  * deltaT = 0 in stage "first" or when int(t) is not an interior frame (move_model.py:122-133).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from .camera import interpolate_camera_deltas, subexposure_times
from .motion import compute_poses_all, compute_poses_fg, compute_transforms
from .scene import render_subexposures

NUM_SUBEXPOSURES = 11  # scene_model.py:248


class SubexposureXys:
    """Stands for ``info["means2d"]`` of ONE sub-exposure in the densifier side channel (scene_model.py:456-459).
    The fused pass holds the screen-space centres of all N sub-exposures in one non-leaf tensor [N,G,2]; the trainer
    reads ``_current_xys[ii].grad`` per sub-exposure (trainer.py:975), so ``.grad`` is that slice, [1,G,2]."""

    def __init__(self, fused: Tensor, index: int):
        self._fused, self._index = fused, index

    @property
    def grad(self) -> Optional[Tensor]:
        g = self._fused.grad
        return None if g is None else g[self._index:self._index + 1]

    @property
    def shape(self):
        return torch.Size((1,) + tuple(self._fused.shape[1:]))

    def detach(self) -> Tensor:
        return self._fused.detach()[self._index:self._index + 1]


# ------------------------------------------------------------------------------------------------ #
# camera motion / exposure model (move_model.py:66-166)
# ------------------------------------------------------------------------------------------------ #
def _posenc(x: Tensor, n_freqs: int = 5) -> Tensor:
    """[x, sin(2^k x), cos(2^k x)]_{k<n_freqs} in the reference's order (move_model.py:12-60)."""
    out = [x]
    for k in range(n_freqs):
        out += [torch.sin(x * (2.0 ** k)), torch.cos(x * (2.0 ** k))]
    return torch.cat(out, dim=-1)


def _series(x2: Tensor, kind: str, terms: int = 11) -> Tensor:
    """Taylor series A = sin x / x, B = (1 - cos x) / x^2 in x^2 (spline_utils.py:26-45)."""
    acc, denom, p = torch.zeros_like(x2), 1.0, torch.ones_like(x2)
    for i in range(terms):
        if kind == "A":
            denom *= (2 * i) * (2 * i + 1) if i > 0 else 1.0
        else:
            denom *= (2 * i + 1) * (2 * i + 2)
        acc = acc + ((-1.0) ** i) * p / denom
        p = p * x2
    return acc


def se3_log_of_pose(Rt: Tensor, eps: float = 1e-8) -> Tensor:
    """[3,4] world-to-camera -> 6-vector [w, u] (spline_utils.py:177-201, SE3_to_se3 / SO3_to_so3)."""
    R, t = Rt[:, :3], Rt[:, 3:]
    trace = R[0, 0] + R[1, 1] + R[2, 2]
    theta = torch.acos(((trace - 1) / 2).clamp(-1 + 1e-7, 1 - 1e-7)) % torch.pi
    lnR = (R - R.t()) / (2 * _series(theta * theta, "A") + 1e-8)
    w = torch.stack([lnR[2, 1], lnR[0, 2], lnR[1, 0]])
    wx = torch.zeros(3, 3, dtype=Rt.dtype, device=Rt.device)
    wx[0, 1], wx[0, 2], wx[1, 0], wx[1, 2], wx[2, 0], wx[2, 1] = -w[2], w[1], w[2], -w[0], -w[1], w[0]
    th2 = (w * w).sum()
    A, B = _series(th2, "A"), _series(th2, "B")
    invV = torch.eye(3, dtype=Rt.dtype, device=Rt.device) - 0.5 * wx + (1 - A / (2 * B)) / (th2 + eps) * (wx @ wx)
    return torch.cat([w, (invV @ t)[:, 0]])


class CameraMotionModel(nn.Module):
    """Pose MLP + learnable exposure (parameter names as in move_model.py:66-109)."""

    def __init__(self, width: int = 64, slope: float = 0.01):
        super().__init__()
        act = lambda: nn.LeakyReLU(slope)
        self.RT_main = nn.Sequential(nn.Linear(66, width), act(), nn.Linear(width, width), act(),
                                     nn.Linear(width, width), act(), nn.Linear(width, width), act(),
                                     nn.Linear(width, width))
        self.RT_head0 = nn.Sequential(nn.Linear(width, width), act(), nn.Linear(width, 6))
        self.RT_head1 = nn.Sequential(nn.Linear(width, width), act(), nn.Linear(width, 6))
        self.time_params = nn.Parameter(torch.full((1, 8), 0.5))
        for head in (self.RT_head0, self.RT_head1):  # zero-initialised heads: identity deltas at start
            nn.init.zeros_(head[-1].weight)
            nn.init.zeros_(head[-1].bias)

    def heads(self, R: Tensor, T: Tensor) -> Tuple[Tensor, Tensor]:
        x = self.RT_main(_posenc(se3_log_of_pose(torch.cat([R, T], dim=-1))[None]))
        return self.RT_head0(x), self.RT_head1(x)

    def exposure(self, t, stage: str) -> Tuple[Tensor, Tensor]:
        zero = torch.zeros_like(self.time_params[:, 0])
        if stage == "first":
            return zero, zero
        idx = int(t)
        if idx <= 0 or idx >= self.time_params.shape[-1] - 1:
            return zero, zero
        d = F.relu(self.time_params[:, idx]).clamp(0.1, 0.9)
        return -d, d

    def forward_start_end_mid(self, info: Dict, num_cameras: int = NUM_SUBEXPOSURES, mode: str = "uniform",
                              stage: str = "second"):
        """-> RTs [N,3,4], times [1,N], deltaT [1,1] (move_model.py:138-166, mode 'uniform')."""
        start6, end6 = self.heads(info["R"], info["T"])
        RTs = interpolate_camera_deltas(start6, end6, num_cameras)
        d0, d1 = self.exposure(info["timestep"], stage)
        times = subexposure_times(float(info["timestep"]), d0, d1, num_cameras)[None]
        return RTs, times, torch.abs(d1)[:, None]


# ------------------------------------------------------------------------------------------------ #
# the frame renderer
# ------------------------------------------------------------------------------------------------ #
class FrameRenderer(nn.Module):
    """Holds the raw scene parameters the way the reference's SceneModel does (fg / bg GaussianParams,
    MotionBases, Ks, w2cs; scene_model.py:14-36) and renders blurry / sharp frames."""

    def __init__(self, Ks: Tensor, w2cs: Tensor, fg: Dict[str, Tensor], rots: Tensor, transls: Tensor,
                 bg: Optional[Dict[str, Tensor]] = None, move_model: Optional[CameraMotionModel] = None):
        super().__init__()
        self.fg = nn.ParameterDict({k: nn.Parameter(v) for k, v in fg.items()})
        self.bg = nn.ParameterDict({k: nn.Parameter(v) for k, v in bg.items()}) if bg is not None else None
        self.rots, self.transls = nn.Parameter(rots), nn.Parameter(transls)
        self.register_buffer("Ks", Ks)
        self.register_buffer("w2cs", w2cs)
        self.move_model = move_model if move_model is not None else CameraMotionModel()
        self.num_frames = rots.shape[1]
        self._current_xys = self._current_radii = self._current_img_wh = None
        self._fused_xys = self._fused_radii = None
        # sync-free tile binning: one RenderCapacity per (image size, sub-exposures, Gaussians) the renderer has seen
        self.sync_free = True
        self._capacities = {}

    @classmethod
    def from_scene(cls, scene, move_model: Optional[CameraMotionModel] = None) -> "FrameRenderer":
        fg = dict(means=scene.fg_means, quats=scene.fg_quats, scales=scene.fg_scales, colors=scene.fg_colors,
                  opacities=scene.fg_opacities, motion_coefs=scene.motion_coefs)
        bg = dict(means=scene.bg_means, quats=scene.bg_quats, scales=scene.bg_scales, colors=scene.bg_colors,
                  opacities=scene.bg_opacities) if scene.num_bg else None
        return cls(scene.K, scene.w2c, fg, scene.rots, scene.transls, bg, move_model)

    # -- sizes / poses (scene_model.py:38-120) ------------------------------------------------------
    @property
    def num_fg_gaussians(self) -> int:
        return self.fg["means"].shape[0]

    @property
    def num_bg_gaussians(self) -> int:
        return 0 if self.bg is None else self.bg["means"].shape[0]

    @property
    def num_gaussians(self) -> int:
        return self.num_fg_gaussians + self.num_bg_gaussians

    def compute_transforms(self, ts: Tensor, inds: Optional[Tensor] = None) -> Tensor:
        coefs = F.softmax(self.fg["motion_coefs"], dim=-1)
        return compute_transforms(ts, coefs if inds is None else coefs[inds], self.rots, self.transls)

    def compute_poses_fg(self, ts: Tensor):
        return compute_poses_fg(self.fg["means"], self.fg["quats"], self.fg["motion_coefs"], self.rots, self.transls, ts)

    def compute_poses_all(self, ts: Tensor):
        if self.bg is None:
            return self.compute_poses_fg(ts)
        return compute_poses_all(self.fg["means"], self.fg["quats"], self.fg["motion_coefs"], self.bg["means"],
                                 self.bg["quats"], self.rots, self.transls, ts)

    # -- render -----------------------------------------------------------------------------------------
    def _capacity_for(self, key):
        if not self.sync_free:
            return None
        from .rendering import RenderCapacity
        if key not in self._capacities:
            self._capacities[key] = RenderCapacity()
        return self._capacities[key]

    def check_capacity(self):
        """After a device synchronisation: raise if any sync-free render since the last check overflowed its
        binning capacity (rendering.RenderCapacity.check)."""
        for cap in self._capacities.values():
            cap.check()

    def _subset(self, fg_only: bool, bg_only: bool):
        """Raw parameter groups of the rendered subset: (fg dict | None, bg dict | None)."""
        assert not (fg_only and bg_only)
        if fg_only:
            return self.fg, None
        if bg_only or self.num_fg_gaussians == 0:
            return None, self.bg
        return self.fg, self.bg

    def render(self, t, w2cs: Tensor, Ks: Tensor, img_wh: Tuple[int, int], target_ts: Optional[Tensor] = None,
               target_w2cs: Optional[Tensor] = None, bg_color=1.0, colors_override: Optional[Tensor] = None,
               means=None, quats=None, target_means: Optional[Tensor] = None, return_color: bool = True,
               return_depth: bool = False, return_mask: bool = False, fg_only: bool = False, bg_only: bool = False,
               filter_mask=None, epoch=1, mode: str = "mid", stage: str = "second") -> Dict[str, Tensor]:
        if filter_mask is not None or means is not None or quats is not None:
            raise NotImplementedError("filter_mask / means / quats overrides are never passed by the reference's callers")
        assert w2cs.shape[0] == 1
        dev = w2cs.device
        W, H = img_wh
        fg, bg = self._subset(fg_only, bg_only)
        cat = lambda k: torch.cat([g[k] for g in (fg, bg) if g is not None], 0)
        n_fg = 0 if fg is None else fg["means"].shape[0]
        G = n_fg + (0 if bg is None else bg["means"].shape[0])

        # per-Gaussian feature vector: rgb | mask | 3-D track targets (scene_model.py:203-289)
        feats, bgc, widths = [], [], {}
        if colors_override is None:
            colors_override = torch.sigmoid(cat("colors")) if return_color else torch.zeros(G, 0, device=dev)
        feats.append(colors_override)
        widths["img"] = colors_override.shape[-1]
        if isinstance(bg_color, float):
            bg_color = torch.full((1, widths["img"]), bg_color, device=dev)
        bgc.append(bg_color)
        if return_mask:
            # fg rows 1, bg rows 0 (scene_model.py:233-246); fg_only / bg_only render an all-ones mask, and a full render
            # of a scene WITHOUT foreground Gaussians marks none (mask_values[:0] = 1 touches nothing)
            m = torch.ones(G, 1, device=dev)
            if not fg_only and not bg_only:
                m[n_fg:] = 0.0
            feats.append(m)
            bgc.append(torch.zeros(1, 1, device=dev))
            widths["mask"] = 1
        B = 0
        if target_ts is not None:
            B = target_ts.shape[0]
            if target_means is None:
                target_means, _ = (self.compute_poses_fg if bg is None else self.compute_poses_all)(target_ts)
            if target_w2cs is not None:
                target_means = torch.einsum("bij,pbj->pbi", target_w2cs[:, :3], F.pad(target_means, (0, 1), value=1.0))
            feats.append(target_means.flatten(-2))
            bgc.append(torch.zeros(1, 3 * B, device=dev))
            widths["tracks_3d"] = 3 * B
        colors = torch.cat(feats, dim=-1)
        backgrounds = torch.cat(bgc, dim=-1)
        if return_depth:
            widths["depth"] = 1

        # sub-exposure schedule (scene_model.py:248-256, 313-321)
        RTs, times, deltaT = self.move_model.forward_start_end_mid(
            {"R": w2cs[0, :3, :3], "T": w2cs[0, :3, 3:4], "timestep": t}, num_cameras=NUM_SUBEXPOSURES, stage=stage)
        pick = {"mid": NUM_SUBEXPOSURES // 2, "start": 0, "end": NUM_SUBEXPOSURES - 1}.get(mode)
        if pick is not None:
            RTs, times = RTs[pick:pick + 1], times[:, pick:pick + 1]
        N = RTs.shape[0]

        K_rank = self.rots.shape[0]
        empty = lambda *s: torch.zeros(*s, device=dev)
        out = render_subexposures(
            fg["means"] if fg is not None else empty(0, 3), fg["quats"] if fg is not None else empty(0, 4),
            fg["motion_coefs"] if fg is not None else empty(0, K_rank),
            bg["means"] if bg is not None else None, bg["quats"] if bg is not None else None,
            self.rots, self.transls, times[0], RTs, torch.exp(cat("scales")), torch.sigmoid(cat("opacities")), colors,
            w2cs, Ks, W, H, backgrounds=backgrounds, render_mode="RGB+ED" if return_depth else "RGB",
            combine=True, ref_quirk=True, capacity=self._capacity_for((W, H, N, G)))

        # densifier side channel (scene_model.py:456-461).  The reference stashes one means2d / radii tensor per
        # sub-exposure and Trainer._prepare_control_step (trainer.py:967-989) indexes them: radii[ii] is [1,G],
        # xys[ii].grad is [1,G,2].  Here they are per-sub-exposure views of the fused [N,G,.] tensors, which
        # control.accumulate_densify_stats consumes whole (``_fused_xys`` / ``_fused_radii``).
        if out["means2d"].requires_grad:
            out["means2d"].retain_grad()
            self._fused_xys, self._fused_radii = out["means2d"], out["radii"]
            self._current_xys = [SubexposureXys(out["means2d"], ii) for ii in range(N)]
            self._current_radii = [out["radii"][ii:ii + 1] for ii in range(N)]
            self._current_img_wh = img_wh

        res: Dict[str, Tensor] = {}
        pieces = torch.split(out["img"], list(widths.values()), dim=-1)
        for (name, _), x in zip(widths.items(), pieces):
            res[name] = x.reshape(1, H, W, B, 3) if name == "tracks_3d" else x
        res["acc"] = out["acc"]
        res["deltaT"] = deltaT[None]
        res["RTs"] = RTs
        res["pred_sharp_img"] = out["pred_sharp_img"]
        # the reference's stack holds the combined image in its last slot (in-place alias, :391-393 + :482-486)
        res["exposure_imgs"] = torch.cat([out["exposure_imgs"][:-1], out["img"][None]], 0) if N > 1 else out["img"][None]
        return res
