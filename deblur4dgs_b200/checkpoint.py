"""Reference checkpoint / state-dict interchange.

The reference saves ``{"model": SceneModel.state_dict(), "move_model": ..., "optimizers": ...,
"global_step", "epoch"}`` (flow3d/trainer.py:126-140) and restores the scene with
``SceneModel.init_from_state_dict`` (flow3d/scene_model.py:145-160), whose key layout is

    fg.params.{means,quats,scales,colors,opacities,motion_coefs}   (flow3d/params.py:52-64)
    bg.params.{means,quats,scales,colors,opacities}                (optional)
    motion_bases.params.{rots,transls}                             (flow3d/params.py:134-139)
    Ks [T,3,3], w2cs [T,4,4]

``scene_from_state_dict`` reads exactly that layout into the ``synthetic.Scene`` container the
benchmark and the fused render path use, so trained Deblur4DGS scenes can be replayed through
``scene.render_subexposures`` (``bench.py --checkpoint``); ``scene_to_state_dict`` writes it back.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from .synthetic import Scene

_FG = ["means", "quats", "scales", "colors", "opacities", "motion_coefs"]
_BG = ["means", "quats", "scales", "colors", "opacities"]


def scene_from_state_dict(state_dict: Dict[str, Tensor], width: int, height: int, frame: int = 0, N: int = 9,
                          delta_t: float = 0.5, prefix: str = "", times: Optional[Tensor] = None,
                          RTs: Optional[Tensor] = None) -> Scene:
    """Build a Scene for training frame ``frame`` from a reference ``SceneModel`` state dict.

    ``times`` / ``RTs`` are the N sub-exposure timestamps / camera deltas ``MoveModel`` would produce
    (move_model.py:138-166); they default to ``linspace(frame - delta_t, frame + delta_t, N)`` and identity."""
    sd = state_dict
    req = [f"{prefix}fg.params.{k}" for k in _FG] + [f"{prefix}motion_bases.params.rots",
                                                    f"{prefix}motion_bases.params.transls", f"{prefix}Ks", f"{prefix}w2cs"]
    missing = [k for k in req if k not in sd]
    if missing:
        raise KeyError(f"not a Deblur4DGS SceneModel state dict, missing {missing}")
    fg = {k: sd[f"{prefix}fg.params.{k}"].detach().float().contiguous() for k in _FG}
    has_bg = any(k.startswith(f"{prefix}bg.") for k in sd)
    if has_bg:
        bg = {k: sd[f"{prefix}bg.params.{k}"].detach().float().contiguous() for k in _BG}
    else:
        z = fg["means"].new_zeros
        bg = {"means": z((0, 3)), "quats": z((0, 4)), "scales": z((0, 3)), "colors": z((0, 3)), "opacities": z((0,))}
    rots = sd[f"{prefix}motion_bases.params.rots"].detach().float().contiguous()
    transls = sd[f"{prefix}motion_bases.params.transls"].detach().float().contiguous()
    Ks, w2cs = sd[f"{prefix}Ks"].float(), sd[f"{prefix}w2cs"].float()
    T = rots.shape[1]
    if not (0 <= frame < w2cs.shape[0]):
        raise IndexError(f"frame {frame} outside the checkpoint's {w2cs.shape[0]} cameras")
    if times is None:
        times = torch.linspace(frame - delta_t, frame + delta_t, N) if N > 1 else torch.tensor([float(frame)])
    if RTs is None:
        RTs = torch.eye(4)[:3][None].repeat(times.shape[0], 1, 1)
    G = fg["means"].shape[0] + bg["means"].shape[0]
    assert fg["motion_coefs"].shape[1] == rots.shape[0] and transls.shape[:2] == rots.shape[:2] and T >= 1
    return Scene(fg_means=fg["means"], fg_quats=fg["quats"], fg_scales=fg["scales"], fg_colors=fg["colors"],
                 fg_opacities=fg["opacities"], motion_coefs=fg["motion_coefs"], bg_means=bg["means"],
                 bg_quats=bg["quats"], bg_scales=bg["scales"], bg_colors=bg["colors"], bg_opacities=bg["opacities"],
                 rots=rots, transls=transls, w2c=w2cs[frame:frame + 1].contiguous(), K=Ks[frame:frame + 1].contiguous(),
                 times=times.float().contiguous(), RTs=RTs.float().contiguous(),
                 extra_channels=torch.zeros(G, 0), width=width, height=height)


def load_checkpoint(path: str, width: int, height: int, **kw) -> Tuple[Scene, dict]:
    """Read a checkpoint written by Trainer.save_checkpoint (trainer.py:126-140)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    sd = ckpt["model"] if "model" in ckpt else ckpt
    extras = {k: ckpt[k] for k in ("global_step", "epoch", "move_model") if isinstance(ckpt, dict) and k in ckpt}
    return scene_from_state_dict(sd, width, height, **kw), extras


def scene_to_state_dict(scene: Scene, Ks: Optional[Tensor] = None, w2cs: Optional[Tensor] = None,
                        prefix: str = "") -> Dict[str, Tensor]:
    """The inverse mapping (reference key layout), e.g. to hand a synthetic scene to the reference trainer."""
    sd = {f"{prefix}fg.params.{k}": getattr(scene, "motion_coefs" if k == "motion_coefs" else f"fg_{k}") for k in _FG}
    if scene.num_bg:
        sd.update({f"{prefix}bg.params.{k}": getattr(scene, f"bg_{k}") for k in _BG})
    sd[f"{prefix}motion_bases.params.rots"] = scene.rots
    sd[f"{prefix}motion_bases.params.transls"] = scene.transls
    sd[f"{prefix}Ks"] = scene.K if Ks is None else Ks
    sd[f"{prefix}w2cs"] = scene.w2c if w2cs is None else w2cs
    return sd
