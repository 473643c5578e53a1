"""Seeded synthetic scenes for parity tests and ``bench.py``.

The generator runs on the CPU with ``torch.Generator().manual_seed(seed)`` so
that the oracle, the CUDA path and every rank see identical bits.  The layout
of a scene mirrors what the reference keeps in ``SceneModel``
(``flow3d/scene_model.py:14-36``): raw (pre-activation) foreground and
background Gaussian parameters (``flow3d/params.py:10-43``), the motion bases
(``flow3d/params.py:121-139``), one camera and the N sub-exposure timestamps /
camera deltas that ``MoveModel.forward_start_end_mid`` would produce
(``flow3d/models/move_model.py:138-166``).
"""
from __future__ import annotations

import dataclasses
import math

import torch


@dataclasses.dataclass
class Scene:
    # raw parameters (pre-activation), foreground first then background
    fg_means: torch.Tensor  # [Gf,3]
    fg_quats: torch.Tensor  # [Gf,4] wxyz raw
    fg_scales: torch.Tensor  # [Gf,3] log
    fg_colors: torch.Tensor  # [Gf,3] logit
    fg_opacities: torch.Tensor  # [Gf] logit
    motion_coefs: torch.Tensor  # [Gf,K] raw (softmax inside)
    bg_means: torch.Tensor
    bg_quats: torch.Tensor
    bg_scales: torch.Tensor
    bg_colors: torch.Tensor
    bg_opacities: torch.Tensor
    rots: torch.Tensor  # [K,T,6]
    transls: torch.Tensor  # [K,T,3]
    w2c: torch.Tensor  # [1,4,4]
    K: torch.Tensor  # [1,3,3]
    times: torch.Tensor  # [N] sub-exposure timestamps
    RTs: torch.Tensor  # [N,3,4] camera sub-exposure deltas
    extra_channels: torch.Tensor  # [G, D0-4] track channels (may have 0 columns)
    width: int
    height: int

    @property
    def num_fg(self):
        return self.fg_means.shape[0]

    @property
    def num_bg(self):
        return self.bg_means.shape[0]

    @property
    def G(self):
        return self.num_fg + self.num_bg

    @property
    def N(self):
        return self.times.shape[0]

    def to(self, device):
        kw = {}
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            kw[f.name] = v.to(device) if isinstance(v, torch.Tensor) else v
        return Scene(**kw)

    def tensors(self):
        return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)
                if isinstance(getattr(self, f.name), torch.Tensor)}

    # activations exactly as flow3d/params.py:39-43,70-84 and the fg|bg concat of
    # flow3d/scene_model.py:122-143
    def scales_all(self):
        return torch.exp(torch.cat([self.fg_scales, self.bg_scales], 0))

    def opacities_all(self):
        return torch.sigmoid(torch.cat([self.fg_opacities, self.bg_opacities], 0))

    def colors_all(self, d0: int):
        """[G, d0] feature vector the reference assembles in render()
        (scene_model.py:205-289): rgb, fg-mask, track channels."""
        rgb = torch.sigmoid(torch.cat([self.fg_colors, self.bg_colors], 0))
        if d0 == 3:
            return rgb
        mask = torch.zeros(self.G, 1, dtype=rgb.dtype, device=rgb.device)
        mask[: self.num_fg] = 1.0
        out = torch.cat([rgb, mask, self.extra_channels], dim=-1)
        assert out.shape[1] >= d0
        return out[:, :d0].contiguous()


CONFIGS = {
    # name: (G, W, H, K, N, seed)   -- BASELINE.json configs / SURVEY 8(d)
    "c1": (1_000, 512, 288, 6, 1, 1),
    "c2": (100_000, 512, 288, 6, 5, 2),
    "c3": (300_000, 1280, 720, 10, 9, 3),
    "c5": (1_000_000, 1280, 720, 16, 13, 5),
}


def make_scene(G: int, width: int, height: int, K: int, N: int, seed: int, T: int = 8,
               fg_fraction: float = 0.3, d_extra: int = 12, scale_mult: float = 1.0,
               t_center: float = 3.0, delta_t: float = 0.5, cam_motion: float = 0.01) -> Scene:
    g = torch.Generator().manual_seed(seed)
    U = lambda *s: torch.rand(*s, generator=g)
    Nrm = lambda *s: torch.randn(*s, generator=g)
    f = 0.8 * width
    z = 2.0 + 8.0 * U(G)
    u, v = U(G), U(G)
    x = (u - 0.5) * 1.2 * z * width / f
    y = (v - 0.5) * 1.2 * z * height / f
    means = torch.stack([x, y, z], -1)
    lo, hi = math.log(0.005 * scale_mult), math.log(0.03 * scale_mult)
    scales = lo + (hi - lo) * U(G, 3)
    quats = Nrm(G, 4)
    op = 0.1 + 0.85 * U(G)
    opac = torch.log(op / (1 - op))
    colors = Nrm(G, 3)
    Gf = int(round(G * fg_fraction))
    coefs = 2.0 * Nrm(Gf, K)
    base6 = torch.tensor([1.0, 0, 0, 0, 1.0, 0])
    rots = base6 + 0.05 * torch.cumsum(Nrm(K, T, 6), dim=1) / math.sqrt(T)
    transls = 0.05 * torch.cumsum(Nrm(K, T, 3), dim=1) / math.sqrt(T)
    extra = Nrm(G, d_extra)
    if N > 1:
        times = torch.linspace(t_center - delta_t, t_center + delta_t, N)
    else:
        times = torch.tensor([t_center - 0.37 * delta_t])
    # camera sub-exposure deltas: small rigid motions interpolated start->end
    # (the a7 interpolation itself is exercised by its own op; here plain
    # Rodrigues of a linearly interpolated tangent keeps the generator simple)
    xi0, xi1 = cam_motion * Nrm(6), cam_motion * Nrm(6)
    RTs = []
    for i in range(N):
        a = 0.5 if N == 1 else i / (N - 1)
        xi = (1 - a) * xi0 + a * xi1
        w = xi[:3]
        th = w.norm()
        Kx = torch.tensor([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        R = torch.eye(3) + (torch.sin(th) / th) * Kx + ((1 - torch.cos(th)) / (th * th)) * (Kx @ Kx)
        RTs.append(torch.cat([R, xi[3:, None]], dim=1))
    RTs = torch.stack(RTs).float()
    w2c = torch.eye(4)[None].clone()
    Kmat = torch.tensor([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1.0]])[None]
    return Scene(fg_means=means[:Gf].contiguous(), fg_quats=quats[:Gf].contiguous(),
                 fg_scales=scales[:Gf].contiguous(), fg_colors=colors[:Gf].contiguous(),
                 fg_opacities=opac[:Gf].contiguous(), motion_coefs=coefs,
                 bg_means=means[Gf:].contiguous(), bg_quats=quats[Gf:].contiguous(),
                 bg_scales=scales[Gf:].contiguous(), bg_colors=colors[Gf:].contiguous(),
                 bg_opacities=opac[Gf:].contiguous(), rots=rots.contiguous(), transls=transls.contiguous(),
                 w2c=w2c, K=Kmat, times=times, RTs=RTs.contiguous(), extra_channels=extra,
                 width=width, height=height)


def make_config(name: str, **overrides) -> Scene:
    G, W, H, K, N, seed = CONFIGS[name]
    kw = dict(G=G, width=W, height=H, K=K, N=N, seed=seed)
    kw.update(overrides)
    return make_scene(**kw)


def cotangents(shape_colors, shape_alphas, seed: int = 1234):
    """Fixed random cotangents for backward parity."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape_colors, generator=g), torch.randn(*shape_alphas, generator=g)
