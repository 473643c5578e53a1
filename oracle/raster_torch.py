"""oracle/raster_torch.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Dense, differentiable pure-torch restatement of gsplat-1.1.1
``rasterization(packed=False)``, written
independently of ``raster_oracle.c`` and used ONLY to pin that C oracle:
forward against forward, and the C oracle's hand-derived backward against
torch autograd of this function.  O(pixels x Gaussians) memory: small cases
only.  PARITY UNPINNED with respect to real gsplat (not installable here).

All discrete decisions (cull, radius, tile rectangle, depth order, alpha
threshold, early stop) are taken in fp32; ``dtype=torch.float64`` re-evaluates
the smooth arithmetic in double *with those fp32 decisions*, which gives a
summation-noise-free gradient reference.
"""
from __future__ import annotations

import math

import torch


def _quat_to_rotmat(q):
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(q.shape[:-1] + (3, 3))


def _project(means, quats, scales, viewmat, K, width, height, eps2d, near, far, radius_clip):
    """One camera.  Returns means2d, depths, conics, radius(float, no grad), valid, plus clamp masks."""
    Rwc, twc = viewmat[:3, :3], viewmat[:3, 3]
    mc = means @ Rwc.T + twc
    x, y, z = mc.unbind(-1)
    R = _quat_to_rotmat(quats)
    M = R * scales[:, None, :]
    covar = M @ M.transpose(-1, -2)
    cc = Rwc @ covar @ Rwc.T
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    lim_x = 1.3 * (0.5 * width / fx)
    lim_y = 1.3 * (0.5 * height / fy)
    rz = 1.0 / z
    tx = z * torch.minimum(lim_x, torch.maximum(-lim_x, x * rz))
    ty = z * torch.minimum(lim_y, torch.maximum(-lim_y, y * rz))
    O = torch.zeros_like(z)
    J = torch.stack([fx * rz, O, -fx * tx * rz * rz, O, fy * rz, -fy * ty * rz * rz], dim=-1).reshape(-1, 2, 3)
    cov2d = J @ cc @ J.transpose(-1, -2)
    means2d = torch.stack([fx * x * rz + cx, fy * y * rz + cy], dim=-1)
    c00 = cov2d[:, 0, 0] + eps2d
    c11 = cov2d[:, 1, 1] + eps2d
    c01 = cov2d[:, 0, 1]
    det = c00 * c11 - c01 * c01
    valid = (z >= near) & (z <= far) & (det > 0)
    det_s = torch.where(valid, det, torch.ones_like(det))
    conics = torch.stack([c11 / det_s, -c01 / det_s, c00 / det_s], dim=-1)
    b = 0.5 * (c00 + c11)
    radius = torch.ceil(3.0 * torch.sqrt(b + torch.sqrt(torch.clamp(b * b - det, min=0.01)))).detach()
    valid = valid & (radius > radius_clip)
    m2 = means2d.detach()
    valid = valid & ~((m2[:, 0] + radius <= 0) | (m2[:, 0] - radius >= width) |
                      (m2[:, 1] + radius <= 0) | (m2[:, 1] - radius >= height))
    return means2d, z, conics, radius, valid


def rasterization_torch(means, quats, scales, opacities, colors, viewmats, Ks, width, height,
                        near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, tile_size=16,
                        backgrounds=None, render_mode="RGB", dtype=torch.float32, decisions=None):
    """Inputs are torch tensors (any float dtype; cast to ``dtype``).  Returns
    (render_colors [C,H,W,D], render_alphas [C,H,W,1], meta) where
    meta["decisions"] can be passed back in to pin the discrete choices."""
    C = viewmats.shape[0]
    G = scales.shape[0]
    cast = lambda t: t.to(dtype)
    means, quats, scales, opacities, colors = map(cast, (means, quats, scales, opacities, colors))
    viewmats, Ks = cast(viewmats), cast(Ks)
    if backgrounds is not None:
        backgrounds = cast(backgrounds)
    tw = math.ceil(width / tile_size)
    th = math.ceil(height / tile_size)
    ys, xs = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
    px = (xs.reshape(-1).to(dtype) + 0.5)
    py = (ys.reshape(-1).to(dtype) + 0.5)
    ptx = (xs.reshape(-1) // tile_size)
    pty = (ys.reshape(-1) // tile_size)
    out_colors, out_alphas, dec_out = [], [], []
    metas = dict(radii=[], means2d=[], depths=[], conics=[])
    for c in range(C):
        m_c = means if means.dim() == 2 else means[c]
        q_c = quats if quats.dim() == 2 else quats[c]
        means2d, depths, conics, radius, valid = _project(m_c, q_c, scales, viewmats[c], Ks[c], width, height,
                                                          eps2d, near_plane, far_plane, radius_clip)
        if decisions is not None:
            d = decisions[c]
            radius, valid = d["radius"].to(dtype), d["valid"]
        col = colors if colors.dim() == 2 else colors[c]
        bg = backgrounds[c] if backgrounds is not None else None
        if render_mode in ("RGB+ED", "RGB+D"):
            col = torch.cat([col, depths[:, None]], dim=-1)
            if bg is not None:
                bg = torch.cat([bg, bg.new_zeros(1)])
        # tile rectangles (decisions)
        if decisions is None:
            m2 = means2d.detach().float()
            r = radius.float()
            tr = r / tile_size
            x0 = torch.floor(m2[:, 0] / tile_size - tr).clamp(0, tw).long()
            y0 = torch.floor(m2[:, 1] / tile_size - tr).clamp(0, th).long()
            x1 = torch.ceil(m2[:, 0] / tile_size + tr).clamp(0, tw).long()
            y1 = torch.ceil(m2[:, 1] / tile_size + tr).clamp(0, th).long()
            dbits = depths.detach().float().contiguous().view(torch.int32).long()
            dbits = torch.where(valid, dbits, torch.full_like(dbits, 1 << 40))
            order = torch.sort(dbits, stable=True)[1]
            order = order[valid[order]]
        else:
            x0, y0, x1, y1, order = d["x0"], d["y0"], d["x1"], d["y1"], d["order"]
        go = order
        in_tile = ((ptx[:, None] >= x0[go][None]) & (ptx[:, None] < x1[go][None]) &
                   (pty[:, None] >= y0[go][None]) & (pty[:, None] < y1[go][None]))
        dx = means2d[go, 0][None, :] - px[:, None]
        dy = means2d[go, 1][None, :] - py[:, None]
        ca, cb, cd = conics[go, 0][None], conics[go, 1][None], conics[go, 2][None]
        sigma = 0.5 * (ca * dx * dx + cd * dy * dy) + cb * dx * dy
        op = opacities[go][None]
        vis = torch.exp(-sigma)
        alpha_raw = op * vis
        alpha = torch.clamp(alpha_raw, max=0.999)
        if decisions is None:
            pair_valid = in_tile & (sigma >= 0) & (alpha >= 1.0 / 255.0)
            a_eff = torch.where(pair_valid, alpha, torch.zeros_like(alpha))
            next_T = torch.cumprod(1 - a_eff, dim=1)
            stop = pair_valid & (next_T <= 1e-4)
            included = pair_valid & (torch.cumsum(stop.long(), dim=1) == 0)
        else:
            included = d["included"]
        a_inc = torch.where(included, alpha, torch.zeros_like(alpha))
        T_incl = torch.cumprod(1 - a_inc, dim=1)
        T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], dim=1)
        w = a_inc * T_excl
        T_final = T_incl[:, -1] if T_incl.shape[1] > 0 else torch.ones_like(px)
        img = w @ col[go]
        if bg is not None:
            img = img + T_final[:, None] * bg[None]
        alpha_img = 1 - T_final
        if render_mode == "RGB+ED":
            img = torch.cat([img[:, :-1], img[:, -1:] / alpha_img[:, None].clamp(min=1e-10)], dim=-1)
        out_colors.append(img.reshape(height, width, -1))
        out_alphas.append(alpha_img.reshape(height, width, 1))
        dec_out.append(dict(radius=radius.detach(), valid=valid, x0=x0, y0=y0, x1=x1, y1=y1, order=order,
                            included=included))
        metas["radii"].append(torch.where(valid, radius, torch.zeros_like(radius)).to(torch.int32))
        metas["means2d"].append(means2d)
        metas["depths"].append(depths)
        metas["conics"].append(conics)
    meta = {k: torch.stack(v) for k, v in metas.items()}
    meta["means2d_list"] = metas["means2d"]
    meta["decisions"] = dec_out
    return torch.stack(out_colors), torch.stack(out_alphas), meta
