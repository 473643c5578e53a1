#!/usr/bin/env python
"""A/B the backward blend formulations on one config (default c3): for each D4_BWD / D4_BWD_GP_CFG
setting run a few full steps, report the mean d4_blend_bwd time (CUDA events around the C-ABI call)
and the deviation of every parameter gradient from the first mode's.

    python scripts/ab_blend_bwd.py [--config c3] [--steps 5] [--modes shfl gp:0 gp:1]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--modes", nargs="+", default=["shfl", "gp:0", "gp:1"])
    ap.add_argument("--d0", type=int, default=None, help="feature channels (default: the config's 16)")
    args = ap.parse_args()
    from deblur4dgs_b200 import _cabi
    from deblur4dgs_b200.scene import assemble_gaussians, render_subexposures
    from deblur4dgs_b200.synthetic import CONFIGS, make_config

    dev = torch.device("cuda", 0)
    G, W, H, K, N, seed = CONFIGS[args.config]
    sc = make_config(args.config).to(dev)
    extra = sc.extra_channels
    if args.d0 is not None:
        extra = extra[:, :max(0, args.d0 - 4)]
    D0 = 4 + extra.shape[1]
    bg = torch.zeros(1, D0, device=dev)
    g = torch.Generator().manual_seed(1234)
    w_img = torch.randn(1, H, W, D0 + 1, generator=g).to(dev)
    w_acc = torch.randn(1, H, W, 1, generator=g).to(dev)
    names = ["fg_means", "fg_quats", "fg_scales", "fg_colors", "fg_opacities", "motion_coefs", "bg_means",
             "bg_quats", "bg_scales", "bg_colors", "bg_opacities", "rots", "transls"]

    def step():
        p = {k: getattr(sc, k).detach().requires_grad_(True) for k in names}
        scales, opac, colors = assemble_gaussians(p["fg_scales"], p["bg_scales"], p["fg_opacities"], p["bg_opacities"],
                                                  p["fg_colors"], p["bg_colors"],
                                                  extra=extra if extra.shape[1] else None, with_mask=True)
        o = render_subexposures(p["fg_means"], p["fg_quats"], p["motion_coefs"], p["bg_means"], p["bg_quats"],
                                p["rots"], p["transls"], sc.times, sc.RTs, scales, opac, colors, sc.w2c, sc.K, W, H,
                                backgrounds=bg, render_mode="RGB+ED", combine=True, ref_quirk=True)
        torch.autograd.backward([o["img"], o["acc"]], [w_img, w_acc])
        g = {k: p[k].grad for k in names}
        g["img"], g["acc"] = o["img"].detach(), o["acc"].detach()
        return g, int(o["meta"]["isect_ids"].numel())

    base = None
    for mode in args.modes:
        if "=" in mode:  # generic form: ENV=VAL[,ENV=VAL...]
            for kv in mode.split(","):
                k, _, v = kv.partition("=")
                os.environ[k] = v
        else:  # shorthand: shfl | gp[:cfg]
            kind, _, cfg = mode.partition(":")
            os.environ["D4_BWD"] = kind
            if cfg:
                os.environ["D4_BWD_GP_CFG"] = cfg
        for _ in range(2):
            grads, n_isects = step()
        torch.cuda.synchronize()
        prof = {}
        _cabi.PROFILE = prof
        for _ in range(args.steps):
            grads, n_isects = step()
        torch.cuda.synchronize()
        _cabi.PROFILE = None
        ms = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in prof.items()}
        dev_rel = {}
        if base is None:
            base = {k: v.clone() for k, v in grads.items()}
        else:
            for k in base:
                scale = base[k].abs().max().item() + 1e-30
                dev_rel[k] = (grads[k] - base[k]).abs().max().item() / scale
        print(json.dumps({"mode": mode, "config": args.config, "D": D0 + 1, "n_isects": n_isects,
                          "blend_bwd_ms": ms.get("d4_blend_bwd"), "blend_fwd_ms": ms.get("d4_blend_fwd"),
                          "step_ms_kernels": sum(ms.values()),
                          "ms": {k.replace("d4_", ""): round(v, 3) for k, v in ms.items()},
                          "finite": all(bool(torch.isfinite(v).all()) for v in grads.values()),
                          "max_rel_dev_vs_first": max(dev_rel.values()) if dev_rel else 0.0,
                          "worst": max(dev_rel, key=dev_rel.get) if dev_rel else None}), flush=True)


if __name__ == "__main__":
    main()
