#!/usr/bin/env python
"""A/B the blend paths on one config (default c3): per C-ABI call CUDA-event times of full fwd+bwd steps.

    python scripts/ab_paths.py [--config c3] [--steps 5] [--d0 16] [--paths slab direct]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from deblur4dgs_b200 import _cabi, rendering  # noqa: E402
from deblur4dgs_b200.scene import render_subexposures  # noqa: E402
from deblur4dgs_b200.synthetic import make_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--d0", type=int, default=16)
    ap.add_argument("--paths", nargs="+", default=["slab", "direct"])
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    sc = make_config(args.config).to(dev)
    scales, opac, colors = sc.scales_all(), sc.opacities_all(), sc.colors_all(args.d0)
    bg = torch.zeros(1, args.d0, device=dev)
    g = torch.Generator().manual_seed(1)
    w_img = torch.randn(1, sc.height, sc.width, args.d0 + 1, generator=g).to(dev)
    w_acc = torch.randn(1, sc.height, sc.width, 1, generator=g).to(dev)

    def step():
        p = {k: getattr(sc, k).detach().requires_grad_(True) for k in ["fg_means", "fg_quats", "motion_coefs", "rots", "transls"]}
        cg = colors.detach().requires_grad_(True)
        o = render_subexposures(p["fg_means"], p["fg_quats"], p["motion_coefs"], sc.bg_means, sc.bg_quats, p["rots"],
                                p["transls"], sc.times, sc.RTs, scales, opac, cg, sc.w2c, sc.K, sc.width, sc.height,
                                backgrounds=bg, render_mode="RGB+ED")
        torch.autograd.backward([o["img"], o["acc"]], [w_img, w_acc])
        return o

    for path in args.paths:
        name, _, mode = path.partition(":")
        rendering.BLEND_PATH = name
        rendering.BWD_MODE = int(mode or 0) if name != "slab" else 0
        # slab:<bwd>[:<fwd>] -- e.g. slab:2:1 = tensor-core backward + queued tensor-core forward
        bm, _, fm = mode.partition(":")
        rendering.SLAB_BWD_VARIANT = int(bm) if (name == "slab" and bm) else None
        rendering.SLAB_FWD_VARIANT = int(fm) if (name == "slab" and fm) else None
        for _ in range(3):
            o = step()
        torch.cuda.synchronize()
        prof = {}
        _cabi.PROFILE = prof
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        _cabi.PROFILE = None
        ms = {k: round(sum(a.elapsed_time(b) for a, b in v) / args.steps, 4) for k, v in prof.items()}
        print(json.dumps({"path": path, "config": args.config, "d0": args.d0, "step_ms": round(e0.elapsed_time(e1) / args.steps, 3),
                          "n_isects": int(o["meta"]["isect_ids"].numel()), "ms": ms}), flush=True)


if __name__ == "__main__":
    main()
