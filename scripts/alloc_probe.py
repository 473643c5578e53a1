#!/usr/bin/env python
"""Which part of the end-to-end loop makes the caching allocator grow (cudaMalloc inside the loop)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deblur4dgs_b200.rendering import RenderCapacity
from deblur4dgs_b200.scene import assemble_gaussians, render_subexposures
from deblur4dgs_b200.synthetic import CONFIGS, make_config

dev = torch.device("cuda", 0)
G, W, H, K, N, seed = CONFIGS["c3"]
sc = make_config("c3").to(dev)
D0 = 4 + sc.extra_channels.shape[1]
bg = torch.zeros(1, D0, device=dev)
w_img = torch.randn(1, H, W, D0 + 1, device=dev)
w_acc = torch.randn(1, H, W, 1, device=dev)
names = ["fg_means", "fg_quats", "fg_scales", "fg_colors", "fg_opacities", "motion_coefs", "bg_means", "bg_quats",
         "bg_scales", "bg_colors", "bg_opacities", "rots", "transls"]


def step(cap):
    p = {k: getattr(sc, k).detach().requires_grad_(True) for k in names}
    scales, opac, colors = assemble_gaussians(p["fg_scales"], p["bg_scales"], p["fg_opacities"], p["bg_opacities"],
                                              p["fg_colors"], p["bg_colors"], extra=sc.extra_channels, with_mask=True)
    o = render_subexposures(p["fg_means"], p["fg_quats"], p["motion_coefs"], p["bg_means"], p["bg_quats"], p["rots"],
                            p["transls"], sc.times, sc.RTs, scales, opac, colors, sc.w2c, sc.K, W, H,
                            backgrounds=bg, render_mode="RGB+ED", combine=True, ref_quirk=True, capacity=cap)
    torch.autograd.backward([o["img"], o["acc"]], [w_img, w_acc])
    return [o["img"], o["acc"]] + [p[k].grad for k in names]


def allocs():
    s = torch.cuda.memory_stats()
    return s.get("num_device_alloc", 0), s.get("reserved_bytes.all.current", 0) >> 20


def run(tag, use_cap, hold, sync_hold, steps=24):
    cap = RenderCapacity() if use_cap else None
    pending = [None, None]
    a0 = None
    for k in range(steps):
        outs = step(cap)
        if hold:
            ev = torch.cuda.Event()
            ev.record()
            if sync_hold and pending[k & 1] is not None:
                pending[k & 1][0].synchronize()
            pending[k & 1] = (ev, outs)
        if k == 7:
            a0 = allocs()
    torch.cuda.synchronize()
    print(tag, "cudaMallocs / reserved MiB after 8 steps", a0, "after", steps, allocs(), flush=True)


run("sync mode, outputs dropped", False, False, False)
run("capacity mode, outputs dropped", True, False, False)
run("capacity mode, outputs held 2 steps (no host wait)", True, True, False)
run("capacity mode, outputs held 2 steps + host waits for step k-2", True, True, True)
run("sync mode, outputs held 2 steps + host waits for step k-2", False, True, True)
