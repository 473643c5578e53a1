#!/usr/bin/env python
"""Where the HOST time of one eager step goes (cProfile over N steps of c3, fwd+bwd): the GPU work of a step is
~8 ms, so Python / ctypes / autograd overhead beyond that shows up as idle GPU time."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from deblur4dgs_b200.rendering import RenderCapacity  # noqa: E402
from deblur4dgs_b200.scene import assemble_gaussians, render_subexposures  # noqa: E402
from deblur4dgs_b200.synthetic import make_config  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    sc = make_config("c3").to(dev)
    D0 = 4 + sc.extra_channels.shape[1]
    bg = torch.zeros(1, D0, device=dev)
    w_img = torch.randn(1, sc.height, sc.width, D0 + 1, device=dev)
    w_acc = torch.randn(1, sc.height, sc.width, 1, device=dev)
    names = ["fg_means", "fg_quats", "fg_scales", "fg_colors", "fg_opacities", "motion_coefs", "bg_means", "bg_quats",
             "bg_scales", "bg_colors", "bg_opacities", "rots", "transls"]
    cap = RenderCapacity()

    def step():
        p = {k: getattr(sc, k).detach().requires_grad_(True) for k in names}
        scales, opac, colors = assemble_gaussians(p["fg_scales"], p["bg_scales"], p["fg_opacities"], p["bg_opacities"],
                                                  p["fg_colors"], p["bg_colors"], extra=sc.extra_channels, with_mask=True)
        o = render_subexposures(p["fg_means"], p["fg_quats"], p["motion_coefs"], p["bg_means"], p["bg_quats"], p["rots"],
                                p["transls"], sc.times, sc.RTs, scales, opac, colors, sc.w2c, sc.K, sc.width, sc.height,
                                backgrounds=bg, render_mode="RGB+ED", combine=True, ref_quirk=True, capacity=cap)
        torch.autograd.backward([o["img"], o["acc"]], [w_img, w_acc])

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    n = 20
    # host time of a step when the GPU is never waited for (enqueue only)
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, with drain {1e3 * (t2 - t0) / n:.3f} ms/step", flush=True)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(35)


if __name__ == "__main__":
    main()
