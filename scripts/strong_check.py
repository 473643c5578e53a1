#!/usr/bin/env python
"""Multi-GPU parity of the strong-scaling partitions (torchrun, N >= 2 ranks): the frame rendered band-major
("rows") and as 2-D units ("bands") over the ranks against the SAME frame rendered by one GPU -- image, alpha and
the all-reduced gradient of every leaf.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/strong_check.py [--config c2]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from deblur4dgs_b200.parallel import allreduce_sum_, render_frame_banded, render_frame_rows  # noqa: E402
from deblur4dgs_b200.scene import assemble_gaussians, render_subexposures  # noqa: E402
from deblur4dgs_b200.synthetic import make_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    args = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    sc = make_config(args.config).to(dev)
    W, H = sc.width, sc.height
    D0 = 4 + sc.extra_channels.shape[1]
    bg = torch.zeros(1, D0, device=dev)
    g = torch.Generator().manual_seed(7)
    w_img = torch.randn(1, H, W, D0 + 1, generator=g).to(dev)
    w_acc = torch.randn(1, H, W, 1, generator=g).to(dev)
    names = ["fg_means", "fg_quats", "fg_scales", "fg_colors", "fg_opacities", "motion_coefs", "bg_means", "bg_quats",
             "bg_scales", "bg_colors", "bg_opacities", "rots", "transls"]

    def run(mode):
        p = {k: getattr(sc, k).detach().requires_grad_(True) for k in names}
        scales, opac, colors = assemble_gaussians(p["fg_scales"], p["bg_scales"], p["fg_opacities"], p["bg_opacities"],
                                                  p["fg_colors"], p["bg_colors"], extra=sc.extra_channels, with_mask=True)

        def local(times, RTs, combine, row_windows=None, camera_of=None):
            return render_subexposures(p["fg_means"], p["fg_quats"], p["motion_coefs"], p["bg_means"], p["bg_quats"],
                                       p["rots"], p["transls"], times, RTs, scales, opac, colors, sc.w2c, sc.K, W, H,
                                       backgrounds=bg, render_mode="RGB+ED", combine=combine, ref_quirk=True,
                                       row_windows=row_windows, camera_of=camera_of)

        def units(t, r, camera_of, row0, band_h):
            o = local(t, r, False, (row0, band_h), camera_of)
            return o["exposure_imgs"], o["exposure_alphas"]

        if mode == "single":
            o = local(sc.times, sc.RTs, True)
            img, acc = o["img"], o["acc"]
        elif mode == "rows":
            img, acc = render_frame_rows(sc.times, sc.RTs, H, units, ref_quirk=True)
        else:
            img, acc = render_frame_banded(sc.times, sc.RTs, H, units, ref_quirk=True)
        torch.autograd.backward([img, acc], [w_img, w_acc])
        grads = [p[k].grad for k in names]
        if mode != "single":
            allreduce_sum_(grads)
        return img.detach(), acc.detach(), grads

    ref = run("single")
    out = {}
    for mode in ("rows", "bands"):
        img, acc, grads = run(mode)
        worst = {"img": float((img - ref[0]).abs().max() / ref[0].abs().max()),
                 "acc": float((acc - ref[1]).abs().max())}
        for k, a, b in zip(names, grads, ref[2]):
            worst[k] = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        out[mode] = worst
        ok = all(v <= 2e-5 for v in worst.values())
        if rank == 0:
            print(json.dumps({"mode": mode, "config": args.config, "world": world, "ok": ok, "max_rel_dev": worst}), flush=True)
        assert ok, (mode, worst)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
