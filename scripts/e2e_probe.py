#!/usr/bin/env python
"""Where does the end-to-end step lose time against the resident step?  Times the c3 step with
(upload, download) toggled.  python scripts/e2e_probe.py

Finding (B200, session 2): the upload costs 0.1 ms, the download 1.3 ms per step although it runs on a side
stream -- the 16-byte intersection-count read-back of the NEXT step queues behind the 87 MB result download on
the D2H copy engine.  Publishing the count through mapped pinned memory with a kernel (+ event wait) removed that
stall (11.85 ms/step) but made the end-to-end loop bimodal (sporadic 30-100 ms host-side stalls, also with a
blocking event), so the copy-engine read-back stays."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deblur4dgs_b200.scene import assemble_gaussians, render_subexposures
from deblur4dgs_b200.synthetic import CONFIGS, make_config

dev = torch.device("cuda", 0)
G, W, H, K, N, seed = CONFIGS["c3"]
sc_cpu = make_config("c3")
host = {k: v.pin_memory() for k, v in sc_cpu.tensors().items()}
sc = sc_cpu.to(dev)
D0 = 4 + sc.extra_channels.shape[1]
bg = torch.zeros(1, D0, device=dev)
g = torch.Generator().manual_seed(1234)
w_img = torch.randn(1, H, W, D0 + 1, generator=g).to(dev)
w_acc = torch.randn(1, H, W, 1, generator=g).to(dev)
names = ["fg_means", "fg_quats", "fg_scales", "fg_colors", "fg_opacities", "motion_coefs", "bg_means", "bg_quats",
         "bg_scales", "bg_colors", "bg_opacities", "rots", "transls"]

def step(scn):
    p = {k: getattr(scn, k).detach().requires_grad_(True) for k in names}
    scales, opac, colors = assemble_gaussians(p["fg_scales"], p["bg_scales"], p["fg_opacities"], p["bg_opacities"],
                                              p["fg_colors"], p["bg_colors"], extra=scn.extra_channels, with_mask=True)
    o = render_subexposures(p["fg_means"], p["fg_quats"], p["motion_coefs"], p["bg_means"], p["bg_quats"], p["rots"],
                            p["transls"], scn.times, scn.RTs, scales, opac, colors, scn.w2c, scn.K, W, H,
                            backgrounds=bg, render_mode="RGB+ED", combine=True, ref_quirk=True)
    torch.autograd.backward([o["img"], o["acc"]], [w_img, w_acc])
    return [o["img"], o["acc"]] + [p[k].grad for k in names]

copy_stream, up_stream = torch.cuda.Stream(), torch.cuda.Stream()

def run(upload, download, steps=10):
    out_host, pending, uploaded = [None, None], [None, None], {}
    def up(k):
        with torch.cuda.stream(up_stream):
            scn = type(sc)(**{kk: v.to(dev, non_blocking=True) for kk, v in host.items()}, width=W, height=H)
            ev = torch.cuda.Event(); ev.record(up_stream)
        uploaded[k] = (scn, ev)
    def one(k):
        if upload:
            if k not in uploaded: up(k)
            scn, ev = uploaded.pop(k)
            torch.cuda.current_stream().wait_event(ev)
            up(k + 1)
        else:
            scn = sc
        outs = step(scn)
        if upload:
            for t in scn.tensors().values(): t.record_stream(torch.cuda.current_stream())
        if download:
            buf = k & 1
            if out_host[buf] is None:
                out_host[buf] = [torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in outs]
            if pending[buf] is not None: pending[buf][0].synchronize()
            ready = torch.cuda.Event(); ready.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                for h, o in zip(out_host[buf], outs):
                    h.copy_(o, non_blocking=True); o.record_stream(copy_stream)
                done = torch.cuda.Event(); done.record(copy_stream)
            pending[buf] = (done, outs)
    for k in range(3): one(k)
    torch.cuda.synchronize(); uploaded.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for k in range(steps): one(k)
    torch.cuda.current_stream().wait_stream(copy_stream); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps

for up_, down_ in [(False, False), (True, False), (False, True), (True, True)]:
    ms, wall = run(up_, down_)
    print(f"upload={up_} download={down_}: {ms:.3f} ms/step (wall {wall:.3f})", flush=True)
