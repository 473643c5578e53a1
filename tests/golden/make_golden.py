"""Generate the committed golden fixtures under tests/golden/.

Run HERE (the authoring container), never on the GPU box:

    python tests/golden/make_golden.py

1. ``deform_*.npz`` -- produced by the REFERENCE'S OWN code imported from
   /root/reference: ``flow3d.params.GaussianParams`` / ``MotionBases``
   (params.py:10-180), ``flow3d.transforms.cont_6d_to_rmat``
   (transforms.py:41-53) and the unbound ``SceneModel.compute_poses_fg`` /
   ``compute_poses_all`` / ``compute_transforms`` methods
   (scene_model.py:67-120), plus the camera sub-exposure transform lines
   (scene_model.py:352-353) restated verbatim below.  Third-party modules that
   are absent here are stubbed in ``sys.modules``: ``roma`` by
   ``oracle/roma_shim.py`` (a restatement), ``gsplat`` / ``pypose`` /
   ``jaxtyping`` / ``cv2`` by empty placeholders (never called on this path).
   Gradients are torch autograd of that reference code for fixed cotangents.

1b. ``scene_render.npz`` -- the reference's REAL ``SceneModel.render`` (scene_model.py:162-487: flag handling, feature
   vector assembly, the loop over the 11 sub-exposures, the in-place combine, the densifier side channel, the
   out_dict) and the reference's real ``MoveModel`` MLP / exposure code (move_model.py:66-135, 148-165), run on the
   CPU.  What is substituted, because it is absent here: ``gsplat.rendering.rasterization`` by an autograd wrapper
   around the oracle (oracle/raster.py), the pypose SE(3) interpolation inside ``forward_start_end_mid`` by
   ``oracle/camera.py::camera_interp`` (both restatements: "parity unpinned" for those two), ``.cuda()`` by a
   no-op, and the debug ``cv2.imwrite`` / ``os.makedirs`` to a hard-coded absolute path (scene_model.py:375-378)
   by no-ops.  Also stored: the running stats the reference's ``Trainer._prepare_control_step`` loop
   (trainer.py:967-989, restated literally) accumulates from that render.

2. ``raster_*.npz`` -- produced by ``oracle/raster.py`` (the C oracle) on the
   seeded synthetic scenes.  These are NOT reference outputs (gsplat cannot be
   run here: "parity unpinned"); they freeze the oracle so that a later edit
   of the oracle cannot silently move the target.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from deblur4dgs_b200.synthetic import make_scene  # noqa: E402
from oracle import roma_shim  # noqa: E402


def _import_reference():
    sys.modules["roma"] = roma_shim
    g = types.ModuleType("gsplat")
    gr = types.ModuleType("gsplat.rendering")
    gr.rasterization = None
    g.rendering = gr
    sys.modules["gsplat"] = g
    sys.modules["gsplat.rendering"] = gr
    pp = types.ModuleType("pypose")
    pp.LieTensor = object
    sys.modules["pypose"] = pp
    jt = types.ModuleType("jaxtyping")
    jt.Float = type("Float", (), {"__class_getitem__": classmethod(lambda cls, k: cls)})
    sys.modules["jaxtyping"] = jt
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            sys.modules["cv2"] = types.ModuleType("cv2")
    sys.path.insert(0, "/root/reference")
    from flow3d.params import GaussianParams, MotionBases
    from flow3d.scene_model import SceneModel
    return GaussianParams, MotionBases, SceneModel


def deform_golden(name, G, K, N, seed, int_ts=False, out_of_range=False):
    GaussianParams, MotionBases, SceneModel = _import_reference()
    sc = make_scene(G=G, width=64, height=48, K=K, N=N, seed=seed)
    if out_of_range:
        sc.times = torch.linspace(-0.75, 8.25, N)  # exercises the clamp of floor/ceil (params.py:152-153)
    t = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sc.tensors().items()}
    fg = GaussianParams(t["fg_means"], t["fg_quats"], t["fg_scales"], t["fg_colors"], t["fg_opacities"],
                        motion_coefs=t["motion_coefs"])
    bg = GaussianParams(t["bg_means"], t["bg_quats"], t["bg_scales"], t["bg_colors"], t["bg_opacities"])
    mb = MotionBases(t["rots"], t["transls"])
    # a stand-in for `self` carrying exactly what the unbound methods touch
    me = types.SimpleNamespace(fg=fg, bg=bg, motion_bases=mb, has_bg=True)
    me.compute_transforms = lambda ts, inds=None: SceneModel.compute_transforms(me, ts, inds)
    me.compute_poses_fg = lambda ts, inds=None: SceneModel.compute_poses_fg(me, ts, inds)
    me.compute_poses_bg = lambda: SceneModel.compute_poses_bg(me)
    times = t["times"]
    RTs = t["RTs"]
    leaves = dict(fg_means=fg.params["means"], fg_quats=fg.params["quats"], motion_coefs=fg.params["motion_coefs"],
                  bg_means=bg.params["means"], bg_quats=bg.params["quats"], rots=mb.params["rots"],
                  transls=mb.params["transls"], times=times, RTs=RTs)
    out = {}
    # (i) compute_transforms at B timestamps at once (trainer.py:478,485 style call)
    ts_b = times.detach().clone()
    if int_ts:
        ts_b = torch.arange(0, min(N, 8))
    transfms = SceneModel.compute_transforms(me, ts_b)
    out["transforms_ts"] = ts_b.numpy()
    out["transforms"] = transfms.detach().numpy()
    # (ii) the render loop body for every sub-exposure (scene_model.py:323-353)
    all_m, all_q = [], []
    for ii in range(N):
        time = times[None, ii:ii + 1]
        means, quats = SceneModel.compute_poses_all(me, time)
        means, quats = means[:, 0], quats[:, 0]
        transR, transT = RTs[ii][:3, :3], RTs[ii][:3, 3:4]
        means = transR @ means.permute(1, 0) + transT
        means = means.permute(1, 0)
        all_m.append(means)
        all_q.append(quats)
    M, Q = torch.stack(all_m), torch.stack(all_q)
    out["out_means"], out["out_quats"] = M.detach().numpy(), Q.detach().numpy()
    g = torch.Generator().manual_seed(seed + 100)
    vM, vQ = torch.randn(M.shape, generator=g), torch.randn(Q.shape, generator=g)
    out["v_out_means"], out["v_out_quats"] = vM.numpy(), vQ.numpy()
    grads = torch.autograd.grad((M * vM).sum() + (Q * vQ).sum(), list(leaves.values()), allow_unused=True)
    for k, gr in zip(leaves.keys(), grads):
        out["grad_" + k] = gr.detach().numpy()
        out["in_" + k] = leaves[k].detach().numpy()
    np.savez_compressed(os.path.join(HERE, f"deform_{name}.npz"), **out)
    print("wrote deform_%s.npz" % name, {k: v.shape for k, v in out.items() if k.startswith("out")})


def raster_golden(name, G, W, H, seed, d0, mode, scale_mult=1.0, C=1):
    from oracle import raster as orc
    sc = make_scene(G=G, width=W, height=H, K=4, N=1, seed=seed, scale_mult=scale_mult)
    means = torch.cat([sc.fg_means, sc.bg_means]).numpy()
    quats = torch.cat([sc.fg_quats, sc.bg_quats]).numpy()
    scales, opac, colors = sc.scales_all().numpy(), sc.opacities_all().numpy(), sc.colors_all(d0).numpy()
    vm = sc.w2c.repeat(C, 1, 1).numpy().copy()
    for c in range(1, C):
        vm[c, 0, 3] = 0.15 * c
        vm[c, 2, 3] = 0.1 * c
    Ks = sc.K.repeat(C, 1, 1).numpy()
    g = torch.Generator().manual_seed(seed + 7)
    bg = torch.rand(C, d0, generator=g).numpy()
    rc, ra, meta = orc.rasterization(means, quats, scales, opac, colors, vm, Ks, W, H, backgrounds=bg,
                                     render_mode=mode)
    vc = torch.randn(rc.shape, generator=g).numpy()
    va = torch.randn(ra.shape, generator=g).numpy()
    vc[meta["edge"] != 0] = 0  # no cotangent on knife-edge pixels (see tests/test_gpu_parity.py)
    va[meta["edge"] != 0] = 0
    grads = orc.rasterization_backward(meta, ra, vc, va)
    out = dict(means=means, quats=quats, scales=scales, opacities=opac, colors=colors, viewmats=vm, Ks=Ks,
               backgrounds=bg, width=W, height=H, render_mode=mode,
               render_colors=rc.astype(np.float16 if False else np.float32), render_alphas=ra,
               radii=meta["radii"], means2d=meta["means2d"], depths=meta["depths"], conics=meta["conics"],
               tiles_per_gauss=meta["tiles_per_gauss"], isect_ids=meta["isect_ids"],
               flatten_ids=meta["flatten_ids"], isect_offsets=meta["isect_offsets"], last_ids=meta["last_ids"],
               edge=meta["edge"], v_render_colors=vc, v_render_alphas=va)
    for k, v in grads.items():
        if v is not None:
            out["grad_" + k] = v.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, f"raster_{name}.npz"), **out)
    print("wrote raster_%s.npz" % name, "n_isects", meta["isect_ids"].shape[0], "edge px", int(meta["edge"].sum()))


def camera_golden():
    """Reference's own spline_utils.se3_to_SE3 (pure torch) on seeded 6-vectors, with autograd grads."""
    _import_reference()
    from flow3d.models.utils.spline_utils import se3_to_SE3
    g = torch.Generator().manual_seed(77)
    wu = torch.cat([0.3 * torch.randn(64, 6, generator=g), 1e-4 * torch.randn(16, 6, generator=g), torch.zeros(1, 6)])
    wu.requires_grad_(True)
    Rt = se3_to_SE3(wu)
    v = torch.randn(Rt.shape, generator=g)
    (gw,) = torch.autograd.grad((Rt * v).sum(), wu)
    np.savez_compressed(os.path.join(HERE, "camera_se3.npz"), wu=wu.detach().numpy(), Rt=Rt.detach().numpy(),
                        v=v.numpy(), grad_wu=gw.numpy())
    print("wrote camera_se3.npz", tuple(Rt.shape))


class _CpuRasterization(torch.autograd.Function):
    """gsplat.rendering.rasterization on the CPU through the oracle (forward + hand-derived backward).  The
    screen-space gradient dL/dmeans2d -- what ``info["means2d"].grad`` holds in gsplat -- is parked in ``stash``."""

    @staticmethod
    def forward(ctx, means, quats, scales, opacities, colors, backgrounds, viewmats, Ks, width, height, mode, stash):
        from oracle import raster as orc
        n = lambda t: t.detach().contiguous().numpy()
        rc, ra, meta = orc.rasterization(n(means), n(quats), n(scales), n(opacities), n(colors), n(viewmats), n(Ks),
                                         width, height, backgrounds=n(backgrounds), render_mode=mode)
        ctx.meta, ctx.ra, ctx.stash = meta, ra, stash
        stash["meta"] = meta
        means2d = torch.from_numpy(meta["means2d"].copy())
        ctx.mark_non_differentiable()
        return torch.from_numpy(rc.copy()), torch.from_numpy(ra.copy()), means2d

    @staticmethod
    def backward(ctx, v_rc, v_ra, _v_means2d):
        from oracle import raster as orc
        g = orc.rasterization_backward(ctx.meta, ctx.ra, v_rc.contiguous().numpy(), v_ra.contiguous().numpy(),
                                       want_viewmats=False)
        ctx.stash["means2d_grad"] = torch.from_numpy(np.asarray(g["means2d"], np.float32).copy())
        t = lambda k: torch.from_numpy(np.asarray(g[k], np.float32).copy())
        return (t("means"), t("quats"), t("scales"), t("opacities"), t("colors"), None, None, None, None, None, None,
                None)


def scene_render_golden():
    GaussianParams, MotionBases, SceneModel = _import_reference()
    import flow3d.scene_model as sm_mod
    import flow3d.models.move_model as mm_mod
    from oracle import camera as ocam
    from oracle import raster as orc
    orc.set_num_threads(8)

    stashes = []

    def rasterization(means, quats, scales, opacities, colors, backgrounds, viewmats, Ks, width, height, packed=False,
                      render_mode="RGB"):
        assert packed is False
        stash = {}
        stashes.append(stash)
        rc, ra, means2d = _CpuRasterization.apply(means, quats, scales, opacities, colors, backgrounds, viewmats, Ks,
                                                  width, height, render_mode, stash)
        info = {"means2d": means2d, "radii": torch.from_numpy(stash["meta"]["radii"].copy()), "width": width,
                "height": height}
        return rc, ra, info

    def forward_start_end_mid(self, info, num_cameras=10, mode="uniform", stage="second"):
        """move_model.py:138-166 with the pypose calls (:145-146) replaced by oracle/camera.py::camera_interp."""
        R, T, time = info["R"], info["T"], info["timestep"]
        RT_start, RT_end, time_start, time_end = self.forward(R, T, time, stage=stage)
        RTs = ocam.camera_interp(RT_start[0], RT_end[0], num_cameras)
        num_fg = time_start.shape[0]
        time_start = time_start.unsqueeze(-1).repeat(1, num_cameras)
        time_end = time_end.unsqueeze(-1).repeat(1, num_cameras)
        weights = (torch.arange(num_cameras) / (num_cameras - 1)).to(RTs.device)
        weights = weights.unsqueeze(0).repeat(num_fg, 1)
        times = (time_start + time) * (1. - weights) + (time_end + time) * weights
        times = times.reshape(num_fg, num_cameras)
        deltaT = torch.abs(time_end[:, num_cameras - 1:])
        return RTs, times, deltaT

    sm_mod.rasterization = rasterization
    sm_mod.cv2 = types.SimpleNamespace(imwrite=lambda *a, **k: True)
    sm_mod.os = types.SimpleNamespace(makedirs=lambda *a, **k: None, path=os.path)
    mm_mod.MoveModel.forward_start_end_mid = forward_start_end_mid
    cuda0 = torch.nn.Module.cuda
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        W, H, t_frame = 64, 48, 3
        sc = make_scene(G=1500, width=W, height=H, K=5, N=3, seed=55, scale_mult=2.5)
        fg = GaussianParams(sc.fg_means, sc.fg_quats, sc.fg_scales, sc.fg_colors, sc.fg_opacities, motion_coefs=sc.motion_coefs)
        bg = GaussianParams(sc.bg_means, sc.bg_quats, sc.bg_scales, sc.bg_colors, sc.bg_opacities)
        mb = MotionBases(sc.rots, sc.transls)
        model = SceneModel(sc.K, sc.w2c, fg, mb, bg)
    finally:
        torch.nn.Module.cuda = cuda0
    torch.manual_seed(3)
    with torch.no_grad():  # non-trivial camera deltas and exposure (the heads are zero-initialised)
        for head in (model.move_model.RT_head0, model.move_model.RT_head1):
            head[-1].bias.copy_(0.01 * torch.randn(6))
        model.move_model.time_params.copy_(torch.tensor([[0.5, 0.3, 0.7, 0.45, 0.2, 0.95, 0.6, 0.5]]))
    target_ts = torch.tensor([1.0, 2.0, 4.0, 5.0])
    target_w2cs = sc.w2c.repeat(4, 1, 1).clone()
    target_w2cs[:, 0, 3] = torch.tensor([0.05, -0.05, 0.1, -0.1])
    out = model.render(t_frame, sc.w2c, sc.K, (W, H), target_ts=target_ts, target_w2cs=target_w2cs, return_depth=True,
                       return_mask=True, mode="blury", stage="second")
    N = len(stashes)
    assert N == 11 and len(model._current_xys) == 11 and out["exposure_imgs"].shape == (11, 1, H, W, 17)
    edge_any = np.zeros((1, H, W), bool)
    for st in stashes:
        edge_any |= st["meta"]["edge"] != 0
    ok = torch.from_numpy(~edge_any)
    # a scalar loss over what the trainer uses; no cotangent on knife-edge pixels, nor on the max (mask) / min (depth)
    # channels (their arg-extremum may legitimately differ between two fp32 evaluations)
    g = torch.Generator().manual_seed(5)
    keys = ["img", "tracks_3d", "acc"]
    wts = {k: torch.randn(out[k].shape, generator=g) for k in keys}
    loss = 0
    for k in keys:
        m = ok.reshape((1, H, W) + (1,) * (out[k].dim() - 3))
        wts[k] = wts[k] * m
        loss = loss + (out[k] * wts[k]).sum()
    params = {"fg." + k: v for k, v in fg.params.items()}
    params.update({"bg." + k: v for k, v in bg.params.items()})
    params.update({"motion_bases." + k: v for k, v in mb.params.items()})
    mm_params = dict(model.move_model.named_parameters())
    grads = torch.autograd.grad(loss, list(params.values()) + list(mm_params.values()), allow_unused=True)
    # the trainer's control-step loop (trainer.py:967-989) restated literally on the reference's side channel
    G = sc.G
    stats = {"xys_grad_norm_acc": torch.zeros(G), "vis_count": torch.zeros(G, dtype=torch.int64),
             "max_radii": torch.zeros(G)}
    xys_grads = [st["means2d_grad"] for st in stashes]
    batch_size = 1
    for ii in range(len(model._current_xys)):
        _current_radii, _current_img_wh = model._current_radii, model._current_img_wh
        sel = _current_radii[ii] > 0
        gidcs = torch.where(sel)[1]
        xys_grad = xys_grads[ii].clone()
        xys_grad[..., 0] *= _current_img_wh[0] / 2.0 * batch_size * len(model._current_xys)
        xys_grad[..., 1] *= _current_img_wh[1] / 2.0 * batch_size * len(model._current_xys)
        stats["xys_grad_norm_acc"].index_add_(0, gidcs, xys_grad[sel].norm(dim=-1))
        stats["vis_count"].index_add_(0, gidcs, torch.ones_like(gidcs, dtype=torch.int64))
    save = {"width": W, "height": H, "t": t_frame, "target_ts": target_ts.numpy(), "target_w2cs": target_w2cs.numpy(),
            "w2c": sc.w2c.numpy(), "K": sc.K.numpy(), "ok": ok.numpy()}
    for k, v in sc.tensors().items():
        save["scene_" + k] = v.numpy()
    for k, v in model.move_model.state_dict().items():
        save["mm_" + k] = v.numpy()
    for k in ["img", "mask", "tracks_3d", "depth", "acc", "deltaT", "RTs", "pred_sharp_img"]:
        save["out_" + k] = out[k].detach().numpy()
    save["out_exposure_first"] = out["exposure_imgs"][0].detach().numpy()
    save["out_exposure_last"] = out["exposure_imgs"][-1].detach().numpy()
    save["radii"] = torch.stack(model._current_radii).numpy()                # [11,1,G]
    save["means2d_grad"] = torch.stack(xys_grads).numpy().astype(np.float32)  # [11,1,G,2]
    for k in keys:
        save["w_" + k] = wts[k].numpy()
    for (k, _), gr in zip(list(params.items()) + [("mm." + k, v) for k, v in mm_params.items()], grads):
        if gr is not None:
            save["grad_" + k] = gr.numpy()
    save["stat_xys_grad_norm_acc"] = stats["xys_grad_norm_acc"].numpy()
    save["stat_vis_count"] = stats["vis_count"].numpy()
    np.savez_compressed(os.path.join(HERE, "scene_render.npz"), **save)
    print("wrote scene_render.npz", {k: tuple(v.shape) for k, v in save.items() if k.startswith("out_")},
          "edge px", int(edge_any.sum()), "visible", int((save["radii"] > 0).sum()))


def checkpoint_golden():
    """A checkpoint in the reference's own on-disk layout (trainer.py:126-140): the state dicts come from
    the reference's GaussianParams / MotionBases modules, nested exactly as SceneModel registers them
    (scene_model.py:24-30: fg, motion_bases, bg, buffers bg_scene_scale / Ks / w2cs)."""
    GaussianParams, MotionBases, _ = _import_reference()
    sc = make_scene(G=900, width=96, height=64, K=5, N=3, seed=31)
    fg = GaussianParams(sc.fg_means, sc.fg_quats, sc.fg_scales, sc.fg_colors, sc.fg_opacities, motion_coefs=sc.motion_coefs)
    bg = GaussianParams(sc.bg_means, sc.bg_quats, sc.bg_scales, sc.bg_colors, sc.bg_opacities, scene_scale=1.3)
    mb = MotionBases(sc.rots, sc.transls)
    T = sc.rots.shape[1]
    sd = {}
    sd.update({"fg." + k: v.detach().clone() for k, v in fg.state_dict().items()})
    sd.update({"motion_bases." + k: v.detach().clone() for k, v in mb.state_dict().items()})
    sd.update({"bg." + k: v.detach().clone() for k, v in bg.state_dict().items()})
    sd["bg_scene_scale"] = torch.tensor(1.3)
    sd["Ks"] = sc.K.repeat(T, 1, 1)
    w2cs = sc.w2c.repeat(T, 1, 1).clone()
    w2cs[:, 0, 3] = torch.linspace(-0.1, 0.1, T)
    sd["w2cs"] = w2cs
    torch.save({"model": sd, "optimizers": {}, "schedulers": {}, "global_step": 1234, "epoch": 7,
                "move_model": {"time_params": torch.full((1, 8), 0.5)}}, os.path.join(HERE, "ckpt_small.pt"))
    print("wrote ckpt_small.pt", sorted(sd.keys()))


if __name__ == "__main__":
    torch.set_num_threads(4)
    camera_golden()
    checkpoint_golden()
    scene_render_golden()
    deform_golden("k6_n5", G=600, K=6, N=5, seed=11)
    deform_golden("k10_n9", G=900, K=10, N=9, seed=12)
    deform_golden("k3_n1_int", G=257, K=3, N=1, seed=13, int_ts=True)
    deform_golden("k16_n13_oob", G=300, K=16, N=13, seed=14, out_of_range=True)
    raster_golden("small_rgb", G=500, W=96, H=64, seed=21, d0=3, mode="RGB", scale_mult=3.0)
    raster_golden("small_ed5", G=800, W=112, H=80, seed=22, d0=4, mode="RGB+ED", scale_mult=2.0)
    raster_golden("small_ed17_c2", G=700, W=100, H=70, seed=23, d0=16, mode="RGB+ED", scale_mult=2.5, C=2)
