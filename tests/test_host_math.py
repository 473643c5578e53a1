"""CPU checks (-m "not gpu") of the product's per-Gaussian arithmetic.

csrc/project_math.cuh and csrc/deform_math.cuh are __host__ __device__; the
test-only harness tests/host_harness/harness.cpp compiles them with g++ so the
very arithmetic the CUDA kernels execute is compared with the oracle here,
where there is no GPU: projection bit-exactly (radii, means2d, depths, conics,
tiles_per_gauss), projection backward and the dual-number deformation VJP to
fp32 round-off.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from deblur4dgs_b200.synthetic import make_scene
from oracle import deform as odef
from oracle import raster as orc
from util import quat_sign_align, rel_err, scale_err

HH_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_harness")


@pytest.fixture(scope="module")
def hh():
    src, lib = os.path.join(HH_DIR, "harness.cpp"), os.path.join(HH_DIR, "libhh.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", lib, src],
                   check=True)
    return ctypes.CDLL(lib)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("seed,scale_mult", [(1, 1.0), (2, 4.0), (3, 0.3)])
def test_projection_bit_exact_and_backward(hh, seed, scale_mult):
    W, H, G = 512, 288, 20000
    sc = make_scene(G=G, width=W, height=H, K=4, N=1, seed=seed, scale_mult=scale_mult)
    means = torch.cat([sc.fg_means, sc.bg_means]).numpy().copy()
    means[:50, 2] = -1.0  # behind the camera
    means[50:100, 2] = 0.005  # in front of the near plane
    quats = torch.cat([sc.fg_quats, sc.bg_quats]).numpy()
    scales = sc.scales_all().numpy()
    V = sc.w2c[0].numpy().copy()
    a = 0.07
    V[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
    V[:3, 3] = [0.05, -0.02, 0.1]
    K = sc.K[0].numpy()
    radii, m2, dep, con = orc.project_fwd(means, quats, scales, V[None], K[None], W, H)
    tpg, _, _ = orc.isect_tiles(m2, radii, dep, 16, 32, 18, sort=False)
    r2 = np.zeros(G, np.int32); m22 = np.zeros((G, 2), np.float32); d2 = np.zeros(G, np.float32)
    c2 = np.zeros((G, 3), np.float32); t2 = np.zeros(G, np.int32)
    hh.hh_project_fwd(_p(means), _p(quats), _p(scales), _p(V), _p(K), G, W, H, ctypes.c_float(0.3),
                      ctypes.c_float(0.01), ctypes.c_float(1e10), ctypes.c_float(0.0), 16, 32, 18, _p(r2), _p(m22),
                      _p(d2), _p(c2), _p(t2))
    assert (radii[0] > 0).sum() > 1000 and (radii[0] == 0).sum() > 100
    assert np.array_equal(r2, radii[0])
    assert np.array_equal(m22.view(np.int32), m2[0].view(np.int32))
    assert np.array_equal(d2.view(np.int32), dep[0].view(np.int32))
    assert np.array_equal(c2.view(np.int32), con[0].view(np.int32))
    assert np.array_equal(t2, tpg[0])
    # backward
    rng = np.random.default_rng(seed)
    vm2 = rng.standard_normal((1, G, 2)).astype(np.float32)
    vd = rng.standard_normal((1, G)).astype(np.float32)
    vc = rng.standard_normal((1, G, 3)).astype(np.float32)
    om, oq, os_, ov = orc.project_bwd(means, quats, scales, V[None], K[None], W, H, radii, con, vm2, vd, vc)
    vmeans = np.zeros((G, 3), np.float32); vquats = np.zeros((G, 4), np.float32); vscales = np.zeros((G, 3), np.float32)
    vview = np.zeros(16, np.float64)
    hh.hh_project_bwd(_p(means), _p(quats), _p(scales), _p(V), _p(K), G, W, H, _p(radii), _p(con), _p(vm2), _p(vd),
                      _p(vc), _p(vmeans), _p(vquats), _p(vscales), _p(vview))
    assert scale_err(vmeans, om) < 1e-6 and scale_err(vquats, oq) < 1e-6 and scale_err(vscales, os_) < 1e-6
    assert scale_err(vview.reshape(4, 4), ov[0]) < 1e-6


def test_deform_point_and_dual_vjp(hh):
    n = 4000
    g = torch.Generator().manual_seed(5)
    bl = torch.randn(n, 9, generator=g)
    bl[:, 3:] = torch.tensor([1.0, 0, 0, 0, 1.0, 0]) + 0.6 * torch.randn(n, 6, generator=g)  # all 4 quat branches
    bl[: n // 4, 3:] = torch.randn(n // 4, 6, generator=g)
    mu = torch.randn(n, 3, generator=g)
    q = torch.randn(n, 4, generator=g)
    vm, vq = torch.randn(n, 3, generator=g), torch.randn(n, 4, generator=g)
    leaves = [x.clone().requires_grad_(True) for x in (bl, mu, q)]
    R = odef.cont_6d_to_rmat(leaves[0][:, 3:])
    om_ref = torch.einsum("pij,pj->pi", R, leaves[1]) + leaves[0][:, :3]
    qh = torch.nn.functional.normalize(leaves[2], dim=-1)
    oq_ref = odef.roma.quat_xyzw_to_wxyz(odef.roma.quat_product(odef.roma.rotmat_to_unitquat(R),
                                                                odef.roma.quat_wxyz_to_xyzw(qh)))
    oq_ref = torch.nn.functional.normalize(oq_ref, dim=-1)
    om = np.zeros((n, 3), np.float32); oq = np.zeros((n, 4), np.float32)
    bln, mun, qn = bl.numpy(), mu.numpy(), q.numpy()
    hh.hh_deform_point(_p(bln), _p(mun), _p(qn), n, _p(om), _p(oq))
    assert rel_err(om, om_ref.detach().numpy()) < 1e-4
    assert np.all((oq * oq_ref.detach().numpy()).sum(-1) > 0.999)  # same sign, same branch
    assert rel_err(oq, oq_ref.detach().numpy()) < 1e-4
    gb, gm, gq = torch.autograd.grad((om_ref * vm).sum() + (oq_ref * vq).sum(), leaves)
    grad = np.zeros((n, 16), np.float32)
    vmn, vqn = vm.numpy(), vq.numpy()
    hh.hh_deform_point_vjp(_p(bln), _p(mun), _p(qn), _p(vmn), _p(vqn), n, _p(grad))
    ref = np.concatenate([gb.numpy(), gm.numpy(), gq.numpy()], axis=1)
    err = np.abs(grad - ref) / (np.abs(ref) + 1e-3 * np.abs(ref).max(axis=1, keepdims=True) + 1e-6)
    assert np.quantile(err, 0.999) < 1e-3, np.quantile(err, [0.5, 0.99, 0.999, 1.0])
    assert scale_err(grad, ref) < 1e-4
    # the hand-derived reverse-mode VJP (what deform_fg_bwd_kernel runs) against torch autograd and against the
    # dual-number evaluation of the same function; it also returns the deformed mean for the camera-delta gradient
    grad_rev = np.zeros((n, 16), np.float32)
    om_rev = np.zeros((n, 3), np.float32)
    hh.hh_deform_point_vjp_rev(_p(bln), _p(mun), _p(qn), _p(vmn), _p(vqn), n, _p(grad_rev), _p(om_rev))
    assert np.array_equal(om_rev, om)
    err = np.abs(grad_rev - ref) / (np.abs(ref) + 1e-3 * np.abs(ref).max(axis=1, keepdims=True) + 1e-6)
    assert np.quantile(err, 0.999) < 1e-3, np.quantile(err, [0.5, 0.99, 0.999, 1.0])
    assert scale_err(grad_rev, ref) < 1e-4
    assert scale_err(grad_rev, grad) < 5e-5  # two fp32 evaluation orders of the same derivative (measured 1.3e-5)


def test_camera_se3_pinned_to_reference_and_interp(hh):
    """Row a7.  (1) oracle/camera.se3_to_SE3 and the product's se3_to_SE3_mat vs the fixture produced by
    the reference's own spline_utils.se3_to_SE3; (2) the product's camera_interp_one (float and
    Dual<12>) vs the oracle restatement and its autograd."""
    from oracle import camera as ocam
    from util import golden
    g = golden("camera_se3.npz")
    wu = torch.from_numpy(g["wu"]).requires_grad_(True)
    Rt = ocam.se3_to_SE3(wu)
    assert rel_err(Rt.detach().numpy(), g["Rt"]) < 1e-5
    (gw,) = torch.autograd.grad((Rt * torch.from_numpy(g["v"])).sum(), wu)
    assert scale_err(gw.numpy(), g["grad_wu"]) < 1e-5
    out = np.zeros((g["wu"].shape[0], 12), np.float32)
    wun = np.ascontiguousarray(g["wu"])
    hh.hh_se3_to_SE3(_p(wun), wun.shape[0], _p(out))
    assert rel_err(out.reshape(-1, 3, 4), g["Rt"]) < 1e-5
    # interpolation: generic, tiny-angle (Taylor branches) and exactly-zero heads (the init state)
    gen = torch.Generator().manual_seed(3)
    for scale in (0.3, 0.01, 1e-5, 0.0):
        s6 = (scale * torch.randn(6, generator=gen)).requires_grad_(True)
        e6 = (scale * torch.randn(6, generator=gen)).requires_grad_(True)
        N = 11
        ref = ocam.camera_interp(s6, e6, N)
        us = torch.linspace(0, 1, N).numpy().copy()
        got = np.zeros((N, 12), np.float32)
        sn, en = s6.detach().numpy().copy(), e6.detach().numpy().copy()
        hh.hh_camera_interp(_p(sn), _p(en), _p(us), N, _p(got))
        assert rel_err(got.reshape(N, 3, 4), ref.detach().numpy()) < 1e-4, scale
        v = torch.randn(N, 3, 4, generator=gen)
        gs, ge = torch.autograd.grad((ref * v).sum(), [s6, e6])
        g12 = np.zeros(12, np.float32)
        vn = v.numpy().reshape(N, 12).copy()
        hh.hh_camera_interp_vjp(_p(sn), _p(en), _p(us), N, _p(vn), _p(g12))
        refg = np.concatenate([gs.numpy(), ge.numpy()])
        assert scale_err(g12, refg) < 2e-3, (scale, g12, refg)


def test_3xtf32_split_is_fp32_grade():
    """The tensor-core backward (csrc/blend_slab_bwd_tc.cu) feeds mma.sync.m16n8k8 TF32 with a = a_hi + a_lo,
    b = b_hi + b_lo (hi = the value truncated to TF32's 10 mantissa bits, lo = the exact fp32 remainder, itself
    truncated by the MMA) and accumulates a_lo b_hi + a_hi b_lo + a_hi b_hi in fp32.  Emulated here in numpy: the
    dropped terms are bounded by 2^-21 |a||b| per product, i.e. the contraction over 32 pixels stays within ~1e-6 of the
    exact sum relative to sum |a||b| -- the same order as a plain fp32 FMA chain, which is why the parity tolerances did
    not move when the contractions went to the tensor cores."""
    rng = np.random.default_rng(0)

    def tf32(x):
        return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)

    worst = 0.0
    for scale in (1.0, 1e-3, 37.0):
        a = (rng.standard_normal((4096, 32)) * scale).astype(np.float32)
        b = rng.standard_normal((4096, 32)).astype(np.float32)
        a_hi, b_hi = tf32(a), tf32(b)
        a_lo, b_lo = tf32(a - a_hi), tf32(b - b_hi)  # a - a_hi is exact in fp32; the MMA truncates the operand again
        assert np.array_equal(a_hi + (a - a_hi), a)
        acc = np.zeros(4096, np.float32)
        for k in range(32):  # fp32 accumulation, products of TF32 operands are exact in fp32
            for x, y in ((a_lo, b_hi), (a_hi, b_lo), (a_hi, b_hi)):
                acc = (acc + x[:, k] * y[:, k]).astype(np.float32)
        exact = (a.astype(np.float64) * b.astype(np.float64)).sum(1)
        mass = (np.abs(a).astype(np.float64) * np.abs(b)).sum(1)
        worst = max(worst, float((np.abs(acc - exact) / mass).max()))
        plain = np.zeros(4096, np.float32)
        for k in range(32):
            plain = (plain + a[:, k] * b[:, k]).astype(np.float32)
        plain_err = float((np.abs(plain - exact) / mass).max())
        assert worst <= 4 * max(plain_err, 2.0 ** -22), (scale, worst, plain_err)
    assert worst <= 1.5e-6, worst
