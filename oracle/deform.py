"""oracle/deform.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (torch, fp32 or fp64, differentiable through autograd) restatement of the
reference's motion-basis deformation, SURVEY.md section 8 rows a1-a6:

* activations                      flow3d/params.py:39-43, 70-84
* MotionBases.compute_transforms   flow3d/params.py:142-180
* cont_6d_to_rmat                  flow3d/transforms.py:41-53
* SceneModel.compute_poses_fg/all  flow3d/scene_model.py:76-120
* camera sub-exposure transform    flow3d/scene_model.py:352-353

Pinned: ``tests/golden/make_golden.py`` runs the reference's OWN
``flow3d/params.py`` / ``flow3d/transforms.py`` / ``SceneModel.compute_poses_*``
code (imported from /root/reference with ``oracle/roma_shim.py`` standing in
for the absent ``roma``) and commits its outputs and autograd gradients under
``tests/golden/deform_*.npz``; ``tests/test_oracle.py`` checks this
restatement against those fixtures.  The roma arithmetic itself (rotation
matrix -> quaternion, quaternion product) is a restatement of an absent
third-party package: that part is "parity unpinned".

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import roma_shim as roma


def cont_6d_to_rmat(cont_6d):
    """transforms.py:41-53 -- Gram-Schmidt, columns x, y, z."""
    a, b = cont_6d[..., 0:3], cont_6d[..., 3:6]
    x = F.normalize(a, dim=-1)
    y = F.normalize(b - (b * x).sum(dim=-1, keepdim=True) * x, dim=-1)
    z = torch.linalg.cross(x, y, dim=-1)
    return torch.stack([x, y, z], dim=-1)


def compute_transforms(ts, coefs, rots, transls):
    """params.py:142-180.  ts [B] or [1,B]; coefs [G,K] (already softmaxed);
    rots [K,T,6]; transls [K,T,3] -> [G,B,3,4]."""
    if ts.dim() == 1:
        ts = ts.unsqueeze(0)
    T = transls.shape[1]
    ts_pre = torch.floor(ts).clamp(0.0, T - 1).int()
    ts_next = torch.ceil(ts).clamp(0.0, T - 1).int()
    ip, inx = ts_pre[0].long(), ts_next[0].long()
    tp = torch.einsum("pk,kni->pni", coefs, transls[:, ip])
    rp = torch.einsum("pk,kni->pni", coefs, rots[:, ip])
    tn = torch.einsum("pk,kni->pni", coefs, transls[:, inx])
    rn = torch.einsum("pk,kni->pni", coefs, rots[:, inx])
    w = (ts - ts_pre).to(coefs.dtype)[..., None]  # [1,B,1]
    t = (1.0 - w) * tp + w * tn
    r = (1.0 - w) * rp + w * rn
    return torch.cat([cont_6d_to_rmat(r), t[..., None]], dim=-1)


def compute_poses_fg(fg_means, fg_quats_raw, motion_coefs_raw, rots, transls, ts):
    """scene_model.py:76-106 -> means [Gf,B,3], quats [Gf,B,4] (wxyz)."""
    quats = F.normalize(fg_quats_raw, dim=-1, p=2)
    coefs = F.softmax(motion_coefs_raw, dim=-1)
    transfms = compute_transforms(ts, coefs, rots, transls)
    means = torch.einsum("pnij,pj->pni", transfms, F.pad(fg_means, (0, 1), value=1.0))
    q = roma.quat_xyzw_to_wxyz(roma.quat_product(roma.rotmat_to_unitquat(transfms[..., :3, :3]),
                                                 roma.quat_wxyz_to_xyzw(quats[:, None])))
    return means, F.normalize(q, p=2, dim=-1)


def compute_poses_all(fg_means, fg_quats_raw, motion_coefs_raw, bg_means, bg_quats_raw, rots, transls, ts):
    """scene_model.py:108-120 -- fg (deformed) first, bg (static) after."""
    means, quats = compute_poses_fg(fg_means, fg_quats_raw, motion_coefs_raw, rots, transls, ts)
    B = means.shape[1]
    bq = F.normalize(bg_quats_raw, dim=-1, p=2)
    means = torch.cat([means, bg_means[:, None].expand(-1, B, -1)], dim=0)
    quats = torch.cat([quats, bq[:, None].expand(-1, B, -1)], dim=0)
    return means, quats


def deform_subexposures(fg_means, fg_quats_raw, motion_coefs_raw, bg_means, bg_quats_raw, rots, transls,
                        times, RTs):
    """The per-sub-exposure loop body of SceneModel.render (scene_model.py:323-353)
    for all N sub-exposures: times [N], RTs [N,3,4] -> means [N,G,3], quats [N,G,4]."""
    outs_m, outs_q = [], []
    for ii in range(times.shape[0]):
        time = times[None, ii:ii + 1]  # [1,1] like `times[:, ii:ii+1]`
        m, q = compute_poses_all(fg_means, fg_quats_raw, motion_coefs_raw, bg_means, bg_quats_raw, rots,
                                 transls, time)
        m, q = m[:, 0], q[:, 0]
        transR, transT = RTs[ii][:3, :3], RTs[ii][:3, 3:4]
        m = (transR @ m.permute(1, 0) + transT).permute(1, 0)
        outs_m.append(m)
        outs_q.append(q)
    return torch.stack(outs_m), torch.stack(outs_q)
