"""oracle/roma_shim.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restatement of the four ``roma==1.5.0`` functions the reference's deformation
calls (``flow3d/scene_model.py:94-101``; roma is pinned at
``requirements.txt:385`` but is not installed here and cannot be fetched).
Quaternions are XYZW.  Follows roma's published implementation
(``rotmat_to_unitquat`` is the SciPy-style 4-branch construction, sign not
canonicalised).  Installing this module as
``sys.modules['roma']`` lets the reference's own ``flow3d/params.py``,
``flow3d/transforms.py`` and ``flow3d/scene_model.py`` import on CPU, which is
how ``tests/golden/make_golden.py`` pins the deformation oracle.
"""
from __future__ import annotations

import torch


def rotmat_to_unitquat(R: torch.Tensor) -> torch.Tensor:
    batch_shape = R.shape[:-2]
    matrix = R.reshape(-1, 3, 3)
    n = matrix.shape[0]
    decision = torch.empty((n, 4), dtype=matrix.dtype, device=matrix.device)
    decision[:, :3] = matrix.diagonal(dim1=1, dim2=2)
    decision[:, -1] = decision[:, :3].sum(dim=1)
    choices = decision.argmax(dim=1)
    quat = torch.empty((n, 4), dtype=matrix.dtype, device=matrix.device)
    ind = torch.nonzero(choices != 3, as_tuple=True)[0]
    i = choices[ind]
    j = (i + 1) % 3
    k = (j + 1) % 3
    quat[ind, i] = 1 - decision[ind, -1] + 2 * matrix[ind, i, i]
    quat[ind, j] = matrix[ind, j, i] + matrix[ind, i, j]
    quat[ind, k] = matrix[ind, k, i] + matrix[ind, i, k]
    quat[ind, 3] = matrix[ind, k, j] - matrix[ind, j, k]
    ind = torch.nonzero(choices == 3, as_tuple=True)[0]
    quat[ind, 0] = matrix[ind, 2, 1] - matrix[ind, 1, 2]
    quat[ind, 1] = matrix[ind, 0, 2] - matrix[ind, 2, 0]
    quat[ind, 2] = matrix[ind, 1, 0] - matrix[ind, 0, 1]
    quat[ind, 3] = 1 + decision[ind, -1]
    quat = quat / torch.norm(quat, dim=1)[:, None]
    return quat.reshape(batch_shape + (4,))


def quat_product(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    vector = (p[..., None, 3] * q[..., :3] + q[..., None, 3] * p[..., :3]
              + torch.cross(p[..., :3].expand_as(torch.broadcast_tensors(p[..., :3], q[..., :3])[0]),
                            q[..., :3].expand_as(torch.broadcast_tensors(p[..., :3], q[..., :3])[0]), dim=-1))
    last = p[..., 3] * q[..., 3] - torch.sum(p[..., :3] * q[..., :3], dim=-1)
    return torch.cat((vector, last[..., None]), dim=-1)


def quat_xyzw_to_wxyz(xyzw: torch.Tensor) -> torch.Tensor:
    return torch.roll(xyzw, 1, dims=-1)


def quat_wxyz_to_xyzw(wxyz: torch.Tensor) -> torch.Tensor:
    return torch.roll(wxyz, -1, dims=-1)
