"""deblur4dgs_b200 -- B200 (sm_100a) implementation of the Deblur4DGS per-frame render hot path.

Public surface (mirrors what ``flow3d/`` calls):

* ``rendering.rasterization``        drop-in for ``gsplat.rendering.rasterization`` (gsplat==1.1.1)
* ``motion.deform_subexposures``     fused motion-basis deformation at N sub-exposure timestamps
* ``motion.compute_poses_fg/all``    reference-named wrappers (``[G,B,*]`` layout)
* ``scene.render_subexposures``      the whole loop of ``SceneModel.render`` + N-way combine
* ``synthetic``                      seeded synthetic scenes (BASELINE.json configs)
* ``parallel``                       sub-exposure sharding over ranks (NCCL / gloo)

All compute goes through the C ABI of ``libd4gs.so`` (``include/d4gs.h``); there is no CPU
or eager fallback -- ops raise if the library is missing or tensors are not on a CUDA device.
"""
from . import _cabi  # noqa: F401

__all__ = ["rasterization", "deform_subexposures", "render_subexposures"]


def __getattr__(name):
    if name == "rasterization":
        from .rendering import rasterization
        return rasterization
    if name == "deform_subexposures":
        from .motion import deform_subexposures
        return deform_subexposures
    if name == "render_subexposures":
        from .scene import render_subexposures
        return render_subexposures
    raise AttributeError(name)
