"""``gsplat.rendering`` as seen by flow3d/scene_model.py:5 -- forwards to deblur4dgs_b200."""
from deblur4dgs_b200.rendering import (  # noqa: F401
    fully_fused_projection,
    isect_offset_encode,
    isect_tiles,
    rasterization,
    rasterize_to_pixels,
)
