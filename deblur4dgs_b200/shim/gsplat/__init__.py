"""Import shim: put ``deblur4dgs_b200/shim`` on PYTHONPATH and the reference's
``from gsplat.rendering import rasterization`` (flow3d/scene_model.py:5) resolves to the
B200 implementation with no source change.  See INTEGRATION.md."""
from . import rendering  # noqa: F401

__version__ = "1.1.1+d4gs.b200"
