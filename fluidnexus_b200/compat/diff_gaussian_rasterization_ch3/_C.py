"""Drop-in for the reference's pybind module `diff_gaussian_rasterization_ch3._C` (R3/ext.cpp:15-19, signatures
R3/rasterize_points.h:18-64), backed by libfnx: the level at which the reference itself binds its native code.  The reference's
own wrapper package (`__init__.py`: `from . import _C`) runs unchanged on top of it (tests/test_raster_gpu.py)."""
from fluidnexus_b200.rasterizer import make_C as _make

rasterize_gaussians, rasterize_gaussians_backward, mark_visible = _make(3)
__all__ = ["rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"]
