"""Drop-in for the reference package `diff_gaussian_rasterization_ch3`
(FluidDynamics/submodules/gaussian_rasterization_ch3/diff_gaussian_rasterization_ch3/__init__.py), backed by libfnx.

    from diff_gaussian_rasterization_ch3 import GaussianRasterizationSettings, GaussianRasterizer

is what FD/helpers/helper_pipe.py:14-41 imports; put `fluidnexus_b200/compat` on sys.path (or call
`fluidnexus_b200.install_compat()`) and the reference's renderer/ and entries_* run unchanged.
"""
from fluidnexus_b200.rasterizer import make_module as _make

GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians, _RasterizeGaussians = _make(3)
__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
