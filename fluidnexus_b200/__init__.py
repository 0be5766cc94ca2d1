"""fluidnexus_b200 -- B200-native hot path of FluidNexus' FluidDynamics stage.

The product is `libfnx.so` (hand-written sm_100a CUDA behind the C ABI in include/fnx.h) plus this thin host
layer that mirrors the reference's plugin interfaces.  There is no CPU fallback anywhere in this package.
"""
import os
import sys

__version__ = "0.1.0"

COMPAT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat")


def install_compat():
    """Make `diff_gaussian_rasterization_ch1/_ch3`, `simple_knn`, `torch_cluster`, `torch_scatter` resolve to the
    libfnx-backed drop-ins (the five imports that form the reference's plugin boundary)."""
    if COMPAT_DIR not in sys.path:
        sys.path.insert(0, COMPAT_DIR)
