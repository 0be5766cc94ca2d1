"""Drop-in for `simple_knn._C` (`from simple_knn._C import distCUDA2`, FD/gaussian_splatting/gm_fluid.py:7;
binding KNN/ext.cpp, signature KNN/spatial.h:14), backed by libfnx's grid search."""
from fluidnexus_b200.physics import distCUDA2

__all__ = ["distCUDA2"]
