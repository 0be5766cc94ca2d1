"""Drop-in for the one `torch_scatter` function FluidNexus imports (FD/gaussian_splatting/gm_fluid.py:10:
`from torch_scatter import scatter_min`), backed by libfnx."""
from fluidnexus_b200.physics import scatter_min

__all__ = ["scatter_min"]
