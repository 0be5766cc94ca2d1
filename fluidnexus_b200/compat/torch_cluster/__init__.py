"""Drop-in for the two `torch_cluster` functions FluidNexus imports (FD/gaussian_splatting/gm_fluid.py:9:
`from torch_cluster import radius, radius_graph`), backed by libfnx's grid-hash search."""
from fluidnexus_b200.physics import radius, radius_graph

__all__ = ["radius", "radius_graph"]
