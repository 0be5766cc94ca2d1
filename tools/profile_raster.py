"""Run a few forward+backward passes of one scene (for ncu launch lists / full captures)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fluidnexus_b200 import rasterizer as R  # noqa: E402
from fluidnexus_b200 import synthetic as S  # noqa: E402


def main():
    C = int(os.environ.get("FNX_C", 3))
    nf, nb = int(os.environ.get("FNX_NF", 20000)), int(os.environ.get("FNX_NB", 180000))
    size, V = int(os.environ.get("FNX_SIZE", 512)), int(os.environ.get("FNX_V", 1))
    iters = int(os.environ.get("FNX_ITERS", 3))
    gs = S.fluid_gaussians(nf, C, seed=0)
    if nb:
        gs = S.cat_sets(gs, S.background_gaussians(nb, C, seed=1))
    cams = S.make_cameras(5, size)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    inps = [S.raster_inputs(gs, c, np.zeros(C, np.float32)) for c in cams[:V]]
    inp = inps[0]
    view = torch.stack([t(i["view"]) for i in inps]) if V > 1 else t(inp["view"])
    proj = torch.stack([t(i["proj"]) for i in inps]) if V > 1 else t(inp["proj"])
    a = [t(inp[k]) for k in ("bg", "means3D", "colors", "opacities", "scales", "rotations")]
    for it in range(iters):
        ctx, col, rad, dep = R.raster_forward(C, a[0], a[1], a[2], a[3], a[4], a[5], 1.0, None, view, proj, inp["tan_fov_x"],
                                              inp["tan_fov_y"], inp["H"], inp["W"], speculative=True)
        dL = torch.ones_like(col)
        g = R.raster_backward(ctx, dL)
    torch.cuda.synchronize()
    print("R", ctx.num_rendered, "visible", int((rad > 0).sum()))


if __name__ == "__main__":
    main()
