"""Dev tool (GPU box): where do libfnx's pixels differ from the compiled reference at BASELINE size, with and without the
opacity-aware tile culling?"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fluidnexus_b200 import rasterizer as R  # noqa: E402
from fluidnexus_b200 import synthetic as S  # noqa: E402
from oracle import ref_ext  # noqa: E402
import test_baseline_sizes_gpu as T  # noqa: E402


def main():
    for name in sys.argv[1:] or ["c4", "c5"]:
        fluid, bg, C, size = T._sets(name)
        gs = fluid if bg is None else S.cat_sets(fluid, bg)
        cams = S.make_cameras(5, size)
        dL = torch.zeros((5, C, size, size), device="cuda")
        common, inp0, ref, _ = T._reference_views(C, gs, cams, dL)
        view_all = torch.stack([c.world_view_transform for c in cams]).cuda().contiguous()
        proj_all = torch.stack([c.full_proj_transform for c in cams]).cuda().contiguous()
        for exact in (True, False):
            ctx, col, rad, dep = R.raster_forward(C, *common, 1.0, None, view_all, proj_all, inp0["tan_fov_x"], inp0["tan_fov_y"], size, size,
                                                  speculative=False, exact_rect=exact)
            st = R.read_image_state(ctx)
            for v in range(5):
                d = (col[v] - ref[v]["color"]).abs().amax(0)
                n5, n4 = int((d > 1e-5).sum()), int((d > 1e-4).sum())
                iy, ix = np.unravel_index(int(d.argmax()), d.shape)
                print(f"{name} exact_rect={exact} view {v}: max|d| {float(d.max()):.3e} at (x={ix}, y={iy}) pixels>1e-5: {n5} >1e-4: {n4}; "
                      f"fnx {col[v, :, iy, ix].tolist()} ref {ref[v]['color'][:, iy, ix].tolist()} final_T {float(st['final_T'][v, iy, ix]):.4e} "
                      f"n_contrib {int(st['n_contrib'][v, iy, ix])}", flush=True)


if __name__ == "__main__":
    main()
