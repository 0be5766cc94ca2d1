"""Generate tests/golden/ref_*.npz by running the COMPILED, UNMODIFIED REFERENCE rasterizer on a GPU.

    gpurun -- 'python tools/make_golden.py gpurun_out/golden'   then copy gpurun_out/golden/*.npz to tests/golden/

Inputs are the seeded scenes of tests/scenes.py (regenerated from the seed by the tests, so only outputs are
stored): image, median depth, radii, instance count, the reference's own intermediate state (means2D, depths,
conic_opacity carved out of its geomBuffer following R3/cuda_rasterizer/rasterizer_impl.cu:144-160) and the 7
gradient tensors for the seeded dL/dpixel of tests/scenes.py:dL_dpix.  Also records run-to-run gradient noise of
the reference (float atomics).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scenes  # noqa: E402
from oracle.ref_ext import RefRaster  # noqa: E402


from oracle.ref_ext import carve_geom  # noqa: E402,F401


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    dev = "cuda"
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    for name in scenes.SCENES:
        gs, cam, bg, inp = scenes.build(name)
        C = inp["colors"].shape[1]
        rr = RefRaster(C)
        args = (t(inp["bg"]), t(inp["means3D"]), t(inp["colors"]), t(inp["opacities"]), t(inp["scales"]), t(inp["rotations"]),
                1.0, t(inp["view"]), t(inp["proj"]), inp["tan_fov_x"], inp["tan_fov_y"], inp["H"], inp["W"])
        out = rr.forward(*args)
        torch.cuda.synchronize()
        P = inp["means3D"].shape[0]
        depths, m2, co = carve_geom(out["geom"], P)
        dL = t(scenes.dL_dpix(name, tuple(out["color"].shape)))
        g = rr.backward(dL)
        g2 = rr.backward(dL)
        torch.cuda.synchronize()
        noise = {k: float((g[k] - g2[k]).norm() / (g[k].norm() + 1e-30)) for k in g}
        save = dict(scene=name, color=out["color"].cpu().numpy(), depth=out["depth"].cpu().numpy(),
                    radii=out["radii"].cpu().numpy(), num_rendered=np.int64(out["num_rendered"]),
                    ref_depths=depths, ref_means2D=m2, ref_conic_opacity=co,
                    noise=np.array([noise[k] for k in sorted(noise)]))
        for k, v in g.items():
            save["g_" + k] = v.cpu().numpy()
        np.savez_compressed(os.path.join(outdir, f"ref_{name}.npz"), **save)
        print(name, "R", out["num_rendered"], "visible", int((out["radii"] > 0).sum()), "grad noise", noise)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
