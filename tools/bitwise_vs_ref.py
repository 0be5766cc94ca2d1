"""Dev tool (GPU box): bit-level comparison of libfnx's intermediate state with the compiled reference on larger
scenes, plus a first timing of both.  Writes a text report to gpurun_out/bitwise_vs_ref.txt."""
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from fluidnexus_b200 import rasterizer as R  # noqa: E402
from fluidnexus_b200 import synthetic as S  # noqa: E402
from make_golden import carve_geom  # noqa: E402
from oracle.ref_ext import RefRaster  # noqa: E402


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    lines = []

    def log(*a):
        s = " ".join(str(x) for x in a)
        print(s, flush=True)
        lines.append(s)

    for C, nf, nb, size in [(3, 20000, 180000, 512), (1, 150000, 0, 512), (3, 50000, 0, 400)]:
        gs = S.fluid_gaussians(nf, C, seed=0)
        if nb:
            gs = S.cat_sets(gs, S.background_gaussians(nb, C, seed=1))
        cam = S.make_cameras(5, size)[1]
        inp = S.raster_inputs(gs, cam, np.zeros(C, np.float32))
        P = gs.P
        args = (t(inp["bg"]), t(inp["means3D"]), t(inp["colors"]), t(inp["opacities"]), t(inp["scales"]), t(inp["rotations"]),
                1.0, t(inp["view"]), t(inp["proj"]), inp["tan_fov_x"], inp["tan_fov_y"], inp["H"], inp["W"])
        rr = RefRaster(C)
        ro = rr.forward(*args)
        rd, rm2, rco = carve_geom(ro["geom"], P)
        for exact in (True, False):
            ctx, col, rad, dep = R.raster_forward(C, args[0], args[1], args[2], args[3], args[4], args[5], 1.0, None, args[7], args[8],
                                                  args[9], args[10], args[11], args[12], exact_rect=exact, speculative=False)
            g = R.read_geom(ctx)
            vis = (ro["radii"] > 0).cpu().numpy()
            xy = g["xy"][0].cpu().numpy(); de = g["depth"][0].cpu().numpy(); co = g["conic_opacity"][0].cpu().numpy()
            log(f"--- C={C} P={P} size={size} exact_rect={exact}: R_ref={ro['num_rendered']} R_fnx={ctx.num_rendered} visible={vis.sum()}")
            log("   radii mismatches", int((rad != ro["radii"]).sum().item()),
                "| xy bit mismatches", int((xy[vis].view(np.uint32) != rm2[vis].view(np.uint32)).sum()),
                "| depth bit mismatches", int((de[vis].view(np.uint32) != rd[vis].view(np.uint32)).sum()),
                "| conic bit mismatches", int((co[vis].view(np.uint32) != rco[vis].view(np.uint32)).sum()))
            d = (col - ro["color"]).abs()
            log("   image max|d|", float(d.max()), "pixels > 1e-5:", int((d > 1e-5).sum()), " > 1e-3:", int((d > 1e-3).sum()),
                "| depth mismatches", int((dep != ro["depth"]).sum()))
            dL = torch.randn(col.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
            gr = rr.backward(dL)
            gf = R.raster_backward(ctx, dL)
            for k in ("means2D", "colors", "opacity", "means3D", "scales", "rotations"):
                a, b = gf[k].reshape(gr[k].shape), gr[k]
                log(f"   grad {k:10s} rel-L2 {float((a - b).norm() / (b.norm() + 1e-30)):.3e}  max|d| {float((a - b).abs().max()):.3e}")
        # timings
        t_ref_f = timeit(lambda: rr.forward(*args))
        dLr = torch.randn(ro["color"].shape, device="cuda")
        t_ref_b = timeit(lambda: rr.backward(dLr))
        for exact in (True, False):
            for spec in (False, True):
                f = lambda: R.raster_forward(C, args[0], args[1], args[2], args[3], args[4], args[5], 1.0, None, args[7], args[8],
                                             args[9], args[10], args[11], args[12], exact_rect=exact, speculative=spec)
                t_f = timeit(lambda: f())
                ctx = f()[0]
                t_b = timeit(lambda: R.raster_backward(ctx, dLr))
                log(f"   time ms: ref fwd {t_ref_f:.3f} bwd {t_ref_b:.3f} | fnx(exact_rect={exact}, speculative={spec}) fwd {t_f:.3f} bwd {t_b:.3f}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "bitwise_vs_ref.txt"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
