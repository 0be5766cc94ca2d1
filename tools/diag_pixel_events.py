"""Dev tool (GPU box): for given (scene, view, x, y) list the blend events that sit on a decision threshold."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fluidnexus_b200 import synthetic as S  # noqa: E402
from oracle import ref_ext  # noqa: E402
import test_baseline_sizes_gpu as T  # noqa: E402


def main():
    cases = [("c5", 0, 40, 439), ("c4", 2, 469, 29), ("c4", 0, 173, 56)]
    cache = {}
    for name, v, px, py in cases:
        if name not in cache:
            fluid, bg, C, size = T._sets(name)
            gs = fluid if bg is None else S.cat_sets(fluid, bg)
            cams = S.make_cameras(5, size)
            dL = torch.zeros((5, C, size, size), device="cuda")
            cache[name] = (gs, T._reference_views(C, gs, cams, dL), size)
        gs, (common, inp0, ref, _), size = cache[name]
        r = ref[v]
        rad = r["radii"].cpu().numpy()
        xy, co, dep = r["xy"].astype(np.float32), r["conic"].astype(np.float32), r["gdepth"]
        col = gs.colors.astype(np.float32)
        vis = rad > 0
        # rect test (auxiliary.h getRect) for this pixel's tile
        tx, ty = px // 16, py // 16
        grid = (size + 15) // 16
        rmin_x = np.clip(((xy[:, 0] - rad) / 16).astype(np.int32), 0, grid); rmax_x = np.clip(((xy[:, 0] + rad + 15) / 16).astype(np.int32), 0, grid)
        rmin_y = np.clip(((xy[:, 1] - rad) / 16).astype(np.int32), 0, grid); rmax_y = np.clip(((xy[:, 1] + rad + 15) / 16).astype(np.int32), 0, grid)
        inr = vis & (rmin_x <= tx) & (tx < rmax_x) & (rmin_y <= ty) & (ty < rmax_y)
        idx = np.nonzero(inr)[0]
        order = np.lexsort((idx, dep[idx].view(np.uint32)))
        idx = idx[order]
        f = np.float32
        dx = xy[idx, 0] - f(px); dy = xy[idx, 1] - f(py)
        a, b, c, o = co[idx, 0], co[idx, 1], co[idx, 2], co[idx, 3]
        # unfused fp32 and fp64 powers
        p32 = f(-0.5) * (a * dx * dx + c * dy * dy) - b * dx * dy
        p64 = -0.5 * (a.astype(np.float64) * dx.astype(np.float64) ** 2 + c.astype(np.float64) * dy.astype(np.float64) ** 2) - b.astype(np.float64) * dx * dy
        al64 = np.minimum(0.99, o.astype(np.float64) * np.exp(p64))
        Tt = 1.0
        Cacc = np.zeros(col.shape[1])
        print(f"== {name} view {v} pixel ({px},{py}): {idx.size} instances in the tile; ref {r['color'][:, py, px].tolist()}")
        ties = int((np.diff(dep[idx].view(np.uint32).astype(np.int64)) == 0).sum())
        print("   depth ties among the tile's instances:", ties)
        for k in range(idx.size):
            if p64[k] > 0:
                if abs(p64[k]) < 1e-5:
                    print(f"   k={k} power ~ 0: {p64[k]:.3e} (p32 {p32[k]:.3e})")
                continue
            al = al64[k]
            if abs(al - 1 / 255) < 2e-7:
                print(f"   k={k} g={idx[k]} alpha on the 1/255 cut: alpha64 {al:.10f} (1/255 = {1/255:.10f}), rel {abs(al*255-1):.2e}, T={Tt:.4e}, col {col[idx[k]].tolist()}")
            if al < 1 / 255:
                continue
            tT = Tt * (1 - al)
            if abs(tT - 1e-4) < 1e-8:
                print(f"   k={k} g={idx[k]} termination test on the threshold: T(1-a) = {tT:.10e}, alpha {al:.4f}, contribution {al*Tt:.3e}")
            if tT < 1e-4:
                print(f"   terminated at k={k} with T={Tt:.4e}")
                break
            Cacc += col[idx[k]] * al * Tt
            if Tt > 0.5 and tT < 0.5 and abs(tT - 0.5) < 1e-5:
                print(f"   k={k} median-depth crossing within 1e-5: {tT:.8f}")
            Tt = tT
        print("   fp64 emulation colour:", Cacc.tolist(), "final T", Tt)


if __name__ == "__main__":
    main()
