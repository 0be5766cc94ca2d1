"""Dev tool: per-section GPU time vs host wall time of the fused training iteration (one frame, 5 views)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fluidnexus_b200 import _lib as L  # noqa: E402
from fluidnexus_b200.step import FrameState, PhysicalStep, StepParams  # noqa: E402


def main():
    wl = os.environ.get("FNX_WORKLOAD", "smoke")
    iters = int(os.environ.get("FNX_ITERS", 20))
    dev = torch.device("cuda", 0)
    cams, bg, frames, cfg = bench.build_frames(wl, 1, dev)
    prm = StepParams(p0=cfg["p0"], buoyancy_max_y=cfg["bmax"], grey=cfg["grey"], distance_threshold_visual=cfg["thr"])
    ps = PhysicalStep(cams, cfg["C"], prm, device=dev)
    fr = FrameState(frames[0]["hidden"], frames[0]["visual"], frames[0]["fluid"], bg, device=dev, prm=prm)
    gt = torch.rand(5, cfg["C"], cfg["size"], cfg["size"], device=dev) * 0.5
    lib = L.lib()
    views = [0, 1, 2, 3, 4]
    for _ in range(5):
        ps.step(fr, views, gt)
    torch.cuda.synchronize()
    nsec = lib.fnx_profile_sections()
    names = [lib.fnx_profile_section_name(i).decode() for i in range(nsec)]
    for mask, label in ((0, "profiler off"), ((1 << nsec) - 1, "profiler on")):
        lib.fnx_profile_enable(mask)
        lib.fnx_profile_collect(None, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(iters):
            out = ps.step(fr, views, gt)
        e1.record()
        t_queue = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t0
        print(f"[{label}] per iteration: host queue {t_queue / iters * 1e3:.3f} ms, wall {t_wall / iters * 1e3:.3f} ms, "
              f"GPU events {e0.elapsed_time(e1) / iters:.3f} ms, instances {out['ws'].num_rendered()}")
        tot = (C.c_float * nsec)(); cnt = (C.c_int32 * nsec)()
        lib.fnx_profile_collect(tot, cnt)
        if mask:
            s = 0.0
            for i in range(nsec):
                if cnt[i]:
                    print(f"   {names[i]:12s} {tot[i] / iters:8.3f} ms/iter  ({cnt[i] // iters} launches)")
                    s += tot[i] / iters
            print(f"   sum of sections {s:.3f} ms/iter")
    lib.fnx_profile_enable(0)


if __name__ == "__main__" and not os.environ.get("FNX_TILE_STATS"):
    main()


def tile_stats():
    """Distribution of the per-tile / per-patch span lengths of the last forward (what bounds the blend kernels' tail)."""
    wl = os.environ.get("FNX_WORKLOAD", "smoke")
    dev = torch.device("cuda", 0)
    cams, bg, frames, cfg = bench.build_frames(wl, 1, dev)
    prm = StepParams(p0=cfg["p0"], buoyancy_max_y=cfg["bmax"], grey=cfg["grey"], distance_threshold_visual=cfg["thr"])
    ps = PhysicalStep(cams, cfg["C"], prm, device=dev)
    fr = FrameState(frames[0]["hidden"], frames[0]["visual"], frames[0]["fluid"], bg, device=dev, prm=prm)
    gt = torch.rand(5, cfg["C"], cfg["size"], cfg["size"], device=dev) * 0.5
    for _ in range(3):
        out = ps.step(fr, [0, 1, 2, 3, 4], gt)
    ts = out["ws"].tile_state()
    dyn = ts["tile_src"] == 0
    q = lambda a: " ".join(f"{np.percentile(a, p):.0f}" for p in (50, 90, 99, 100))
    print("dynamic tiles", int(dyn.sum()), "of", dyn.size)
    print("merged span length (p50 p90 p99 max):", q((ts["end"] - ts["begin"])[dyn]))
    print("tile_last fwd:", q(ts["tile_last"][dyn]), " sum", int(ts["tile_last"][dyn].sum()))
    print("patch_last fwd:", q(ts["patch_last"][dyn].reshape(-1)))
    print("dyn_last (bwd start):", q(ts["tile_dyn_last"][dyn]), " sum", int(np.minimum(ts["tile_last"], ts["tile_dyn_last"])[dyn].sum()))
    nstat = ts["tile_last"][~dyn]
    print("static-only tiles tile_last:", q(nstat))


if os.environ.get("FNX_TILE_STATS"):
    tile_stats()
