import os, sys, time
sys.path.insert(0, '/root/repo' if os.path.isdir('/root/repo/fluidnexus_b200') else '.')
import torch, bench
from fluidnexus_b200 import rasterizer as R
from fluidnexus_b200.step import FrameState, PhysicalStep, StepParams
dev = torch.device("cuda", 0)
cams, bg, frames, cfg = bench.build_frames("smoke", 1, dev)
prm = StepParams(p0=cfg["p0"], grey=cfg["grey"], distance_threshold_visual=cfg["thr"])
ps = PhysicalStep(cams, cfg["C"], prm, device=dev)
fr = FrameState(frames[0]["hidden"], frames[0]["visual"], frames[0]["fluid"], bg, device=dev, prm=prm)
gt = torch.rand(5, 3, 512, 512, device=dev) * 0.5
views=[0,1,2,3,4]
for _ in range(3): ps.step(fr, views, gt)
torch.cuda.synchronize()
T = {}
def tm(name, fn):
    torch.cuda.synchronize(); t=time.perf_counter(); r=fn(); T[name]=T.get(name,0)+time.perf_counter()-t; torch.cuda.synchronize(); return r
n=10
for _ in range(n):
    tm('physics_forward', lambda: ps.physics_forward(fr))
    ctx, images, radii, depth = tm('render', lambda: ps.render(fr, views))
    l1, ss, g = tm('image_loss', lambda: ps.image_loss(images, gt, 5))
    grads = tm('raster_backward', lambda: R.raster_backward(ctx, g, want_means2D=False))
    tm('physics_backward', lambda: ps.physics_backward_and_update(fr, grads["means3D"]))
for k,v in T.items(): print(f"{k:18s} host-side {v/n*1e3:.3f} ms (GPU idle at call time; includes the in-call sync for render)")
