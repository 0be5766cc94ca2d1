"""Dev tool: time of one static-background training iteration (fluidnexus_b200.background.BackgroundStep) at P = 200k, one
512x512 view, next to the same iteration written the reference's way (oracle/background_ref.py: torch activations + autograd
+ torch.optim.Adam around the libfnx drop-in rasterizer and the fused loss module)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fluidnexus_b200 import losses as FL  # noqa: E402
from fluidnexus_b200 import rasterizer as R  # noqa: E402
from fluidnexus_b200 import synthetic as S  # noqa: E402
from fluidnexus_b200.background import BackgroundModel, BackgroundStep  # noqa: E402
from oracle import background_ref as OB  # noqa: E402


class Args:
    position_lr_init, position_lr_final, position_lr_delay_mult, position_lr_max_steps = 1.6e-4, 1.6e-6, 0.01, 30_000
    color_lr, opacity_lr, scaling_lr, rotation_lr, percent_dense = 2.5e-3, 0.05, 5e-3, 1e-3, 0.01


def timeit(fn, n=30, warm=5):
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
    P = int(os.environ.get("FNX_P", 200_000))
    g = S.background_gaussians(P, 3, seed=1).torch("cuda")
    cams = S.make_cameras(5, 512, device="cuda")
    bg = torch.zeros(3, device="cuda")
    gt = torch.rand(3, 512, 512, device="cuda")
    mine = BackgroundModel(g["xyz"], g["colors"], g["opacity"], g["scales"], g["rotations"])
    mine.training_setup(Args)
    step = BackgroundStep(3, bg_color=bg)
    k = [0]

    def fused():
        k[0] += 1
        step.step(mine, cams[k[0] % 5], gt)

    ref = OB.RefBackgroundModel(g["xyz"], g["colors"], g["opacity"], g["scales"], g["rotations"])
    ref.training_setup(Args)
    Settings, Rasterizer, _, _ = R.make_module(3)

    def reference_style():
        k[0] += 1
        OB.ref_iteration(ref, cams[k[0] % 5], gt, bg, Settings, Rasterizer, FL.l1_loss, FL.ssim)

    t_f, t_r = timeit(fused), timeit(reference_style)
    print(f"background iteration, P={P}, 1 view 512x512: fused BackgroundStep {t_f:.3f} ms | torch activations/autograd/Adam around the "
          f"same rasterizer {t_r:.3f} ms ({t_r / t_f:.2f}x)")


if __name__ == "__main__":
    main()
