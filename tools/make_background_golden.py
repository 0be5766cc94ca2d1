"""Run in the build container (needs /root/reference): drives the reference's OWN gaussian_splatting/gm_background.GaussianModel
on the CPU (plyfile / simple_knn stubbed, device="cuda" redirected to the CPU) through a short training history -- Adam steps
with seeded gradients and the script's statistics updates, densify_and_prune, reset_opacity, one more step -- and stores the
inputs and every intermediate state in tests/golden/pyref_background.npz.  tests/test_reference_background_golden.py replays
the same history through oracle/background_ref.py (which the GPU tests compare the fused kernels with)."""
import os
import sys
import types

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/FluidDynamics"
OUT = os.path.join(ROOT, "tests", "golden", "pyref_background.npz")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from fluidnexus_b200 import synthetic as S  # noqa: E402
from make_physics_golden import cuda_as_cpu  # noqa: E402

NAMES = ("xyz", "color", "opacity", "scaling", "rotation")


class Args:
    position_lr_init, position_lr_final, position_lr_delay_mult, position_lr_max_steps = 1.6e-4, 1.6e-6, 0.01, 30_000
    color_lr, opacity_lr, scaling_lr, rotation_lr, percent_dense = 2.5e-3, 0.05, 5e-3, 1e-3, 0.01


def history(P, seed):
    """Seeded per-iteration inputs: gradients of the five tensors, the screen-space gradient, radii."""
    rng = np.random.default_rng(seed)
    steps = []
    for it in range(4):
        steps.append(dict(g_xyz=rng.normal(0, 1e-3, (P, 3)), g_color=rng.normal(0, 1e-2, (P, 3)), g_opacity=rng.normal(0, 1e-2, (P, 1)),
                          g_scaling=rng.normal(0, 1e-2, (P, 3)), g_rotation=rng.normal(0, 1e-2, (P, 4)),
                          g_screen=rng.normal(0, 3e-4, (P, 3)), radii=rng.integers(0, 40, P)))
    return steps


def snapshot(gm, prefix, out):
    for k in NAMES:
        p = getattr(gm, "_" + k)
        out[f"{prefix}_{k}"] = p.detach().numpy().copy()
        st = gm.optimizer.state.get(p, None)
        if st is not None and "exp_avg" in st:
            out[f"{prefix}_m_{k}"] = st["exp_avg"].numpy().copy()
            out[f"{prefix}_v_{k}"] = st["exp_avg_sq"].numpy().copy()
    out[f"{prefix}_accum"], out[f"{prefix}_denom"] = gm.xyz_gradient_accum.numpy().copy(), gm.denom.numpy().copy()
    out[f"{prefix}_max_radii"] = gm.max_radii2D.numpy().copy()


def apply_step(gm, h, it):
    """One pass of train_background.py:203-262 with the rasterizer's gradients replaced by the recorded ones."""
    gm.update_learning_rate(it)
    for k in NAMES:
        getattr(gm, "_" + k).grad = torch.tensor(h["g_" + k], dtype=torch.float32)
    screen = torch.zeros((gm.get_xyz.shape[0], 3), requires_grad=True)
    screen.grad = torch.tensor(h["g_screen"], dtype=torch.float32)
    radii = torch.tensor(h["radii"], dtype=torch.int32)
    vis = radii > 0
    with torch.no_grad():
        gm.max_radii2D[vis] = torch.max(gm.max_radii2D[vis], radii[vis])
        gm.add_densification_stats(screen, vis)
        gm.optimizer.step()
        gm.optimizer.zero_grad(set_to_none=True)


def main():
    ply = types.ModuleType("plyfile")
    ply.PlyData, ply.PlyElement = object, object
    sk, skc = types.ModuleType("simple_knn"), types.ModuleType("simple_knn._C")
    skc.distCUDA2 = lambda pts: None
    sys.modules.update({"plyfile": ply, "simple_knn": sk, "simple_knn._C": skc})
    sys.path.insert(0, REF)
    from gaussian_splatting.gm_background import GaussianModel
    P = 400
    g = S.background_gaussians(P, 3, seed=4)
    out = dict(init_xyz=g.xyz, init_color=g.colors, init_opacity=g.opacity, init_scales=g.scales, init_rotations=g.rotations * 1.3)
    gm = GaussianModel()
    f = lambda a: nn.Parameter(torch.tensor(np.asarray(a), dtype=torch.float32).requires_grad_(True))
    gm._xyz, gm._color, gm._rotation = f(g.xyz), f(g.colors), f(g.rotations * 1.3)
    gm._opacity = nn.Parameter(torch.log(torch.tensor(g.opacity, dtype=torch.float32) / (1 - torch.tensor(g.opacity, dtype=torch.float32))).requires_grad_(True))
    gm._scaling = nn.Parameter(torch.log(torch.tensor(g.scales, dtype=torch.float32)).requires_grad_(True))
    gm.max_radii2D = torch.zeros(P)
    gm.spatial_lr_scale = 5.0
    hist = history(P, seed=2)
    with cuda_as_cpu():
        gm.training_setup(Args)
        for it in range(1, 4):
            apply_step(gm, hist[it - 1], it)
        snapshot(gm, "after3", out)
        thr = float((gm.xyz_gradient_accum / gm.denom).nan_to_num(0).median())
        torch.manual_seed(123)
        gm.densify_and_prune(thr, 0.005, 2.0, 20)
        snapshot(gm, "densified", out)
        gm.reset_opacity()
        snapshot(gm, "reset", out)
        n = gm.get_xyz.shape[0]
        rng = np.random.default_rng(77)
        last = dict(g_xyz=rng.normal(0, 1e-3, (n, 3)), g_color=rng.normal(0, 1e-2, (n, 3)), g_opacity=rng.normal(0, 1e-2, (n, 1)),
                    g_scaling=rng.normal(0, 1e-2, (n, 3)), g_rotation=rng.normal(0, 1e-2, (n, 4)), g_screen=rng.normal(0, 3e-4, (n, 3)),
                    radii=rng.integers(0, 40, n))
        apply_step(gm, last, 4)
        snapshot(gm, "final", out)
    for i, h in enumerate(hist[:3]):
        out.update({f"hist{i}_{k}": v for k, v in h.items()})
    out.update({f"last_{k}": v for k, v in last.items()})
    out["densify_threshold"] = thr
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "P:", P, "->", n)


if __name__ == "__main__":
    main()
