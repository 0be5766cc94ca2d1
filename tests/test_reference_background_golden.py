"""CPU: oracle/background_ref.py against the reference's OWN gm_background.GaussianModel.  tools/make_background_golden.py
drove that class (imported from /root/reference, CPU, device="cuda" redirected) through Adam steps with recorded gradients,
the training script's statistics updates, densify_and_prune, reset_opacity and one more step; the same history replayed
through RefBackgroundModel must reproduce every tensor (parameters, Adam moments, statistics, the set and order of the
Gaussians after cloning / splitting / pruning).  The GPU tests then hold the fused kernels to RefBackgroundModel."""
import os

import numpy as np
import torch

from oracle import background_ref as OB

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "pyref_background.npz"))
NAMES = ("xyz", "color", "opacity", "scaling", "rotation")


class Args:
    position_lr_init, position_lr_final, position_lr_delay_mult, position_lr_max_steps = 1.6e-4, 1.6e-6, 0.01, 30_000
    color_lr, opacity_lr, scaling_lr, rotation_lr, percent_dense = 2.5e-3, 0.05, 5e-3, 1e-3, 0.01


def _expon(step):  # the xyz schedule of gm_background.py:169-182 (checked against the reference in test_reference_python_golden.py)
    from fluidnexus_b200.background import expon_lr
    return expon_lr(1.6e-4 * 5.0, 1.6e-6 * 5.0, lr_delay_mult=0.01, max_steps=30_000)(step)


def _apply(gm, prefix, it):
    for group in gm.optimizer.param_groups:
        if group["name"] == "xyz":
            group["lr"] = _expon(it)
    for k in NAMES:
        getattr(gm, "_" + k).grad = torch.tensor(Z[f"{prefix}_g_{k}"], dtype=torch.float32)
    screen = torch.zeros((gm.get_xyz.shape[0], 3), requires_grad=True)
    screen.grad = torch.tensor(Z[f"{prefix}_g_screen"], dtype=torch.float32)
    radii = torch.tensor(Z[f"{prefix}_radii"], dtype=torch.int32)
    vis = radii > 0
    with torch.no_grad():
        gm.max_radii2D[vis] = torch.max(gm.max_radii2D[vis], radii[vis])
        gm.add_densification_stats(screen, vis)
        gm.optimizer.step()
        gm.optimizer.zero_grad(set_to_none=True)


def _check(gm, prefix, tol=1e-6):
    for k in NAMES:
        p = getattr(gm, "_" + k)
        ref = Z[f"{prefix}_{k}"]
        assert tuple(p.shape) == ref.shape, (prefix, k, tuple(p.shape), ref.shape)
        assert np.allclose(p.detach().numpy(), ref, rtol=tol, atol=tol), (prefix, k, np.abs(p.detach().numpy() - ref).max())
        st = gm.optimizer.state.get(p, None)
        if f"{prefix}_m_{k}" in Z.files:
            assert np.allclose(st["exp_avg"].numpy(), Z[f"{prefix}_m_{k}"], rtol=tol, atol=1e-9), (prefix, "m", k)
            assert np.allclose(st["exp_avg_sq"].numpy(), Z[f"{prefix}_v_{k}"], rtol=tol, atol=1e-12), (prefix, "v", k)
    assert np.allclose(gm.xyz_gradient_accum.numpy(), Z[f"{prefix}_accum"], rtol=tol, atol=1e-9)
    assert np.array_equal(gm.denom.numpy(), Z[f"{prefix}_denom"]) and np.array_equal(gm.max_radii2D.numpy(), Z[f"{prefix}_max_radii"])


def test_restated_background_model_replays_the_reference_history():
    t = lambda k: torch.tensor(Z[k], dtype=torch.float32)
    gm = OB.RefBackgroundModel(t("init_xyz"), t("init_color"), t("init_opacity"), t("init_scales"), t("init_rotations"), spatial_lr_scale=5.0)
    gm.training_setup(Args)
    for it in range(1, 4):
        _apply(gm, f"hist{it - 1}", it)
    _check(gm, "after3")
    torch.manual_seed(123)                                   # densify_and_split draws its samples from the global generator
    gm.densify_and_prune(float(Z["densify_threshold"]), 0.005, 2.0, 20)
    assert gm.get_xyz.shape[0] == Z["densified_xyz"].shape[0] and gm.get_xyz.shape[0] != Z["init_xyz"].shape[0]
    _check(gm, "densified")
    gm.reset_opacity()
    _check(gm, "reset")
    _apply(gm, "last", 4)
    _check(gm, "final")
