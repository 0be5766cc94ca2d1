"""GPU: the fused static-background iteration (fluidnexus_b200.background) against the restatement of the reference's
loop in oracle/background_ref.py (nn.Parameters in torch.optim.Adam, torch activations and autograd, literal statistics),
both on top of libfnx's rasterizer.  Parameters, Adam moments and densification statistics after several iterations:
rel-L2 < 1e-4 (fp32 on both sides; float atomics in the rasterizer backward make either side run-to-run noisy at 1e-6).
Densification / pruning: identical selections, identical tensors (same torch.Generator for the split samples)."""
import numpy as np
import pytest
import torch

from fluidnexus_b200 import losses as FL
from fluidnexus_b200 import rasterizer as R
from fluidnexus_b200 import synthetic as S
from fluidnexus_b200.background import BackgroundModel, BackgroundStep
from oracle import background_ref as OB

pytestmark = pytest.mark.gpu


class Args:
    position_lr_init, position_lr_final, position_lr_delay_mult, position_lr_max_steps = 1.6e-4, 1.6e-6, 0.01, 30_000
    color_lr, opacity_lr, scaling_lr, rotation_lr, percent_dense = 2.5e-3, 0.05, 5e-3, 1e-3, 0.01


def rel(a, b):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def _models(P=3000, seed=0):
    g = S.background_gaussians(P, 3, seed=seed).torch("cuda")
    mine = BackgroundModel(g["xyz"], g["colors"], g["opacity"], g["scales"], g["rotations"] * 1.3, spatial_lr_scale=5.0)
    ref = OB.RefBackgroundModel(g["xyz"], g["colors"], g["opacity"], g["scales"], g["rotations"] * 1.3, spatial_lr_scale=5.0)
    mine.training_setup(Args)
    ref.training_setup(Args)
    return mine, ref


def _compare(mine, ref, tol):
    names = dict(xyz="_xyz", color="_color", opacity="_opacity", scaling="_scaling", rotation="_rotation")
    for k, attr in names.items():
        assert getattr(mine, attr).shape == getattr(ref, attr).shape, k
        assert rel(getattr(mine, attr), getattr(ref, attr)) < tol, (k, rel(getattr(mine, attr), getattr(ref, attr)))
    for group in ref.optimizer.param_groups:
        stt = ref.optimizer.state.get(group["params"][0], None)
        if stt is None:
            continue
        k = group["name"]
        assert rel(mine.exp_avg[k], stt["exp_avg"]) < 10 * tol, ("exp_avg", k, rel(mine.exp_avg[k], stt["exp_avg"]))
        assert rel(mine.exp_avg_sq[k], stt["exp_avg_sq"]) < 10 * tol, ("exp_avg_sq", k)
    assert rel(mine.xyz_gradient_accum, ref.xyz_gradient_accum) < 10 * tol
    assert torch.equal(mine.denom, ref.denom) and torch.equal(mine.max_radii2D, ref.max_radii2D)


@pytest.mark.parametrize("lam_reg", [0.0, 0.05])
def test_fused_background_iterations_match_the_reference_loop(libfnx, lam_reg):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    mine, ref = _models()
    cams = S.make_cameras(5, 96, device="cuda")
    Settings, Rasterizer, _, _ = R.make_module(3)
    bg = torch.tensor([0.1, 0.1, 0.2], device="cuda")
    step = BackgroundStep(3, lambda_dssim=0.2, lambda_reg_scaling=lam_reg, scaling_reg_ratio_threshold=2.0, bg_color=bg)
    gen = torch.Generator().manual_seed(7)
    for it in range(1, 6):
        cam = cams[it % 5]
        gt = torch.rand(3, 96, 96, generator=gen).cuda()
        for m in (mine,):
            m.update_learning_rate(it)
        for group in ref.optimizer.param_groups:
            if group["name"] == "xyz":
                group["lr"] = mine.lr["xyz"]
        out = step.step(mine, cam, gt)
        r = OB.ref_iteration(ref, cam, gt, bg, Settings, Rasterizer, FL.l1_loss, FL.ssim, lambda_dssim=0.2, lambda_reg_scaling=lam_reg,
                             ratio_threshold=2.0)
        assert abs(float(step.total_loss(out)) - float(r["loss"])) < 1e-5 * abs(float(r["loss"]))
        assert abs(float(out["reg"]) - float(r["reg"])) <= 1e-5 * abs(float(r["reg"])) + 1e-12
    _compare(mine, ref, 1e-4)


def test_densify_prune_and_opacity_reset_match(libfnx):
    mine, ref = _models(P=2500, seed=3)
    cams = S.make_cameras(5, 96, device="cuda")
    Settings, Rasterizer, _, _ = R.make_module(3)
    bg = torch.zeros(3, device="cuda")
    step = BackgroundStep(3, bg_color=bg)
    gen = torch.Generator().manual_seed(9)
    for it in range(1, 4):
        gt = torch.rand(3, 96, 96, generator=gen).cuda()
        step.step(mine, cams[it], gt)
        OB.ref_iteration(ref, cams[it], gt, bg, Settings, Rasterizer, FL.l1_loss, FL.ssim)
    # identical statistics on both sides so that the selections cannot differ by rounding
    mine.xyz_gradient_accum.copy_(ref.xyz_gradient_accum)
    for k, attr in dict(xyz="_xyz", color="_color", opacity="_opacity", scaling="_scaling", rotation="_rotation").items():
        getattr(mine, attr).copy_(getattr(ref, attr).detach())
        stt = ref.optimizer.state[[g_ for g_ in ref.optimizer.param_groups if g_["name"] == k][0]["params"][0]]
        mine.exp_avg[k].copy_(stt["exp_avg"]); mine.exp_avg_sq[k].copy_(stt["exp_avg_sq"])
    thr = float((ref.xyz_gradient_accum / ref.denom).nan_to_num(0).median())
    g1, g2 = torch.Generator(device="cuda").manual_seed(11), torch.Generator(device="cuda").manual_seed(11)
    n0 = mine._xyz.size(0)
    mine.densify_and_prune(thr, 0.005, 2.0, 20, generator=g1)    # extent 2: splats above percent_dense*extent = 0.02 split, the rest clone
    ref.densify_and_prune(thr, 0.005, 2.0, 20, generator=g2)
    assert mine._xyz.size(0) == ref._xyz.size(0) and mine._xyz.size(0) > n0
    _compare(mine, ref, 1e-7)
    mine.reset_opacity(); ref.reset_opacity()
    _compare(mine, ref, 1e-7)
    assert float(mine.get_opacity.max()) <= 0.01 + 1e-6 and float(mine.exp_avg["opacity"].abs().max()) == 0
    # training continues on the new set
    gt = torch.rand(3, 96, 96, generator=gen).cuda()
    step.step(mine, cams[0], gt)
    OB.ref_iteration(ref, cams[0], gt, bg, Settings, Rasterizer, FL.l1_loss, FL.ssim)
    _compare(mine, ref, 1e-4)
