"""Run in the build container (needs /root/reference): imports the reference's OWN pure-Python modules on the CPU
(FluidDynamics/utils/loss_utils.py, graphics_utils.py, general_utils.py, sh_utils.py -- none of them needs a GPU or the
un-installable torch_cluster) and records their outputs on seeded inputs in tests/golden/pyref_python.npz.  Those values pin
oracle/pbf_ref.py's restatements of the image / distance losses, fluidnexus_b200/synthetic.py's camera matrices and the
learning-rate schedule / colour conversion helpers (tests/test_reference_python_golden.py).  Nothing is copied from the
reference; only numbers it computes are stored."""
import math
import os
import sys

import numpy as np
import torch

REF = "/root/reference/FluidDynamics"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pyref_python.npz")


def images(C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    base = torch.stack([0.5 + 0.4 * torch.sin(xx / (5.0 + c) + c) * torch.cos(yy / 7.0) for c in range(C)])
    img = (base + 0.05 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    gt = (base.roll(2, 2) + 0.05 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    return img.double(), gt.double()


def main():
    sys.path.insert(0, REF)
    from utils import general_utils as GU
    from utils import graphics_utils as GR
    from utils import loss_utils as LU
    from utils import sh_utils as SH
    out = {}
    # ---- image losses (loss_utils.py:9-64) and their gradients, fp64 ----
    for tag, (C, H, W, seed) in {"rgb": (3, 48, 40, 1), "grey1": (1, 37, 53, 2)}.items():
        img, gt = images(C, H, W, seed)
        x = img.clone().requires_grad_(True)
        l1 = LU.l1_loss(x, gt)
        ss = LU.ssim(x, gt)
        (0.8 * l1 + 0.2 * (1.0 - ss)).backward()
        out.update({f"loss_{tag}_img": img.numpy(), f"loss_{tag}_gt": gt.numpy(), f"loss_{tag}_l1": l1.item(), f"loss_{tag}_ssim": ss.item(),
                    f"loss_{tag}_grad": x.grad.numpy()})
    # the FluidNexus entries' grey conversion (entries_fluid_nexus/train_physical_particle.py:356-360) through the same functions
    img, gt = images(3, 40, 44, 3)
    x = img.clone().requires_grad_(True)
    gi, gg = torch.mean(x, dim=0, keepdim=True).repeat(3, 1, 1), torch.mean(gt, dim=0, keepdim=True).repeat(3, 1, 1)
    l1, ss = LU.l1_loss(gi, gg), LU.ssim(gi, gg)
    (0.8 * l1 + 0.2 * (1.0 - ss)).backward()
    out.update(loss_greyed_img=img.numpy(), loss_greyed_gt=gt.numpy(), loss_greyed_l1=l1.item(), loss_greyed_ssim=ss.item(),
               loss_greyed_grad=x.grad.numpy())
    # ---- distance_loss (loss_utils.py:98-121) ----
    rng = np.random.default_rng(7)
    p = rng.uniform(0, 0.02, (400, 3))
    p[10] = p[11]
    for thr in (0.004, 0.0005):
        x = torch.tensor(p, dtype=torch.float64, requires_grad=True)
        v = LU.distance_loss(x, thr)
        v.backward()
        out.update({f"dist_{thr}_value": v.item(), f"dist_{thr}_grad": x.grad.numpy()})
    out["dist_points"] = p
    out["l2_value"] = LU.l2_loss(torch.tensor(p), torch.tensor(p[::-1].copy())).item()
    # ---- cameras (graphics_utils.py:24-60, scene/camera.py:90-110) ----
    R = np.array([[0.36, 0.48, -0.8], [-0.8, 0.6, 0.0], [0.48, 0.64, 0.6]])
    T = np.array([0.1, -0.2, 1.3])
    w2v = GR.get_world_2_view2(R, T, np.array([0.0, 0.0, 0.0]), 1.0)
    wvt = torch.tensor(w2v).transpose(0, 1)
    proj = GR.get_projection_matrix(z_near=0.01, z_far=100.0, fovX=0.69, fovY=0.55).transpose(0, 1)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    out.update(cam_R=R, cam_T=T, cam_world_view_transform=wvt.numpy(), cam_projection_matrix=proj.numpy(), cam_full_proj_transform=full.numpy(),
               cam_center=wvt.inverse()[3, :3].numpy(),
               cam_w2v_shifted=GR.get_world_2_view2(R, T, np.array([0.5, -0.25, 0.125]), 2.0))
    # ---- helpers ----
    f = GU.get_expon_lr_func(lr_init=1.6e-4, lr_final=1.6e-6, lr_delay_mult=0.01, max_steps=30_000)
    g = GU.get_expon_lr_func(lr_init=1e-2, lr_final=1e-4, lr_delay_steps=100, lr_delay_mult=0.01, max_steps=1000)
    steps = np.array([0, 1, 50, 100, 999, 15_000, 30_000, 10 ** 6])
    out.update(lr_steps=steps, lr_plain=np.array([f(int(s)) for s in steps]), lr_delayed=np.array([g(int(s)) for s in steps]))
    xs = torch.linspace(0.01, 0.99, 33)
    out.update(invsig_x=xs.numpy(), invsig_y=GU.inv_sigmoid(xs).numpy())
    rgb = rng.uniform(0, 1, (17, 3))
    out.update(sh_rgb=rgb, sh_dc=SH.rgb2sh(rgb))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "with", len(out), "arrays")


if __name__ == "__main__":
    main()
