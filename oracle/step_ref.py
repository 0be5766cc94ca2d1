"""oracle/step_ref.py -- CPU restatement of one physical-particle training iteration.  TEST INFRASTRUCTURE ONLY.

Follows FD/entries_scalar_real/train_physical_particle.py:301-381 literally: a Python loop over the sampled views,
every view adding image + distance + exyz + gas + next-gas losses, `loss.backward()` per view, gradient cache summed
and multiplied by 1/batch (gm_fluid.py:419-430), then torch.optim.Adam(lr group, eps=1e-15).step().
The rasterizer inside is oracle/raster_ref.c (fp64 twin) wrapped as a torch.autograd.Function; the physics terms
are oracle/pbf_ref.py (torch CPU autograd, fp64).
"""
import numpy as np
import torch

from . import pbf_ref as O
from .raster_oracle import RasterOracle


class _OracleRaster(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, colors, opacity, scales, rotations, cam, bg, kind):
        o = RasterOracle(kind)
        out = o.forward(bg=bg, means3D=means3D.detach().numpy(), colors=colors.detach().numpy(),
                        opacities=opacity.detach().numpy(), scales=scales.detach().numpy(),
                        rotations=rotations.detach().numpy(), scale_modifier=1.0, view=cam["view"], proj=cam["proj"],
                        tan_fov_x=cam["tan_fov_x"], tan_fov_y=cam["tan_fov_y"], H=cam["H"], W=cam["W"])
        ctx.o = o
        return torch.from_numpy(out["color"].astype(np.float64))

    @staticmethod
    def backward(ctx, g):
        gr = ctx.o.backward(g.numpy().astype(np.float32))
        return torch.from_numpy(gr["means3D"]), None, None, None, None, None, None, None


def reference_step(prm, st, gauss, cams, gts, e0, adam_state=None, grey=False, kind="f64", do_step=True):
    """st: dict of float64 torch tensors (xyz, estimate_xyz, buoyancy, force, imass, visual_xyz);
    gauss: dict (scales, rotations, opacity [P], colors, bg_xyz [Pb,3] or None) float64 tensors;
    cams: list of dicts (view, proj, tan_fov_x, tan_fov_y, H, W); gts: list of [C,H,W] float64 tensors.
    Returns (losses dict, averaged grad [N,3], updated parameter, optimizer)."""
    e = e0.clone().double().requires_grad_(True)
    opt = adam_state or torch.optim.Adam([{"params": [e], "lr": prm.lr, "name": "estimate_xyz_nn"}], lr=0.0, eps=1e-15)
    if adam_state is not None:
        e = opt.param_groups[0]["params"][0]
    batch = len(cams)
    grad_cache = torch.zeros_like(e)
    logs = []
    for cam, gt in zip(cams, gts):
        vis = O.visual_xyz_from_nn(prm, e, st["xyz"], st["visual_xyz"])
        render_xyz = vis / O.SCALE_FACTOR
        means3D = render_xyz if gauss.get("bg_xyz") is None else torch.cat([render_xyz, gauss["bg_xyz"]], 0)
        image = _OracleRaster.apply(means3D, gauss["colors"], gauss["opacity"], gauss["scales"], gauss["rotations"], cam,
                                    gauss["bg"], kind)
        img_loss, l1, ss = O.image_loss(prm, image, gt, grey=grey)
        dist = O.distance_loss(render_xyz, prm.distance_threshold_visual) if prm.lambda_current_distance > 0 else torch.zeros(())
        exyz = O.l2_loss(e * O.SCALE_FACTOR, st["estimate_xyz"])
        p = O.gas_constraints_from_exyz_nn(prm, e, st["imass"])
        gas = O.l2_loss(p, torch.ones_like(p))
        pn = O.gas_constraints_from_vel_nn_guess(prm, e, st["xyz"], st["buoyancy"], st["force"], st["imass"])
        nxt = O.l2_loss(pn, torch.ones_like(pn))
        loss = (img_loss + prm.lambda_current_distance * dist + prm.lambda_exyz * exyz
                + prm.lambda_gas_constraints * gas + prm.lambda_next_gas_constraints * nxt)
        loss.backward()
        grad_cache += e.grad          # cache_gradient_current
        opt.zero_grad()
        logs.append(dict(l1=float(l1), ssim=float(1 - ss), dist=float(dist), exyz=float(exyz), gas=float(gas),
                         next_gas=float(nxt), total=float(loss)))
    e.grad = grad_cache * (1.0 / batch)  # set_batch_gradient_current
    g = e.grad.clone()
    if do_step:
        opt.step()
        opt.zero_grad()
    return logs, g, e.detach().clone(), opt
