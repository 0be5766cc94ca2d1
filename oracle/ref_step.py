"""oracle/ref_step.py -- the REFERENCE ARM of bench.py (TEST / MEASUREMENT INFRASTRUCTURE ONLY).

One training iteration exactly the way the reference does it, with none of fluidnexus_b200's code on the path:
  * rasterizer: the unmodified reference CUDA extension compiled into oracle/_ref (oracle/build_ref.py), driven like
    R3/diff_gaussian_rasterization_ch3/__init__.py:33-140 (autograd.Function around _C.rasterize_gaussians[_backward]);
  * render pipe: restated from FD/renderer/pipe_dynamics.py:44-180 (five torch.cat of fluid + frozen background per
    view, zero screen-space tensor, settings rebuilt per view);
  * image losses: plain torch on the GPU, FD/utils/loss_utils.py:9-64 (window rebuilt and uploaded per call) with the
    FluidNexus grey conversion (entries_fluid_nexus/train_physical_particle.py:356-360);
  * distance_loss: dense torch.cdist on the GPU (loss_utils.py:98-121);
  * physics terms P1-P4: oracle/pbf_ref.py on the HOST cores (torch CPU autograd; torch_cluster is not installable
    here, see that file's header) -- BASELINE.json's north_star asks for exactly this split;
  * per-view python loop, ~10 .item() reads per view, gradient cache / batch average, torch.optim.Adam(eps=1e-15)
    (train_physical_particle.py:301-381).
"""
import math

import torch

from . import pbf_ref as O
from .ref_ext import RefRaster


class _RefRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, colors, opacity, scales, rotations, rr, bg, view, proj, tfx, tfy, H, W):
        out = rr.forward(bg, means3D, colors, opacity, scales, rotations, 1.0, view, proj, tfx, tfy, H, W)
        ctx.rr, ctx.saved = rr, dict(rr.saved)
        return out["color"], out["radii"], out["depth"]

    @staticmethod
    def backward(ctx, g, _r, _d):
        ctx.rr.saved = ctx.saved
        gr = ctx.rr.backward(g.contiguous())
        return (gr["means3D"], gr["means2D"], gr["colors"], gr["opacity"], gr["scales"], gr["rotations"], None, None, None, None,
                None, None, None, None)


def distance_loss_cdist(positions, threshold):
    """loss_utils.py:98-121 verbatim semantics (dense cdist)."""
    distances = torch.cdist(positions, positions, p=2)
    mask = distances < threshold
    mask.fill_diagonal_(False)
    return ((threshold - distances) * mask.float()).clamp(min=0).pow(2).sum()


class ReferenceTrainer:
    """Holds what gm_fluid.GaussianModel holds for one frame; `iteration(cams, gts_cpu)` is one pass of the hot loop."""

    def __init__(self, prm, hidden, visual_scaled, fluid, background, channels, grey, device="cuda"):
        self.prm, self.C, self.grey, self.dev = prm, channels, grey, torch.device(device)
        f32 = lambda a: torch.as_tensor(a, dtype=torch.float32)
        self.st = dict(xyz=f32(hidden.xyz), estimate_xyz=f32(hidden.estimate_xyz), buoyancy=f32(hidden.buoyancy),
                       force=f32(hidden.force), imass=f32(hidden.imass), visual_xyz=f32(visual_scaled))
        self.e = torch.nn.Parameter((self.st["estimate_xyz"] / O.SCALE_FACTOR).clone())
        self.opt = torch.optim.Adam([{"params": [self.e], "lr": prm.lr, "name": "estimate_xyz_nn"}], lr=0.0, eps=1e-15)
        g = lambda s, k: f32(getattr(s, k)).to(self.dev)
        self.fluid = {k: g(fluid, k) for k in ("scales", "rotations", "opacity", "colors")}
        self.bg = None if background is None else {k: g(background, k) for k in ("xyz", "scales", "rotations", "opacity", "colors")}
        self.rr = RefRaster(channels)
        self.bg_color = torch.zeros(channels, device=self.dev)

    def render(self, cam, render_xyz_gpu):
        b = self.bg
        cat = (lambda a, k: a) if b is None else (lambda a, k: torch.cat([a, b[k]], dim=0))
        means3D = cat(render_xyz_gpu, "xyz")
        screen = torch.zeros_like(means3D, requires_grad=True) + 0
        opacity, scales = cat(self.fluid["opacity"], "opacity"), cat(self.fluid["scales"], "scales")
        rotations, colors = cat(self.fluid["rotations"], "rotations"), cat(self.fluid["colors"], "colors")
        tfx, tfy = math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)
        img, radii, depth = _RefRasterize.apply(means3D.float(), screen.float(), colors.float(), opacity.float(), scales.float(),
                                                rotations.float(), self.rr, self.bg_color.float(), cam.world_view_transform,
                                                cam.full_proj_transform, tfx, tfy, int(cam.image_height), int(cam.image_width))
        return img

    def gpu_part(self, cams, gts_gpu, render_xyz_gpu):
        """Only what the reference runs on the GPU for one iteration: per view render + image loss + backward to the
        Gaussian means (no host physics, no .item()): the tightest comparison with libfnx's rasterizer + loss kernels."""
        leaf = render_xyz_gpu.detach().clone().requires_grad_(True)
        for cam, gt in zip(cams, gts_gpu):
            image = self.render(cam, leaf)
            img_loss, l1, ss = O.image_loss(self.prm, image, gt, grey=self.grey)
            img_loss.backward()
        return leaf.grad

    def iteration(self, cams, gts_cpu, with_distance=True, do_step=True):
        prm, st = self.prm, self.st
        batch = len(cams)
        cache = torch.zeros_like(self.e)
        log = {}
        for cam, gt_cpu in zip(cams, gts_cpu):
            vis = O.visual_xyz_from_nn(prm, self.e, st["xyz"], st["visual_xyz"])              # host cores
            render_xyz = (vis / O.SCALE_FACTOR).to(self.dev)                                  # H2D, autograd-aware
            image = self.render(cam, render_xyz)
            gt_image = gt_cpu.float().to(self.dev)                                            # H2D every view (:325)
            img_loss, l1, ss = O.image_loss(prm, image, gt_image, grey=self.grey)
            dist = distance_loss_cdist(render_xyz, prm.distance_threshold_visual) if with_distance else torch.zeros((), device=self.dev)
            exyz = O.l2_loss(self.e * O.SCALE_FACTOR, st["estimate_xyz"])                     # host
            p = O.gas_constraints_from_exyz_nn(prm, self.e, st["imass"])
            gas = O.l2_loss(p, torch.ones_like(p))
            pn = O.gas_constraints_from_vel_nn_guess(prm, self.e, st["xyz"], st["buoyancy"], st["force"], st["imass"])
            nxt = O.l2_loss(pn, torch.ones_like(pn))
            loss = ((img_loss + prm.lambda_current_distance * dist).cpu() + prm.lambda_exyz * exyz
                    + prm.lambda_gas_constraints * gas + prm.lambda_next_gas_constraints * nxt)
            # the reference logs ten scalars per view through .item() (train_physical_particle.py:359-373)
            log = dict(l1=l1.item(), ssim=ss.item(), dist=dist.item(), exyz=exyz.item(), gas=gas.item(), next_gas=nxt.item(),
                       total=loss.item(), p_ratio=p.mean().item(), next_p_ratio=pn.mean().item())
            loss.backward()
            cache += self.e.grad
            self.opt.zero_grad()
        self.e.grad = cache * (1.0 / batch)
        if do_step:
            self.opt.step()
            self.opt.zero_grad()
        return log
