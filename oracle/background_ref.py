"""Restatement of the reference's static-background training stage (TEST INFRASTRUCTURE ONLY).

`RefBackgroundModel` follows FD/gaussian_splatting/gm_background.py line by line where it matters for parity: nn.Parameter
tensors in a torch.optim.Adam(lr=0.0, eps=1e-15) with five named groups (:155-168), optimizer-state surgery on prune / cat
(:271-347), densify_and_clone / densify_and_split / densify_and_prune (:376-433), reset_opacity (:227-230),
add_densification_stats (:472-476).  `ref_iteration` is the body of FD/entries_fluid_nexus/train_background.py:160-273 for
one camera (loss, backward, statistics, optimizer step).  The rasterizer and the image losses are passed in by the caller
(tests use the libfnx drop-in module through autograd, which is parity-tested against the compiled reference separately,
and oracle/pbf_ref.py's torch losses), so what this file pins is everything AROUND them.  Only tests/ may import it.
"""
import math

import torch
from torch import nn


def inv_sigmoid(x):
    return torch.log(x / (1 - x))


def build_rotation(r):
    """FD/utils/general_utils.py:113-134."""
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - r * z)
    R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y)
    R[:, 2, 1] = 2 * (y * z + r * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


class RefBackgroundModel:
    def __init__(self, xyz, color, opacity, scales, rotations, percent_dense=0.01, spatial_lr_scale=1.0):
        dev = xyz.device
        self.dev = dev
        self._xyz = nn.Parameter(xyz.clone().float().requires_grad_(True))
        self._color = nn.Parameter(color.clone().float().requires_grad_(True))
        self._opacity = nn.Parameter(inv_sigmoid(opacity.clone().float().reshape(-1, 1)).requires_grad_(True))
        self._scaling = nn.Parameter(torch.log(scales.clone().float()).requires_grad_(True))
        self._rotation = nn.Parameter(rotations.clone().float().requires_grad_(True))
        self.percent_dense, self.spatial_lr_scale, self.active_sh_degree = percent_dense, spatial_lr_scale, 0
        self.max_radii2D = torch.zeros((xyz.shape[0]), device=dev)

    get_xyz = property(lambda s: s._xyz)
    get_color = property(lambda s: s._color)
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))
    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s._rotation))

    def training_setup(self, a):
        self.percent_dense = a.percent_dense
        self.xyz_gradient_accum = torch.zeros((self.get_xyz.shape[0], 1), device=self.dev)
        self.denom = torch.zeros((self.get_xyz.shape[0], 1), device=self.dev)
        groups = [
            {"params": [self._xyz], "lr": a.position_lr_init * self.spatial_lr_scale, "name": "xyz"},
            {"params": [self._color], "lr": a.color_lr, "name": "color"},
            {"params": [self._opacity], "lr": a.opacity_lr, "name": "opacity"},
            {"params": [self._scaling], "lr": a.scaling_lr, "name": "scaling"},
            {"params": [self._rotation], "lr": a.rotation_lr, "name": "rotation"},
        ]
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)

    def _rebind(self, tensors):
        self._xyz, self._color, self._opacity = tensors["xyz"], tensors["color"], tensors["opacity"]
        self._scaling, self._rotation = tensors["scaling"], tensors["rotation"]

    def replace_tensor_to_optimizer(self, tensor, name):
        out = {}
        for group in self.optimizer.param_groups:
            if group["name"] == name:
                stored = self.optimizer.state.get(group["params"][0], None)
                stored["exp_avg"] = torch.zeros_like(tensor)
                stored["exp_avg_sq"] = torch.zeros_like(tensor)
                del self.optimizer.state[group["params"][0]]
                group["params"][0] = nn.Parameter(tensor.requires_grad_(True))
                self.optimizer.state[group["params"][0]] = stored
                out[group["name"]] = group["params"][0]
        return out

    def reset_opacity(self):
        new = inv_sigmoid(torch.min(self.get_opacity, torch.ones_like(self.get_opacity) * 0.01))
        self._opacity = self.replace_tensor_to_optimizer(new, "opacity")["opacity"]

    def _prune_optimizer(self, mask):
        out = {}
        for group in self.optimizer.param_groups:
            stored = self.optimizer.state.get(group["params"][0], None)
            if stored is not None:
                stored["exp_avg"] = stored["exp_avg"][mask]
                stored["exp_avg_sq"] = stored["exp_avg_sq"][mask]
                del self.optimizer.state[group["params"][0]]
                group["params"][0] = nn.Parameter((group["params"][0][mask].requires_grad_(True)))
                self.optimizer.state[group["params"][0]] = stored
            else:
                group["params"][0] = nn.Parameter(group["params"][0][mask].requires_grad_(True))
            out[group["name"]] = group["params"][0]
        return out

    def prune_points(self, mask):
        valid = ~mask
        self._rebind(self._prune_optimizer(valid))
        self.xyz_gradient_accum = self.xyz_gradient_accum[valid]
        self.denom = self.denom[valid]
        self.max_radii2D = self.max_radii2D[valid]

    def cat_tensors_to_optimizer(self, d):
        out = {}
        for group in self.optimizer.param_groups:
            ext = d[group["name"]]
            stored = self.optimizer.state.get(group["params"][0], None)
            if stored is not None:
                stored["exp_avg"] = torch.cat((stored["exp_avg"], torch.zeros_like(ext)), dim=0)
                stored["exp_avg_sq"] = torch.cat((stored["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
                del self.optimizer.state[group["params"][0]]
                group["params"][0] = nn.Parameter(torch.cat((group["params"][0], ext), dim=0).requires_grad_(True))
                self.optimizer.state[group["params"][0]] = stored
            else:
                group["params"][0] = nn.Parameter(torch.cat((group["params"][0], ext), dim=0).requires_grad_(True))
            out[group["name"]] = group["params"][0]
        return out

    def densification_postfix(self, new_xyz, new_color, new_opacities, new_scaling, new_rotation):
        d = {"xyz": new_xyz, "color": new_color, "opacity": new_opacities, "scaling": new_scaling, "rotation": new_rotation}
        self._rebind(self.cat_tensors_to_optimizer(d))
        n = self.get_xyz.shape[0]
        self.xyz_gradient_accum = torch.zeros((n, 1), device=self.dev)
        self.denom = torch.zeros((n, 1), device=self.dev)
        self.max_radii2D = torch.zeros((n), device=self.dev)

    def densify_and_split(self, grads, grad_threshold, scene_extent, N=2, generator=None):
        n_init = self.get_xyz.shape[0]
        padded = torch.zeros((n_init), device=self.dev)
        padded[: grads.shape[0]] = grads.squeeze()
        sel = torch.where(padded >= grad_threshold, True, False)
        sel = torch.logical_and(sel, torch.max(self.get_scaling, dim=1).values > self.percent_dense * scene_extent)
        stds = self.get_scaling[sel].repeat(N, 1)
        means = torch.zeros((stds.size(0), 3), device=self.dev)
        samples = torch.normal(mean=means, std=stds, generator=generator)
        rots = build_rotation(self._rotation[sel]).repeat(N, 1, 1)
        new_xyz = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + self.get_xyz[sel].repeat(N, 1)
        new_scaling = torch.log(self.get_scaling[sel].repeat(N, 1) / (0.8 * N))
        self.densification_postfix(new_xyz, self._color[sel].repeat(N, 1), self._opacity[sel].repeat(N, 1), new_scaling,
                                   self._rotation[sel].repeat(N, 1))
        self.prune_points(torch.cat((sel, torch.zeros(N * sel.sum(), device=self.dev, dtype=bool))))

    def densify_and_clone(self, grads, grad_threshold, scene_extent):
        sel = torch.where(torch.norm(grads, dim=-1) >= grad_threshold, True, False)
        sel = torch.logical_and(sel, torch.max(self.get_scaling, dim=1).values <= self.percent_dense * scene_extent)
        self.densification_postfix(self._xyz[sel], self._color[sel], self._opacity[sel], self._scaling[sel], self._rotation[sel])

    def densify_and_prune(self, max_grad, min_opacity, extent, max_screen_size, generator=None):
        grads = self.xyz_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        self.densify_and_clone(grads, max_grad, extent)
        self.densify_and_split(grads, max_grad, extent, generator=generator)
        prune_mask = (self.get_opacity < min_opacity).squeeze()
        if max_screen_size:
            big_vs = self.max_radii2D > max_screen_size
            big_ws = self.get_scaling.max(dim=1).values > 0.1 * extent
            prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_vs), big_ws)
        self.prune_points(prune_mask)

    def add_densification_stats(self, viewspace_point_tensor, update_filter):
        self.xyz_gradient_accum[update_filter] += torch.norm(viewspace_point_tensor.grad[update_filter, :2], dim=-1, keepdim=True)
        self.denom[update_filter] += 1


def ref_iteration(gm, cam, gt, bg, GRsetting, GRzer, l1_loss, ssim, lambda_dssim=0.2, lambda_reg_scaling=0.0, ratio_threshold=5.0,
                  update_stats=True):
    """train_background.py:160-273 for one camera, with render_background (renderer/pipe_background.py) inlined."""
    xyz = gm.get_xyz
    screen = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    screen.retain_grad()
    rs = GRsetting(image_height=int(cam.image_height), image_width=int(cam.image_width), tan_fov_x=math.tan(cam.FoVx * 0.5),
                   tan_fov_y=math.tan(cam.FoVy * 0.5), bg=bg.float(), scale_modifier=1.0, view_matrix=cam.world_view_transform,
                   proj_matrix=cam.full_proj_transform, sh_degree=0, campos=cam.camera_center, prefiltered=False)
    image, radii, _ = GRzer(raster_settings=rs)(means3D=xyz.float(), means2D=screen.float(), shs=None, colors_precomp=gm.get_color.float(),
                                                opacities=gm.get_opacity.float(), scales=gm.get_scaling.float(),
                                                rotations=gm.get_rotation.float(), cov3D_precomp=None)
    l1_value = l1_loss(image, gt)
    ssim_value = 1.0 - ssim(image, gt)
    loss = (1.0 - lambda_dssim) * l1_value + lambda_dssim * ssim_value
    reg = torch.zeros((), device=xyz.device)
    if lambda_reg_scaling > 0:
        scaling = gm.get_scaling
        smax, smin = torch.max(scaling, dim=1).values, torch.min(scaling, dim=1).values
        reg = torch.max(smax / smin - ratio_threshold, torch.zeros_like(smin)).mean()
        loss = loss + lambda_reg_scaling * reg
    loss.backward()
    with torch.no_grad():
        vis = radii > 0
        if update_stats:
            gm.max_radii2D[vis] = torch.max(gm.max_radii2D[vis], radii[vis])
            gm.add_densification_stats(screen, vis)
        gm.optimizer.step()
        gm.optimizer.zero_grad(set_to_none=True)
    return dict(loss=loss.detach(), l1=l1_value.detach(), ssim=1.0 - ssim_value.detach(), reg=reg.detach(), image=image.detach(), radii=radii)
