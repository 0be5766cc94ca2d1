"""Static-background training stage on libfnx.

`BackgroundModel` keeps the state and the method names of FD/gaussian_splatting/gm_background.py:GaussianModel --
raw tensors `_xyz, _color, _opacity, _scaling, _rotation`, the activated accessors `get_*`, the densification statistics
`xyz_gradient_accum, denom, max_radii2D`, `training_setup / update_learning_rate / densify_and_prune / densify_and_clone /
densify_and_split / prune_points / reset_opacity / add_densification_stats` -- but owns its Adam moments as plain tensors
(one per parameter tensor) instead of a torch.optim.Adam object, so pruning and cloning are a row selection / a
concatenation of every per-Gaussian tensor in one place instead of the optimizer-state surgery of
gm_background.py:271-347.

`BackgroundStep.step()` is one iteration of FD/entries_fluid_nexus/train_background.py:160-273 as a launch sequence:
fnx_gs_activate -> rasterizer forward (all five attributes trainable) -> fused L1/SSIM loss -> rasterizer backward ->
fnx_gs_update (activation backward + scaling regulariser + densification statistics + Adam on the five tensors).
There is no CPU fallback and no host synchronisation inside an iteration.
"""
import ctypes as C
import math

import torch

from . import _lib as L
from . import rasterizer as R

_PARAMS = ("xyz", "color", "opacity", "scaling", "rotation")


def inverse_sigmoid(x):
    return torch.log(x / (1 - x))


def expon_lr(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """Log-linear learning-rate decay with an optional eased start (FD/utils/general_utils.py:63-94)."""
    def rate(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        delay = 1.0
        if lr_delay_steps > 0:
            delay = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
        t = min(max(step / max_steps, 0.0), 1.0)
        return delay * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)
    return rate


def quaternion_to_matrix(q):
    """Rotation matrices [n,3,3] of (r,x,y,z) quaternions, normalised first (general_utils.py:113-134)."""
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q.unbind(1)
    rows = [1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
            2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
            2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]
    return torch.stack(rows, dim=1).reshape(-1, 3, 3)


class BackgroundModel:
    def __init__(self, xyz, color, opacity, scales, rotations, device="cuda", percent_dense=0.01, spatial_lr_scale=1.0):
        """xyz [P,3], color [P,C], opacity [P,1] in (0,1), scales [P,3] > 0, rotations [P,4]: ACTIVATED values, stored raw
        (log scale, logit opacity) like create_from_pcd does (gm_background.py:116-144)."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("BackgroundModel lives on a CUDA device (no CPU fallback)")
        f = lambda t: torch.as_tensor(t, dtype=torch.float32).to(dev).contiguous().clone()
        self.dev = dev
        self._xyz, self._color, self._rotation = f(xyz), f(color), f(rotations)
        self._opacity = inverse_sigmoid(f(opacity).reshape(-1, 1))
        self._scaling = torch.log(f(scales))
        self.percent_dense, self.spatial_lr_scale, self.active_sh_degree = percent_dense, spatial_lr_scale, 0
        self.exp_avg = {k: torch.zeros_like(self._raw(k)) for k in _PARAMS}
        self.exp_avg_sq = {k: torch.zeros_like(self._raw(k)) for k in _PARAMS}
        self.step_count = 0
        self._reset_stats()
        self.lr = dict(xyz=0.0, color=0.0, opacity=0.0, scaling=0.0, rotation=0.0)
        self.xyz_scheduler = None

    # -- state access ---------------------------------------------------------------------------------------------
    def _raw(self, name):
        return getattr(self, "_" + name)

    def _reset_stats(self):
        P = self._xyz.size(0)
        self.xyz_gradient_accum = torch.zeros((P, 1), device=self.dev)
        self.denom = torch.zeros((P, 1), device=self.dev)
        self.max_radii2D = torch.zeros((P,), device=self.dev)

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_color(self):
        return self._color

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)

    # -- optimiser configuration (gm_background.py:155-182) ---------------------------------------------------------
    def training_setup(self, training_args):
        """training_args: object with position_lr_init/final/delay_mult/max_steps, color_lr, opacity_lr, scaling_lr,
        rotation_lr, percent_dense (FD/arguments/__init__.py)."""
        self.percent_dense = training_args.percent_dense
        self._reset_stats()
        self.lr = dict(xyz=training_args.position_lr_init * self.spatial_lr_scale, color=training_args.color_lr,
                       opacity=training_args.opacity_lr, scaling=training_args.scaling_lr, rotation=training_args.rotation_lr)
        self.xyz_scheduler = expon_lr(training_args.position_lr_init * self.spatial_lr_scale,
                                      training_args.position_lr_final * self.spatial_lr_scale,
                                      lr_delay_mult=training_args.position_lr_delay_mult, max_steps=training_args.position_lr_max_steps)

    def update_learning_rate(self, iteration):
        self.lr["xyz"] = self.xyz_scheduler(iteration)
        return self.lr["xyz"]

    # -- per-Gaussian row surgery: every per-Gaussian tensor in one place ---------------------------------------------
    def _select(self, keep):
        for k in _PARAMS:
            setattr(self, "_" + k, self._raw(k)[keep].contiguous())
            self.exp_avg[k] = self.exp_avg[k][keep].contiguous()
            self.exp_avg_sq[k] = self.exp_avg_sq[k][keep].contiguous()
        self.xyz_gradient_accum, self.denom = self.xyz_gradient_accum[keep], self.denom[keep]
        self.max_radii2D = self.max_radii2D[keep]

    def prune_points(self, mask):
        self._select(~mask)

    def densification_postfix(self, new_xyz, new_color, new_opacities, new_scaling, new_rotation):
        """Appends new Gaussians with zero Adam moments and clears ALL statistics (gm_background.py:349-374)."""
        new = dict(xyz=new_xyz, color=new_color, opacity=new_opacities, scaling=new_scaling, rotation=new_rotation)
        for k in _PARAMS:
            setattr(self, "_" + k, torch.cat((self._raw(k), new[k]), dim=0).contiguous())
            pad = torch.zeros_like(new[k])
            self.exp_avg[k] = torch.cat((self.exp_avg[k], pad), dim=0)
            self.exp_avg_sq[k] = torch.cat((self.exp_avg_sq[k], pad), dim=0)
        self._reset_stats()

    def reset_opacity(self):
        self._opacity = inverse_sigmoid(torch.min(self.get_opacity, torch.ones_like(self._opacity) * 0.01))
        self.exp_avg["opacity"] = torch.zeros_like(self._opacity)
        self.exp_avg_sq["opacity"] = torch.zeros_like(self._opacity)

    # -- densification (gm_background.py:376-433) ---------------------------------------------------------------------
    def densify_and_clone(self, grads, grad_threshold, scene_extent):
        sel = (torch.norm(grads, dim=-1) >= grad_threshold) & (self.get_scaling.max(dim=1).values <= self.percent_dense * scene_extent)
        self.densification_postfix(self._xyz[sel], self._color[sel], self._opacity[sel], self._scaling[sel], self._rotation[sel])

    def densify_and_split(self, grads, grad_threshold, scene_extent, N=2, generator=None):
        P = self._xyz.size(0)
        padded = torch.zeros((P,), device=self.dev)
        padded[:grads.shape[0]] = grads.squeeze()
        sel = (padded >= grad_threshold) & (self.get_scaling.max(dim=1).values > self.percent_dense * scene_extent)
        stds = self.get_scaling[sel].repeat(N, 1)
        samples = torch.normal(mean=torch.zeros_like(stds), std=stds, generator=generator)
        rots = quaternion_to_matrix(self._rotation[sel]).repeat(N, 1, 1)
        new_xyz = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + self._xyz[sel].repeat(N, 1)
        new_scaling = torch.log(self.get_scaling[sel].repeat(N, 1) / (0.8 * N))
        self.densification_postfix(new_xyz, self._color[sel].repeat(N, 1), self._opacity[sel].repeat(N, 1), new_scaling,
                                   self._rotation[sel].repeat(N, 1))
        self.prune_points(torch.cat((sel, torch.zeros(N * int(sel.sum()), device=self.dev, dtype=torch.bool))))

    def densify_and_prune(self, max_grad, min_opacity, extent, max_screen_size, generator=None, **kwargs):
        grads = self.xyz_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        self.densify_and_clone(grads, max_grad, extent)
        self.densify_and_split(grads, max_grad, extent, generator=generator)
        prune = (self.get_opacity < min_opacity).squeeze()
        if max_screen_size:
            prune = prune | (self.max_radii2D > max_screen_size) | (self.get_scaling.max(dim=1).values > 0.1 * extent)
        self.prune_points(prune)

    # -- point-cloud file of the stage (gm_background.py:203-269), see fluidnexus_b200/io.py ------------------------------
    def save_ply(self, path):
        from . import io as IO
        IO.save_background_ply(path, self._xyz, self._color, self._opacity, self._scaling, self._rotation)

    def load_ply(self, path):
        """Replaces the parameters with the file's (raw values), zeroes the Adam moments and the statistics."""
        from . import io as IO
        d = IO.load_background_ply(path)
        t = lambda a: torch.tensor(a, dtype=torch.float32, device=self.dev).contiguous()
        self._xyz, self._color, self._opacity = t(d["xyz"]), t(d["color"]), t(d["opacity"])
        self._scaling, self._rotation = t(d["scaling"]), t(d["rotation"])
        self.exp_avg = {k: torch.zeros_like(self._raw(k)) for k in _PARAMS}
        self.exp_avg_sq = {k: torch.zeros_like(self._raw(k)) for k in _PARAMS}
        self._reset_stats()

    def add_densification_stats(self, viewspace_point_tensor, update_filter):
        """For callers that run the drop-in rasterizer through autograd; BackgroundStep does this inside fnx_gs_update."""
        self.xyz_gradient_accum[update_filter] += torch.norm(viewspace_point_tensor.grad[update_filter, :2], dim=-1, keepdim=True)
        self.denom[update_filter] += 1


class BackgroundStep:
    """step(model, cam, gt) -> dict(l1, ssim, reg, image, radii): one optimiser iteration on one camera."""

    def __init__(self, channels=3, lambda_dssim=0.2, lambda_reg_scaling=0.0, scaling_reg_ratio_threshold=5.0, bg_color=None,
                 device="cuda"):
        self.C, self.dev = channels, torch.device(device)
        self.lambda_dssim, self.lambda_reg, self.reg_thr = lambda_dssim, lambda_reg_scaling, scaling_reg_ratio_threshold
        self.bg = torch.zeros(channels, device=self.dev) if bg_color is None else torch.as_tensor(bg_color, dtype=torch.float32).to(self.dev)
        self._loss_scratch = {}
        self.reg = torch.zeros(1, device=self.dev)

    def total_loss(self, out):
        """The reference's `loss` (train_background.py:190-201), a device scalar."""
        return ((1.0 - self.lambda_dssim) * out["l1"] + self.lambda_dssim * (1.0 - out["ssim"]) + self.lambda_reg * out["reg"]).sum()

    def step(self, model: BackgroundModel, cam, gt, update_stats=True, update=True):
        lib, dev, P, Cc = L.lib(), self.dev, model._xyz.size(0), self.C
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            scales, opacity, rotation = torch.empty((P, 3), device=dev), torch.empty((P,), device=dev), torch.empty((P, 4), device=dev)
            L.check(lib.fnx_gs_activate(P, model._scaling.data_ptr(), model._opacity.data_ptr(), model._rotation.data_ptr(),
                                        scales.data_ptr(), opacity.data_ptr(), rotation.data_ptr(), st))
            H, W = int(cam.image_height), int(cam.image_width)
            ctx, image, radii, _ = R.raster_forward(Cc, self.bg, model._xyz, model._color, opacity, scales, rotation, 1.0, None,
                                                    cam.world_view_transform, cam.full_proj_transform, math.tan(cam.FoVx * 0.5),
                                                    math.tan(cam.FoVy * 0.5), H, W)
            key = (Cc, H, W)
            if key not in self._loss_scratch:
                self._loss_scratch[key] = (torch.empty(lib.fnx_image_loss_bytes(1, Cc, H, W), dtype=torch.uint8, device=dev),
                                           torch.empty(1, device=dev), torch.empty(1, device=dev), torch.empty((1, Cc, H, W), device=dev))
            scratch, l1, ss, dimg = self._loss_scratch[key]
            gt = gt.to(dev, non_blocking=True).float().contiguous()
            L.check(lib.fnx_image_loss(1, Cc, H, W, image.data_ptr(), gt.data_ptr(), 0, 1.0 - self.lambda_dssim, self.lambda_dssim,
                                       dimg.data_ptr(), l1.data_ptr(), ss.data_ptr(), scratch.data_ptr(), st))
            g = R.raster_backward(ctx, dimg[0])
            out = dict(l1=l1, ssim=ss, reg=self.reg, image=image, radii=radii, grads=g)
            if not update:
                return out
            model.step_count += 1
            state, grads, hp = L.GsState(), L.GsGrads(), L.GsHparams()
            for k in _PARAMS:
                setattr(state, k, model._raw(k).data_ptr())
                setattr(state, "m_" + k, model.exp_avg[k].data_ptr())
                setattr(state, "v_" + k, model.exp_avg_sq[k].data_ptr())
            state.max_radii2D, state.xyz_gradient_accum, state.denom = (model.max_radii2D.data_ptr(), model.xyz_gradient_accum.data_ptr(),
                                                                        model.denom.data_ptr())
            grads.dL_dmeans3D, grads.dL_dmeans2D, grads.dL_dcolors = g["means3D"].data_ptr(), g["means2D"].data_ptr(), g["colors"].data_ptr()
            grads.dL_dopacity, grads.dL_dscales, grads.dL_drotations = g["opacity"].data_ptr(), g["scales"].data_ptr(), g["rotations"].data_ptr()
            hp.lr_xyz, hp.lr_color, hp.lr_opacity = model.lr["xyz"], model.lr["color"], model.lr["opacity"]
            hp.lr_scaling, hp.lr_rotation = model.lr["scaling"], model.lr["rotation"]
            hp.beta1, hp.beta2, hp.eps, hp.step = 0.9, 0.999, 1e-15, model.step_count
            hp.update_stats, hp.lambda_reg_scaling, hp.reg_ratio_threshold = int(bool(update_stats)), self.lambda_reg, self.reg_thr
            L.check(lib.fnx_gs_update(P, Cc, C.byref(state), C.byref(grads), C.byref(hp), radii.data_ptr(), self.reg.data_ptr(), st))
            out["_keep"] = (scales, opacity, rotation, ctx)
        return out
