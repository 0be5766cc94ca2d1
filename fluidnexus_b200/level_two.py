"""The level-two ("visual particle") training iteration as one fused launch sequence on libfnx.

Mirrors one iteration of FD/entries_fluid_nexus/train_visual_particle.py:133-222 (ScalarReal twin :129-218):

    zero_gradient_cache_current_level_two()                                   gm_dynamics.py:474-482
    for each sampled camera:
        render_func(..., pos_type="visual")                                   pipe_dynamics.py:8-180 / pipe_fluid.py:8-135
        l1_loss, 1 - ssim (no grey conversion in this stage)                  loss_utils.py:9-64
        + lambda_consistency_X * l2_loss_consistency(X, prev_X)  X in {color, opacity, scales, rotation}   loss_utils.py:138-146
        + lambda_reg_scaling * mean(max(s_max / s_min - threshold, 0))        train_visual_particle.py:174-184
        loss.backward(); cache_gradient_current_level_two()                   gm_dynamics.py:484-492
    set_batch_gradient_current_level_two(batch); optimizer.step()             gm_dynamics.py:494-503, Adam eps 1e-15 (:380-397)

The positions are fixed (loaded from the physical stage); the trainable tensors are the RAW colour [V,1|3], opacity [V,1] (sigmoid),
scales [V,3] (exp) and rotation [V,4] (normalize) of the V visual particles.  What is fused: the activations are one launch
(fnx_gs_activate), all views are rendered by one batched rasterizer call (the frozen background set, if any, lives in a StaticStream
binned once), the image loss is one fused kernel, and the chain through the activations + the consistency / regulariser gradients +
Adam over the four tensors is one launch (fnx_gs_update_level_two).  The view-independent terms are evaluated once (the reference
adds them per view and averages: weight 1).
"""
import ctypes as C
import math
from dataclasses import dataclass

import torch

from . import _lib as L
from . import rasterizer as R


@dataclass
class LevelTwoParams:
    """FD/arguments/__init__.py + configs/fluid_nexus_smoke_dynamics.json (level-two block)."""
    lambda_dssim: float = 0.2
    lambda_image: float = 1.0
    lambda_consistency_color: float = 10.0
    lambda_consistency_opacity: float = 8.0
    lambda_consistency_scales: float = 0.0
    lambda_consistency_rotation: float = 0.1
    lambda_reg_scaling: float = 1.0
    scaling_reg_ratio_threshold: float = 4.0
    visual_color_lr: float = 0.0025
    visual_opacity_lr: float = 0.05
    visual_scales_lr: float = 0.005
    visual_rotation_lr: float = 0.001
    fit_color: bool = True
    fit_opacity: bool = True
    fit_scales: bool = True
    fit_rotation: bool = True
    adam_eps: float = 1e-15


class LevelTwoState:
    """Per-frame state of the level-two stage: fixed positions (render units), raw trainable attributes with their Adam moments,
    the previous frame's raw attributes (consistency targets; None for the first frame), the frozen background set (activated)."""

    def __init__(self, xyz, color, opacity, scales, rotation, prev=None, background=None, device="cuda"):
        dev = torch.device(device)
        f = lambda a: torch.as_tensor(a, dtype=torch.float32).to(dev).contiguous()
        self.dev = dev
        self.xyz = f(xyz)
        self.V = self.xyz.size(0)
        self.color, self.opacity, self.scales, self.rotation = f(color), f(opacity).reshape(-1, 1), f(scales), f(rotation)
        self.Cp = self.color.size(1)
        z = torch.zeros_like
        self.m = dict(color=z(self.color), opacity=z(self.opacity), scales=z(self.scales), rotation=z(self.rotation))
        self.v = dict(color=z(self.color), opacity=z(self.opacity), scales=z(self.scales), rotation=z(self.rotation))
        self.prev = None
        if prev is not None:
            self.prev = {k: f(prev[k]) for k in ("color", "opacity", "scales", "rotation")}
            self.prev["opacity"] = self.prev["opacity"].reshape(-1, 1)
            self.prev_num = self.prev["color"].size(0)
            assert self.prev_num <= self.V, "Current number of particles must be greater than or equal to the previous"
        self.bg_set = None if background is None else {k: v.contiguous() for k, v in background.torch(dev).items()}
        self.Pb = 0 if background is None else background.P
        self.step_count = 0
        self.static_stream = None
        self.ws = {}
        # activated attributes handed to the rasterizer (rewritten every iteration)
        self.act_scales = torch.empty((self.V, 3), device=dev)
        self.act_opacity = torch.empty((self.V,), device=dev)
        self.act_rotation = torch.empty((self.V, 4), device=dev)
        self.act_colors = None        # [V, C] with the render channel count (allocated by the step)
        self.losses = torch.zeros(5, device=dev)   # consistency colour / opacity / scales / rotation, scaling regulariser


def init_quantities_current_level_two(visual_xyz, color, opacity, scales, rotation, prev=None, fit_color=True, fit_opacity=True,
                                      fit_scales=True, fit_rotation=True, init_scales_w_xyz_dist=False, inherit_prev_color=False,
                                      inherit_prev_opacity=False, inherit_prev_scales=False, inherit_prev_rotation=False, dist2_fn=None):
    """Raw attributes a frame's level-two fit starts from (gm_dynamics.py:363-378, called at train_visual_particle.py:111 right after
    load_visual): optionally log-scales from the mean distance to the 3 nearest particles (distCUDA2, clamped to [-10, 1]); then, for
    every fitted attribute whose `inherit_prev_*` flag is set, the first `prev[X].shape[0]` particles -- the ones that already existed
    in the previous frame, new particles are appended -- take the previous frame's fitted values.  `prev`: dict with any of
    color / opacity / scales / rotation (or None for the first frame).  Flag defaults: FD/arguments/__init__.py:395-401 (the
    FluidNexus configs inherit all four).  Returns new tensors (color, opacity, scales, rotation)."""
    color, opacity, scales, rotation = color.clone(), opacity.clone(), scales.clone(), rotation.clone()
    if fit_scales and init_scales_w_xyz_dist:
        if dist2_fn is None:
            from .physics import distCUDA2 as dist2_fn
        d2 = torch.clamp_min(dist2_fn(visual_xyz.float()), 0.0000001)
        scales = torch.clamp(torch.log(torch.sqrt(d2))[..., None].repeat(1, 3), -10, 1.0)
    out = dict(color=color, opacity=opacity, scales=scales, rotation=rotation)
    flags = dict(color=fit_color and inherit_prev_color, opacity=fit_opacity and inherit_prev_opacity, scales=fit_scales and inherit_prev_scales,
                 rotation=fit_rotation and inherit_prev_rotation)
    for name, on in flags.items():
        p = None if prev is None else prev.get(name)
        if on and p is not None:
            out[name][: p.shape[0]] = p.to(out[name].device, out[name].dtype).reshape(p.shape[0], -1)
    return out["color"], out["opacity"], out["scales"], out["rotation"]


class LevelTwoStep:
    """step(state, view_ids, gt) -> dict of device scalars; `cams` as for PhysicalStep."""

    def __init__(self, cams, channels, prm: LevelTwoParams = None, bg_color=None, device="cuda"):
        self.prm = prm or LevelTwoParams()
        self.dev = torch.device(device)
        self.C = channels
        self.view_all = torch.stack([c.world_view_transform.float() for c in cams]).to(self.dev).contiguous()
        self.proj_all = torch.stack([c.full_proj_transform.float() for c in cams]).to(self.dev).contiguous()
        c0 = cams[0]
        self.W, self.H = int(c0.image_width), int(c0.image_height)
        self.tan_fov_x, self.tan_fov_y = math.tan(c0.FoVx * 0.5), math.tan(c0.FoVy * 0.5)
        self.bg = torch.zeros(channels, device=self.dev) if bg_color is None else torch.as_tensor(bg_color, dtype=torch.float32).to(self.dev)
        self._loss_scratch, self._views = {}, {}
        self.lib = L.lib()
        self.want = ("colors", "opacity", "scales", "rotations")

    def _view_mats(self, view_ids):
        key = tuple(view_ids)
        if key not in self._views:
            idx = torch.tensor(list(key), dtype=torch.long, device=self.dev)
            self._views[key] = (self.view_all[idx].contiguous(), self.proj_all[idx].contiguous())
        return self._views[key]

    def _activate(self, st: LevelTwoState):
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        L.check(self.lib.fnx_gs_activate(st.V, st.scales.data_ptr(), st.opacity.data_ptr(), st.rotation.data_ptr(), st.act_scales.data_ptr(),
                                         st.act_opacity.data_ptr(), st.act_rotation.data_ptr(), stream))
        if st.act_colors is None:
            st.act_colors = torch.empty((st.V, self.C), device=self.dev)
        # grey particles are repeated to the render channels (pipe_dynamics.py:118-120)
        st.act_colors.copy_(st.color.expand(st.V, self.C) if st.Cp == 1 else st.color)

    def _workspace(self, st: LevelTwoState, view_ids):
        key = tuple(view_ids)
        ws = st.ws.get(key)
        # The positions are fixed in this stage, so the instance count only creeps (scales and opacities are trained at small
        # learning rates): re-size as soon as the last FINISHED forward used more than 90 % of the capacity, well before a forward can
        # overflow it (an overflowed forward renders only the background, and this stage's update is not gated on the device).
        if ws is not None and int(ws.count[0]) > 0.9 * ws.capacity:
            torch.cuda.synchronize(self.dev)
            st.ws.pop(key)
            ws = None
        if ws is None:
            vm, pm = self._view_mats(view_ids)
            dyn = dict(means3D=st.xyz, colors=st.act_colors, opacities=st.act_opacity, scales=st.act_scales, rotations=st.act_rotation)
            if st.Pb > 0 and self.C == 3:
                if st.static_stream is None:
                    b = st.bg_set
                    sta = dict(means3D=b["xyz"], colors=b["colors"], opacities=b["opacity"].reshape(-1).contiguous(), scales=b["scales"],
                               rotations=b["rotations"])
                    st.static_stream = R.StaticStream(self.dev, self.view_all.size(0), self.H, self.W, self.bg, sta, self.view_all, self.proj_all,
                                                      self.tan_fov_x, self.tan_fov_y)
                ws = R.MergedRasterWorkspace(self.dev, st.V, len(view_ids), self.H, self.W, self.bg, dyn, None, vm, pm, self.tan_fov_x, self.tan_fov_y,
                                             margin=1.5, static_stream=st.static_stream, view_ids=list(view_ids), want=self.want)
                ws.dyn = dyn
            else:
                if st.Pb > 0:
                    raise NotImplementedError("a frozen background set needs the 3-channel rasterizer (FluidNexus scenes)")
                ctx, _, _, _ = R.raster_forward(self.C, self.bg, st.xyz, st.act_colors, st.act_opacity, st.act_scales, st.act_rotation, 1.0, None,
                                                vm, pm, self.tan_fov_x, self.tan_fov_y, self.H, self.W, speculative=False)
                cap = int(ctx.num_rendered * 1.5) + 65536
                del ctx
                ws = R.RasterWorkspace(self.dev, self.C, st.V, len(view_ids), self.H, self.W, cap, want=self.want)
            st.ws[key] = ws
        return ws

    def step(self, st: LevelTwoState, view_ids, gt, batch=None, update=True):
        """One optimiser iteration.  gt [len(view_ids), C, H, W] (device or host tensor).  Returns device tensors: l1 / ssim per
        view, the five view-independent loss values, the rendered images."""
        prm, lib, dev = self.prm, self.lib, self.dev
        ids = sorted(int(v) for v in view_ids)
        if ids != [int(v) for v in view_ids]:
            order = sorted(range(len(view_ids)), key=[int(v) for v in view_ids].__getitem__)
            gt = gt[order]
        batch = len(ids) if batch is None else batch
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            if not gt.is_cuda:
                gt = gt.to(dev, non_blocking=True)
            gt = gt.float().contiguous()
            self._activate(st)
            ws = self._workspace(st, ids)
            if isinstance(ws, R.MergedRasterWorkspace):
                d = ws.dyn
                img = ws.forward(d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"])
            else:
                vm, pm = self._view_mats(ids)
                img = ws.forward(self.bg, st.xyz, st.act_colors, st.act_opacity, st.act_scales, st.act_rotation, 1.0, vm, pm, self.tan_fov_x,
                                 self.tan_fov_y)
            nv = len(ids)
            key = (nv, self.C, self.H, self.W)
            if key not in self._loss_scratch:
                self._loss_scratch[key] = (torch.empty(lib.fnx_image_loss_bytes(nv, self.C, self.H, self.W), dtype=torch.uint8, device=dev),
                                           torch.empty(nv, device=dev), torch.empty(nv, device=dev), torch.empty((nv, self.C, self.H, self.W), device=dev))
            scratch, l1, ss, dimg = self._loss_scratch[key]
            w_l1 = (1.0 - prm.lambda_dssim) * prm.lambda_image / batch
            w_ss = prm.lambda_dssim * prm.lambda_image / batch
            L.check(lib.fnx_image_loss(nv, self.C, self.H, self.W, img.data_ptr(), gt.data_ptr(), 0, w_l1, w_ss, dimg.data_ptr(), l1.data_ptr(),
                                       ss.data_ptr(), scratch.data_ptr(), stream))
            g = ws.backward(dimg)
            out = dict(l1=l1, ssim=ss, images=img, losses=st.losses, ws=ws, view_ids=ids, grads=g)
            if not update:
                return out
            st.step_count += 1
            state, grads, hp = L.GsState(), L.GsGrads(), L.GsLevelTwo()
            for name, raw in (("color", st.color), ("opacity", st.opacity), ("scaling", st.scales), ("rotation", st.rotation)):
                k = "scales" if name == "scaling" else name
                setattr(state, name, raw.data_ptr())
                setattr(state, "m_" + name, st.m[k].data_ptr())
                setattr(state, "v_" + name, st.v[k].data_ptr())
            grads.dL_dcolors, grads.dL_dopacity = g["colors"].data_ptr(), g["opacity"].data_ptr()
            grads.dL_dscales, grads.dL_drotations = g["scales"].data_ptr(), g["rotations"].data_ptr()
            if st.prev is not None:
                hp.prev_color, hp.prev_opacity = st.prev["color"].data_ptr(), st.prev["opacity"].data_ptr()
                hp.prev_scales, hp.prev_rotation = st.prev["scales"].data_ptr(), st.prev["rotation"].data_ptr()
                hp.prev_num = st.prev_num
            hp.color_channels = st.Cp
            hp.fit_color, hp.fit_opacity, hp.fit_scales, hp.fit_rotation = (int(prm.fit_color), int(prm.fit_opacity), int(prm.fit_scales),
                                                                            int(prm.fit_rotation))
            hp.lambda_consistency_color, hp.lambda_consistency_opacity = prm.lambda_consistency_color, prm.lambda_consistency_opacity
            hp.lambda_consistency_scales, hp.lambda_consistency_rotation = prm.lambda_consistency_scales, prm.lambda_consistency_rotation
            hp.lambda_reg_scaling, hp.reg_ratio_threshold = prm.lambda_reg_scaling, prm.scaling_reg_ratio_threshold
            hp.lr_color, hp.lr_opacity, hp.lr_scaling, hp.lr_rotation = (prm.visual_color_lr, prm.visual_opacity_lr, prm.visual_scales_lr,
                                                                         prm.visual_rotation_lr)
            hp.beta1, hp.beta2, hp.eps, hp.step = 0.9, 0.999, prm.adam_eps, st.step_count
            L.check(lib.fnx_gs_update_level_two(st.V, self.C, C.byref(state), C.byref(grads), C.byref(hp), st.losses.data_ptr(), stream))
        return out

    def total_loss(self, out, batch=None):
        """The reference's per-view `loss`, averaged over the views (device scalar)."""
        prm = self.prm
        b = out["l1"].numel() if batch is None else batch
        img = ((1.0 - prm.lambda_dssim) * prm.lambda_image * out["l1"] + prm.lambda_dssim * prm.lambda_image * (1.0 - out["ssim"])).sum() / b
        ls = out["losses"]
        return (img + prm.lambda_consistency_color * ls[0] * float(prm.fit_color) + prm.lambda_consistency_opacity * ls[1] * float(prm.fit_opacity)
                + (prm.lambda_consistency_scales * ls[2] + prm.lambda_reg_scaling * ls[4]) * float(prm.fit_scales)
                + prm.lambda_consistency_rotation * ls[3] * float(prm.fit_rotation))
