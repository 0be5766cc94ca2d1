"""The FluidDynamics physical-particle training step as ONE fused launch sequence on libfnx.

Mirrors one iteration of the reference's hot loop (FD/entries_scalar_real/train_physical_particle.py:301-381 and its
FluidNexus twin FD/entries_fluid_nexus/train_physical_particle.py:330-420) for one frame and a batch of views:

    zero_gradient_cache_current()                                   gm_fluid.py:419-421
    for each sampled camera:                                         train_physical_particle.py:310
        render_func(..., pos_type="guess_visual_nn", scale=True)    pipe_fluid.py:8-135 / pipe_dynamics.py:8-180
            gm.get_visual_xyz_from_nn()                              gm_fluid.py:1291-1336                [P1]
        l1_loss, 1 - ssim (grey conversion for FluidNexus scenes)    loss_utils.py:9-64                   [L1-L3]
        distance_loss(visual_xyz, thr)                               loss_utils.py:98-121                 [P5]
        exyz l2, gas constraints now / next tick                     gm_fluid.py:1107-1158,846-862        [P2-P4]
        loss.backward(); cache_gradient_current()                    gm_fluid.py:423-425
    set_batch_gradient_current(batch); optimizer.step()              gm_fluid.py:427-430, Adam :349       [P6,P7]

What is fused: all views are rendered by one batched rasterizer call; the view-independent physics terms are
evaluated once (the reference re-evaluates them per view and averages, which gives weight 1); every term is a
hand-written gather kernel (no edge lists, no autograd graph); there is no host synchronisation inside the step
(losses stay on the device until the caller reads them).
"""
import os
from dataclasses import dataclass

import torch

from . import _lib as L
from . import rasterizer as R


@dataclass
class StepParams:
    """Constants of FD/arguments/__init__.py + configs (values: scalar_real.json / fluid_nexus_smoke_dynamics.json)."""
    H: float = 2.0
    KNN_K: int = 100
    p0: float = 1.5
    secs: float = 0.033
    buoyancy_max_y: float = 0.0
    scale_factor: float = 100.0
    lambda_dssim: float = 0.2
    lambda_image: float = 1.0
    lambda_current_distance: float = 0.1
    lambda_exyz: float = 0.1
    lambda_gas_constraints: float = 1.0
    lambda_next_gas_constraints: float = 0.1
    distance_threshold_visual: float = 0.002
    lr: float = 1.6e-4          # position_lr_init * spatial_lr_scale; never rescheduled (gm_fluid.py:401-407)
    adam_eps: float = 1e-15
    grey: bool = False          # FluidNexus entries compare channel-mean images (train_physical_particle.py:356-360)


def _dev_f32(x, dev):
    t = torch.as_tensor(x, dtype=torch.float32)
    return t.to(dev).contiguous()


class HostTensorCache:
    """Device copies of host tensors that are handed in again and again, keyed by the tensor OBJECT: an entry is valid only while
    its weak reference still points at the very tensor that is asked for and that tensor's version counter is unchanged (in-place
    edits bump it).  An address is not an identity -- a freed image's storage can be handed to the next frame's image of the same
    shape -- so neither data_ptr() nor id() alone is trusted.  Bounded (frames come and go); dead entries are dropped first."""

    def __init__(self, max_entries=256):
        self.max_entries = int(max_entries)
        self._entries = {}     # id(tensor) -> (weakref to the tensor, version, device copy)

    def __len__(self):
        return len(self._entries)

    def get(self, t, upload):
        import weakref
        ent = self._entries.get(id(t))
        if ent is not None and ent[0]() is t and ent[1] == t._version:
            return ent[2]
        if len(self._entries) >= self.max_entries:
            dead = [k for k, e in self._entries.items() if e[0]() is None]
            for k in dead or [next(iter(self._entries))]:
                self._entries.pop(k)
        dev_copy = upload(t)
        self._entries[id(t)] = (weakref.ref(t), t._version, dev_copy)
        return dev_copy


LOSS_ROW = 16      # floats per frame in the loss table (FrameState.loss_row)
MAX_ROW_VIEWS = 5  # views whose l1 / ssim fit into the row


class FrameState:
    """Per-frame state the reference keeps in gm_fluid.GaussianModel, laid out for the fused step.

    Gaussians rendered = [V fluid particles] ++ [Pb frozen background Gaussians] (pipe_dynamics.py:51-57,139-148);
    the fluid rows of `means3D` are rewritten by P1 every step, everything else is static within a frame."""

    def __init__(self, hidden, visual_xyz_scaled, fluid, background=None, device="cuda", prm: StepParams = None):
        dev = torch.device(device)
        self.dev = dev
        self.N, self.V = hidden.N, visual_xyz_scaled.shape[0]
        self.xyz = _dev_f32(hidden.xyz, dev)
        self.estimate_xyz = _dev_f32(hidden.estimate_xyz, dev)
        self.buoyancy = _dev_f32(hidden.buoyancy, dev)
        self.force = _dev_f32(hidden.force, dev)
        self.imass = _dev_f32(hidden.imass, dev).reshape(-1)
        self.visual = _dev_f32(visual_xyz_scaled, dev)
        sf = (prm or StepParams()).scale_factor
        # training_setup_current(): _estimate_xyz_nn = estimate_xyz / scale_factor (gm_fluid.py:333-334)
        self.e = (self.estimate_xyz / sf).contiguous()
        self.m = torch.zeros_like(self.e)
        self.v = torch.zeros_like(self.e)
        self.adam_step = 0
        f = fluid.torch(dev)
        parts = [f] + ([background.torch(dev)] if background is not None else [])
        self.Pb = background.P if background is not None else 0
        self.P = self.V + self.Pb
        self.C = f["colors"].shape[1]
        cat = lambda k: torch.cat([p[k] for p in parts], 0).contiguous()
        self.means3D = cat("xyz")          # rows [0,V) are overwritten each step
        self.scales, self.rotations = cat("scales"), cat("rotations")
        self.opacity, self.colors = cat("opacity").reshape(-1).contiguous(), cat("colors")
        lib = L.lib()
        N, V = self.N, self.V
        nb = lambda n: torch.empty(lib.fnx_grid_bytes(n), dtype=torch.uint8, device=dev)
        self.gridX, self.gridY, self.gridP, self.gridVis = nb(N), nb(N), nb(V), nb(V)
        z = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt, device=dev)
        self.X, self.Y, self.dX, self.dY, self.de = z(N, 3), z(N, 3), z(N, 3), z(N, 3), z(N, 3)
        self.kthX, self.kthY, self.kthV = z(N, dt=torch.int32), z(N, dt=torch.int32), z(V, dt=torch.int32)
        self.p, self.pn, self.gp, self.gpn = z(N), z(N), z(N), z(N)
        self.num, self.den, self.dDist = z(V, 3), z(V), z(V, 3)
        # loss row of this frame: [gas, next_gas, exyz, dist | l1 of this process' k-th view (5) | ssim of the k-th view (5) | pad]
        # (one row of parallel.FlatBucket.losses when the frame is part of a sharded job: bind_flat)
        self.loss_row = torch.zeros(LOSS_ROW, dtype=torch.float32, device=dev)
        self.scalars = self.loss_row[:4]
        self.cap_flag = torch.zeros(2, dtype=torch.int32, device=dev)   # "max_num_neighbors binds" flags of P2 / P3
        self.vis_grid_built = False
        self.zero_dmeans = None
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)   # Adam step count (device side: graph-replay safe)
        self.bc_dev = torch.zeros(2, dtype=torch.float32, device=dev)
        self.static_stream = None   # the frozen set binned once for ALL cameras (rasterizer.StaticStream), shared by the workspaces
        self.skipped_iterations = 0  # iterations voided on the device because the rasterizer's instance capacity overflowed
        self.ws = {}       # sorted view ids -> rasterizer workspace
        self.graphs = {}   # (view ids, update, batch, physics) -> two alternating (CUDAGraph, outputs, gt buffer, events) slots

    def bind_flat(self, param, exp_avg, exp_avg_sq, grad, loss_row=None):
        """Alias the trainable state into slots of flat buffers (parallel.FlatBucket.views(f)); the current parameter values
        are kept.  Must be called before the first step (captured graphs hold the pointers)."""
        assert not self.graphs, "bind_flat() after an iteration was captured"
        param.copy_(self.e)
        self.e, self.m, self.v, self.de = param, exp_avg, exp_avg_sq, grad
        if loss_row is not None:
            loss_row.zero_()
            self.loss_row, self.scalars = loss_row, loss_row[:4]


class PhysicalStep:
    """step(frame, view_ids, gt) -> dict of device scalars.  `cams`: list of camera objects with the attributes of
    FD/scene/camera.py (world_view_transform, full_proj_transform, FoVx, FoVy, image_width, image_height)."""

    def __init__(self, cams, channels, prm: StepParams = None, bg_color=None, device="cuda", overlap=True, static_cache=True,
                 static_tile_cache=True):
        import math
        self.prm = prm or StepParams()
        self.dev = torch.device(device)
        self.C = channels
        self.view_all = torch.stack([c.world_view_transform.float() for c in cams]).to(self.dev).contiguous()
        self.proj_all = torch.stack([c.full_proj_transform.float() for c in cams]).to(self.dev).contiguous()
        c0 = cams[0]
        self.W, self.H = int(c0.image_width), int(c0.image_height)
        self.tan_fov_x, self.tan_fov_y = math.tan(c0.FoVx * 0.5), math.tan(c0.FoVy * 0.5)
        self.bg = _dev_f32([0.0] * channels if bg_color is None else bg_color, self.dev)
        self._loss_scratch, self._views = {}, {}
        self.capacity_margin = 1.2   # binning capacity = margin * instances of the sizing forward + slack
        self.capacity_slack = 65536
        self.static_cache = static_cache  # bin the frozen background once per frame (MergedRasterWorkspace)
        self.static_tile_cache = static_tile_cache  # ... and keep the pixels of tiles that hold no fluid instance
        self.lpt_order = os.environ.get("FNX_LPT_ORDER", "1") != "0"   # longest-tile-first start order of the blend CTAs
        self.lib = L.lib()
        # the view-independent physics terms run on a side stream next to the rasterizer (fork/join with events, also
        # inside a captured graph); overlap=False keeps everything on one stream
        self.side = torch.cuda.Stream(device=self.dev) if overlap else None
        self._ev_fork, self._ev_means, self._ev_join = (torch.cuda.Event() for _ in range(3))
        self._side_pending = False
        # host ground truth is uploaded on its own stream so that the copy for the next frame / iteration overlaps the
        # kernels of the current one (the reference uploads it inline, train_physical_particle.py:325)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        # Host ground truth that is handed in again and again (the images of a frame's cameras do not change over the 250-1000
        # iterations of the frame; the reference uploads them every time, train_physical_particle.py:353) is uploaded ONCE: device
        # copies keyed by the host tensor object and its version counter (HostTensorCache).  Off by default: the caller
        # opts in per call (step(..., cache_gt=True)) or per object.
        self.cache_gt = False
        self._gt_cache = HostTensorCache()

    # -- pieces ---------------------------------------------------------------------------------------------
    def physics_forward(self, fr: FrameState, physics=True):
        """P1-forward on the current stream (the rasterizer needs its output) and, when physics=True, P2, P3 and P5
        on a side stream that runs concurrently with the rasterizer (joined in physics_backward_and_update).
        Fills fr.means3D[:V], fr.dX, fr.dY, fr.dDist and the loss scalars.
        physics=False: a rank that renders some views of a frame whose view-independent terms another rank owns."""
        lib, prm = self.lib, self.prm
        N, V = fr.N, fr.V
        ck = L.check
        main = torch.cuda.current_stream(self.dev)
        st = main.cuda_stream
        ck(lib.fnx_pbf_next_tick_fwd(N, fr.e.data_ptr(), fr.xyz.data_ptr(), fr.buoyancy.data_ptr(), fr.force.data_ptr(), prm.secs,
                                     prm.buoyancy_max_y, prm.scale_factor, fr.X.data_ptr(), fr.Y.data_ptr() if physics else None, st))
        ck(lib.fnx_grid_build(fr.X.data_ptr(), N, prm.H, fr.gridX.data_ptr(), st))
        if not fr.vis_grid_built:  # the un-advected visual particles are constant within a frame
            ck(lib.fnx_grid_build(fr.visual.data_ptr(), V, prm.H, fr.gridVis.data_ptr(), st))
            fr.vis_grid_built = True
        if physics and self.side is not None:
            self._ev_fork.record(main)
        # P1 forward straight into the fluid rows of the rasterizer's means3D (render units)
        # (neighbour count, cut-off and the advection sums in one walk: fnx_radius_count + fnx_visual_advect_fwd)
        ck(lib.fnx_visual_advect_fwd_counted(fr.gridX.data_ptr(), fr.X.data_ptr(), fr.xyz.data_ptr(), N, fr.visual.data_ptr(), V, prm.KNN_K,
                                             prm.H, prm.secs, prm.scale_factor, fr.means3D.data_ptr(), fr.num.data_ptr(),
                                             fr.den.data_ptr(), fr.kthV.data_ptr(), st))
        if not physics:
            fr.dX.zero_()
            fr.dDist.zero_()
            fr.scalars.zero_()
            return
        if self.side is not None:
            self._ev_means.record(main)
            self.side.wait_event(self._ev_fork)
            ctx = torch.cuda.stream(self.side)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            st = torch.cuda.current_stream(self.dev).cuda_stream
            for k, (grid, pos, kth, p, gp, lam, dpos) in enumerate((
                    (fr.gridX, fr.X, fr.kthX, fr.p, fr.gp, prm.lambda_gas_constraints, fr.dX),
                    (fr.gridY, fr.Y, fr.kthY, fr.pn, fr.gpn, prm.lambda_next_gas_constraints, fr.dY))):
                if k == 1:
                    ck(lib.fnx_grid_build(pos.data_ptr(), N, prm.H, grid.data_ptr(), st))
                ck(lib.fnx_pbf_density_fwd_counted(grid.data_ptr(), pos.data_ptr(), N, fr.imass.data_ptr(), prm.KNN_K, prm.H, prm.p0,
                                                   kth.data_ptr(), p.data_ptr(), fr.cap_flag[k:].data_ptr(), st))
                # l2_loss(p_ratio, 1) * lambda and its chain to the positions (the loss scalar and dL/dp_ratio are produced inside
                # the backward's own pre-pass)
                ck(lib.fnx_pbf_density_bwd_ratio(grid.data_ptr(), pos.data_ptr(), N, fr.imass.data_ptr(), kth.data_ptr(), prm.H, prm.p0,
                                                 p.data_ptr(), lam, fr.scalars[k:].data_ptr(), gp.data_ptr(), dpos.data_ptr(), 0, st))
            # P5 on the render-unit positions written by P1
            if prm.lambda_current_distance > 0:
                if self.side is not None:
                    self.side.wait_event(self._ev_means)
                thr = prm.distance_threshold_visual
                ck(lib.fnx_grid_build(fr.means3D.data_ptr(), V, thr, fr.gridP.data_ptr(), st))
                ck(lib.fnx_pair_distance_loss(fr.gridP.data_ptr(), fr.means3D.data_ptr(), V, thr, thr, prm.lambda_current_distance,
                                              fr.scalars[3:].data_ptr(), fr.dDist.data_ptr(), st))
            else:
                fr.dDist.zero_()
                fr.scalars[3:].zero_()
            if self.side is not None:
                self._ev_join.record(self.side)
        self._side_pending = self.side is not None

    def _view_mats(self, view_ids):
        key = tuple(view_ids)
        if key not in self._views:
            idx = torch.tensor(list(key), dtype=torch.long, device=self.dev)
            self._views[key] = (self.view_all[idx].contiguous(), self.proj_all[idx].contiguous())
        return self._views[key]

    def workspace(self, fr: FrameState, nviews, view_ids):
        """Persistent rasterizer buffers for this frame and this (sorted) set of cameras; sized from one exact forward (the only
        one that blocks).  With a frozen background set (3 channels) the static stream is binned ONCE per frame for all cameras
        (FrameState.static_stream) and shared by the workspaces of every camera subset an iteration may draw -- the reference
        samples random.sample(cur_viewpoint_set, batch) cameras per iteration (train_physical_particle.py:337); only the fluid
        rows are re-binned per iteration (MergedRasterWorkspace).  A workspace whose last forward overflowed its instance
        capacity is dropped here and rebuilt with a fresh sizing forward."""
        key = tuple(view_ids)
        ws = fr.ws.get(key)
        if ws is not None and ws.overflowed():   # the last finished forward overflowed: grow
            torch.cuda.synchronize(self.dev)
            fr.ws.pop(key)
            fr.graphs.clear()
            fr.skipped_iterations += 1
            self.capacity_margin = max(self.capacity_margin, 1.5)
            ws = None
        if ws is None:
            vm, pm = self._view_mats(view_ids)
            if self.static_cache and fr.Pb > 0 and self.C == 3:
                V = fr.V
                sl = lambda t, a, b: t[a:b]
                dyn = dict(means3D=sl(fr.means3D, 0, V), colors=sl(fr.colors, 0, V), opacities=sl(fr.opacity, 0, V),
                           scales=sl(fr.scales, 0, V), rotations=sl(fr.rotations, 0, V))
                if fr.static_stream is None:
                    sta = dict(means3D=sl(fr.means3D, V, fr.P), colors=sl(fr.colors, V, fr.P), opacities=sl(fr.opacity, V, fr.P),
                               scales=sl(fr.scales, V, fr.P), rotations=sl(fr.rotations, V, fr.P))
                    fr.static_stream = R.StaticStream(self.dev, self.view_all.size(0), self.H, self.W, self.bg, sta, self.view_all, self.proj_all,
                                                      self.tan_fov_x, self.tan_fov_y)
                ws = R.MergedRasterWorkspace(self.dev, V, nviews, self.H, self.W, self.bg, dyn, None, vm, pm, self.tan_fov_x,
                                             self.tan_fov_y, margin=self.capacity_margin, static_tile_cache=self.static_tile_cache,
                                             static_stream=fr.static_stream, view_ids=list(view_ids), slack=self.capacity_slack)
                ws.dyn = dyn
            else:
                ctx, _, _, _ = R.raster_forward(self.C, self.bg, fr.means3D, fr.colors, fr.opacity, fr.scales, fr.rotations, 1.0, None,
                                                vm, pm, self.tan_fov_x, self.tan_fov_y, self.H, self.W, speculative=False)
                cap = int(ctx.num_rendered * self.capacity_margin) + self.capacity_slack
                del ctx
                ws = R.RasterWorkspace(self.dev, self.C, fr.P, nviews, self.H, self.W, cap)
            fr.ws[key] = ws
            fr.graphs.clear()
        return ws

    def render(self, fr: FrameState, view_ids):
        ws = self.workspace(fr, len(view_ids), view_ids)
        first = ws.forwards == 0
        ws = self._render(fr, ws, view_ids)
        if first and self.lpt_order:
            # rank the tiles by the work of this first forward: every later blend launch (also the captured ones) starts the long
            # tiles first.  The scene of a frame barely moves between iterations, so the order is taken once.
            ws.update_tile_order()
        return ws

    def _render(self, fr: FrameState, ws, view_ids):
        if isinstance(ws, R.MergedRasterWorkspace):
            d = ws.dyn
            ws.forward(d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"])
            return ws
        vm, pm = self._view_mats(view_ids)
        # only the fluid rows [0, V) are trainable in the physical stage; the background set is frozen
        # (pipe_dynamics.py:51-57 concatenates it behind the fluid particles) and needs no gradient
        ws.forward(self.bg, fr.means3D, fr.colors, fr.opacity, fr.scales, fr.rotations, 1.0, vm, pm, self.tan_fov_x, self.tan_fov_y,
                   grad_range=(0, fr.V) if fr.Pb > 0 else None)
        return ws

    def image_loss(self, images, gt, batch, fr: FrameState = None):
        """Fused L1 + SSIM (+ grey conversion) and dL/dimage.  With `fr` the per-view means land in the frame's loss row."""
        prm, lib = self.prm, self.lib
        V, Cc, H, W = images.shape
        key = (V, Cc, H, W)
        if key not in self._loss_scratch:
            self._loss_scratch[key] = (torch.empty(lib.fnx_image_loss_bytes(V, Cc, H, W), dtype=torch.uint8, device=self.dev),
                                       torch.empty(V, device=self.dev), torch.empty(V, device=self.dev),
                                       torch.empty((V, Cc, H, W), device=self.dev))
        scratch, l1, ss, g = self._loss_scratch[key]
        if fr is not None and V <= MAX_ROW_VIEWS:
            l1, ss = fr.loss_row[4:4 + V], fr.loss_row[4 + MAX_ROW_VIEWS:4 + MAX_ROW_VIEWS + V]
        w_l1 = (1.0 - prm.lambda_dssim) * prm.lambda_image / batch
        w_ss = prm.lambda_dssim * prm.lambda_image / batch
        L.check(lib.fnx_image_loss(V, Cc, H, W, images.data_ptr(), gt.data_ptr(), int(prm.grey), w_l1, w_ss, g.data_ptr(),
                                   l1.data_ptr(), ss.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream(self.dev).cuda_stream))
        return l1, ss, g

    def physics_backward_and_update(self, fr: FrameState, dL_dmeans3D, grad_extra=None, update=True, physics=True, skip_flag=None):
        """P1-backward (image + distance gradients -> hidden particles), P3 chain, P4, Adam."""
        lib, prm, st = self.lib, self.prm, torch.cuda.current_stream(self.dev).cuda_stream
        N, V = fr.N, fr.V
        ck = L.check
        if self._side_pending:  # join the side stream: dX, dY, dDist and the loss scalars are complete after this
            torch.cuda.current_stream(self.dev).wait_event(self._ev_join)
            self._side_pending = False
        ck(lib.fnx_visual_advect_bwd(fr.gridVis.data_ptr(), fr.X.data_ptr(), fr.xyz.data_ptr(), N, V, fr.kthV.data_ptr(), fr.num.data_ptr(),
                                     fr.den.data_ptr(), dL_dmeans3D.data_ptr(), fr.dDist.data_ptr(), 1.0 / prm.scale_factor, prm.H,
                                     prm.secs, fr.dX.data_ptr(), 1, st))
        if update and grad_extra is None:
            # gradient assembly + Adam in one element-wise pass (the whole batch of views ran here: set_batch_gradient_current +
            # optimizer.step); voided on the device when the forward overflowed its instance capacity (skip_flag)
            ck(lib.fnx_pbf_combine_grad_adam(N, fr.e.data_ptr(), fr.buoyancy.data_ptr(), prm.secs, prm.buoyancy_max_y, prm.scale_factor,
                                             fr.dX.data_ptr(), fr.dY.data_ptr() if physics else None,
                                             fr.estimate_xyz.data_ptr() if physics else None, prm.lambda_exyz, fr.de.data_ptr(),
                                             fr.scalars[2:].data_ptr() if physics else None, fr.m.data_ptr(), fr.v.data_ptr(), prm.lr, 0.9,
                                             0.999, prm.adam_eps, fr.step_dev.data_ptr(), fr.bc_dev.data_ptr(), skip_flag, st))
            return
        ck(lib.fnx_pbf_combine_grad(N, fr.e.data_ptr(), fr.buoyancy.data_ptr(), prm.secs, prm.buoyancy_max_y, prm.scale_factor,
                                    fr.dX.data_ptr(), fr.dY.data_ptr() if physics else None,
                                    fr.estimate_xyz.data_ptr() if physics else None, prm.lambda_exyz, fr.de.data_ptr(),
                                    fr.scalars[2:].data_ptr() if physics else None, st))
        if grad_extra is not None:
            fr.de.add_(grad_extra)
        if update:
            self.adam(fr, skip_flag=skip_flag)

    def adam(self, fr: FrameState, grad=None, grad_scale=1.0, skip_flag=None):
        """torch.optim.Adam(eps=1e-15) step on the frame's trainable tensor.  skip_flag: device address of an int32 that voids the
        update when non-zero (the rasterizer's overflow flag: a forward that overflowed its instance capacity renders only the
        background, its image gradient is zero and the physics-only update must not be applied)."""
        g = fr.de if grad is None else grad
        L.check(self.lib.fnx_adam_step_dev_gated(fr.e.numel(), fr.e.data_ptr(), g.data_ptr(), fr.m.data_ptr(), fr.v.data_ptr(), grad_scale,
                                                 self.prm.lr, 0.9, 0.999, self.prm.adam_eps, fr.step_dev.data_ptr(), fr.bc_dev.data_ptr(),
                                                 skip_flag, torch.cuda.current_stream(self.dev).cuda_stream))

    # -- the step -------------------------------------------------------------------------------------------
    def _iteration(self, fr: FrameState, view_ids, gt, update, batch, physics=True):
        self.physics_forward(fr, physics)
        out, skip = {}, None
        if len(view_ids):
            ws = self.render(fr, view_ids)
            l1, ss, g = self.image_loss(ws.color, gt, batch, fr)
            dmeans = ws.backward(g)["means3D"]
            skip = ws.overflow_flag()
            out.update(l1=l1, ssim=ss, images=ws.color, radii=ws.radii, ws=ws, view_ids=list(view_ids))
        else:
            if fr.zero_dmeans is None:
                fr.zero_dmeans = torch.zeros((fr.P, 3), device=self.dev)
            dmeans = fr.zero_dmeans
        self.physics_backward_and_update(fr, dmeans, update=update, physics=physics, skip_flag=skip)
        out.update(gas=fr.scalars[0], next_gas=fr.scalars[1], exyz=fr.scalars[2], dist=fr.scalars[3], grad=fr.de)
        return out

    def _cached_gt(self, gt):
        return self._gt_cache.get(gt, lambda t: t.to(self.dev, non_blocking=True).float().contiguous())

    def step(self, fr: FrameState, view_ids, gt, update=True, batch=None, graph=False, physics=True, cache_gt=None):
        """One optimiser iteration for one frame.  view_ids: camera indices rendered by THIS process; gt
        [len(view_ids),C,H,W] (device tensor, or a pinned host tensor); `batch` = global number of views of the step
        (defaults to len(view_ids); larger when views are sharded over ranks); physics=False skips the
        view-independent terms (another rank owns them).  graph=True captures the iteration
        into a CUDA graph on first use and replays it afterwards (one host call per iteration).
        Returns device tensors that are overwritten by the next step; there is no host synchronisation."""
        batch = len(view_ids) if batch is None else batch
        if not gt.is_cuda and (self.cache_gt if cache_gt is None else cache_gt):
            gt = self._cached_gt(gt)                      # a host tensor seen before: its device copy (uploaded once)
        # canonical camera order: the workspaces / captured graphs of a frame are keyed by the SET of cameras, so drawing the
        # same cameras in another order (random.sample) reuses them; the per-view outputs (l1, ssim, images) come back in
        # ascending camera order (out["view_ids"])
        ids = [int(v) for v in view_ids]
        perm = sorted(range(len(ids)), key=ids.__getitem__)
        if perm != list(range(len(ids))):
            view_ids, gt = [ids[k] for k in perm], gt[perm]
        else:
            view_ids = ids
        with torch.cuda.device(self.dev):
            if not graph:
                if not gt.is_cuda:
                    gt = gt.to(self.dev, non_blocking=True)
                return self._iteration(fr, view_ids, gt, update, batch, physics)
            key = (tuple(view_ids), bool(update), batch, bool(physics))
            ent = fr.graphs.get(key)
            if ent is not None:
                ws = ent["slots"][0][1].get("ws")
                if ws is not None and ws.overflowed():
                    # A replay overflowed the workspace's instance capacity (the fluid moved into many more tiles).  That
                    # iteration rendered only the background and its update was voided ON THE DEVICE (adam: skip_flag), so the
                    # state is intact: drop the workspace and its graphs, re-size, re-capture and carry on.
                    self.workspace(fr, len(view_ids), view_ids)
                    ent = None
            if ent is None:
                # everything that allocates or blocks happens eagerly first (workspace sizing, visual grid, scratch)
                gt_bufs = [torch.empty((len(view_ids), self.C, self.H, self.W), device=self.dev) for _ in range(2)]
                gt_bufs[0].copy_(gt, non_blocking=True)
                snap = (fr.e.clone(), fr.m.clone(), fr.v.clone(), fr.step_dev.clone())
                self._iteration(fr, view_ids, gt_bufs[0], update, batch, physics)  # eager warm-up (also sizes the workspace)
                for dst, src in zip((fr.e, fr.m, fr.v, fr.step_dev), snap):       # undo its parameter update
                    dst.copy_(src)
                torch.cuda.synchronize(self.dev)
                # The iteration is captured TWICE, once per ground-truth buffer, and the two graphs alternate: the upload
                # of the next iteration's ground truth (copy stream) then never waits for the current iteration, which is
                # still reading the other buffer.  Everything else the two graphs touch is the same memory.
                slots = []
                for gt_buf in gt_bufs:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        out = self._iteration(fr, view_ids, gt_buf, update, batch, physics)
                    slots.append((g, out, gt_buf, torch.cuda.Event(), torch.cuda.Event()))
                    for dst, src in zip((fr.e, fr.m, fr.v, fr.step_dev), snap):   # capture does not execute, but be explicit
                        dst.copy_(src)
                ent = {"slots": slots, "next": 0}
                fr.graphs[key] = ent
            g, out, gt_buf, ev_copied, ev_done = ent["slots"][ent["next"]]
            ent["next"] ^= 1
            main = torch.cuda.current_stream(self.dev)
            if gt.is_cuda:
                gt_buf.copy_(gt, non_blocking=True)
            else:
                cs = self.copy_stream
                cs.wait_event(ev_done)          # the previous replay of THIS graph (two iterations ago) has finished reading gt_buf
                with torch.cuda.stream(cs):
                    gt_buf.copy_(gt, non_blocking=True)
                    ev_copied.record(cs)
                main.wait_event(ev_copied)
            g.replay()
            ev_done.record(main)
            return out

    def total_loss_from_rows(self, rows, batch):
        """The same from loss rows [..., LOSS_ROW] (FrameState.loss_row; summed over ranks when a frame's views are
        sharded): `batch` = number of views of the frame's iteration."""
        prm = self.prm
        l1 = rows[..., 4:4 + MAX_ROW_VIEWS].sum(-1)
        ss = rows[..., 4 + MAX_ROW_VIEWS:4 + 2 * MAX_ROW_VIEWS].sum(-1)
        img = ((1.0 - prm.lambda_dssim) * prm.lambda_image * l1 + prm.lambda_dssim * prm.lambda_image * (batch - ss)) / batch
        return (img + prm.lambda_current_distance * rows[..., 3] + prm.lambda_exyz * rows[..., 2]
                + prm.lambda_gas_constraints * rows[..., 0] + prm.lambda_next_gas_constraints * rows[..., 1])

    def total_loss(self, out, batch=None):
        """The reference's per-view `loss`, averaged over the views (device scalar)."""
        prm = self.prm
        b = out["l1"].numel() if batch is None else batch
        img = ((1.0 - prm.lambda_dssim) * prm.lambda_image * out["l1"] + prm.lambda_dssim * prm.lambda_image * (1.0 - out["ssim"])).sum() / b
        return (img + prm.lambda_current_distance * out["dist"] + prm.lambda_exyz * out["exyz"]
                + prm.lambda_gas_constraints * out["gas"] + prm.lambda_next_gas_constraints * out["next_gas"])
