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


# ======================================================================================================================
# Stock glue: the reference's OWN Python (oracle/_ref/FluidDynamics, see oracle/build_ref.py:stage_python) drives the iteration
# ======================================================================================================================
def make_host_physics_model(model_cls):
    """The reference keeps every tensor on cuda:0 and reaches torch_cluster's CUDA kernels from there.  torch_cluster is
    not installable here and north_star asks for the reference's physics-loss path on the HOST cores, so the particle state
    and the trainable tensor live on the CPU while everything the rasterizer sees lives on the GPU.  Three methods of the
    stock model hard-wire the device; this subclass re-places their tensors and changes nothing else:
      get_visual_xyz_from_nn              stock result, moved to the GPU (autograd-aware H2D) for the stock render pipe
      zero_gradient_cache_current         the stock body allocates the cache with device="cuda" (gm_fluid.py:419-421)
      (cache_/set_batch_gradient_current  are inherited: they then add / scale CPU tensors)"""
    class HostPhysicsModel(model_cls):
        _gpu_leaf = None

        def get_visual_xyz_from_nn(self):
            if self._gpu_leaf is not None:            # gpu_part(): positions given on the GPU, no host physics
                return self._gpu_leaf * self.scale_factor
            return super().get_visual_xyz_from_nn().to("cuda")

        def zero_gradient_cache_current(self):
            self._estimate_xyz_nn_grad = torch.zeros_like(self._estimate_xyz_nn)
    return HostPhysicsModel


class StockTrainer:
    """bench.py's reference arm on the reference's own code: stock GaussianModel (gm_dynamics / gm_fluid) built by its
    constructor + setup_constants on the stock JSON config, stock Camera objects holding the ground truth on the CPU,
    stock render pipe -> the reference's wrapper package -> the compiled, unmodified reference rasterizer (oracle/_ref),
    stock loss_utils on the GPU, stock torch.optim.Adam; one iteration = the body of the hot loop of the stock entry script,
    executed as it is (oracle/ref_python.loop_body)."""

    CASES = {3: ("fluid_nexus_smoke_dynamics", "fluid_nexus_physical_current", "render_dynamics"),
             1: ("scalar_real", "scalar_real_physical_current", "render_fluid")}

    def __init__(self, prm, hidden, visual_scaled, fluid, background, channels, cams, gts_cpu, with_distance=True, config=None,
                 native="reference"):
        """native="reference": the reference arm (compiled reference rasterizer, particle state and physics on the host).
        native="fnx": the same stock Python with the five plugin imports bound to libfnx's drop-ins and EVERYTHING on the GPU,
        as the reference runs it with torch_cluster installed -- what a user of the unchanged scripts gets from the drop-ins."""
        import random
        import tempfile

        from . import ref_python as RP
        RP.use_reference_python(native)
        from helpers.helper_gaussian import get_model
        from helpers.helper_pipe import get_render_pipe
        from scene.camera import Camera
        from utils.loss_utils import distance_loss, l1_loss, l2_loss, ssim
        cfg_name, body, pipe = self.CASES[channels]
        cfg_name = config or cfg_name
        args, model_args, optim_args, pipe_args = RP.parse_args(cfg_name, tempfile.mkdtemp(prefix="fnx_ref_"))
        # the synthetic workload's constants (bench.WORKLOADS) on top of the stock config
        optim_args.p0, optim_args.buoyancy_max_y = prm.p0, prm.buoyancy_max_y
        optim_args.distance_threshold_visual = prm.distance_threshold_visual
        optim_args.batch = len(cams)
        gm = make_host_physics_model(get_model(model_args.model))(model_args.sh_degree)   # (on the GPU the overrides are no-ops)
        gm.setup_constants(optim_args)
        gm.spatial_lr_scale = 1.0
        pdev = "cpu" if native == "reference" else "cuda"
        f32 = lambda a: torch.as_tensor(a, dtype=torch.float32).to(pdev)
        gm._xyz, gm._estimate_xyz = f32(hidden.xyz), f32(hidden.estimate_xyz)
        gm._velocity, gm._force, gm._buoyancy, gm._imass = f32(hidden.velocity), f32(hidden.force), f32(hidden.buoyancy), f32(hidden.imass)
        gm._counts = torch.zeros((hidden.N, 1), device=pdev)
        gm._visual_xyz = f32(visual_scaled)
        gm.training_setup_current(optim_args)           # the Parameter lives where _estimate_xyz lives
        # rendering attributes of the fluid particles and the frozen background set: on the GPU, as raw (pre-activation) values
        dev = "cuda"
        g = lambda s, k: f32(getattr(s, k)).to(dev)
        gm._visual_color = g(fluid, "colors")[:, :1].contiguous() if pipe == "render_dynamics" else g(fluid, "colors")
        gm._visual_scales, gm._visual_rotation = torch.log(g(fluid, "scales")), g(fluid, "rotations")
        gm._visual_opacity = torch.logit(g(fluid, "opacity"))
        if background is not None:
            gm._gs_xyz, gm._gs_color = g(background, "xyz"), g(background, "colors")
            gm._gs_scales, gm._gs_rotation, gm._gs_opacity = torch.log(g(background, "scales")), g(background, "rotations"), torch.logit(g(background, "opacity"))
        elif pipe == "render_dynamics":                  # no frozen set: zero-row tensors where load_ply would have put it
            z = lambda *sh: torch.zeros(sh, device=dev)
            gm._gs_xyz, gm._gs_color, gm._gs_scales, gm._gs_rotation, gm._gs_opacity = z(0, 3), z(0, 3), z(0, 3), z(0, 4), z(0, 1)
        self.gm, self.optim_args, self.pipe_args = gm, optim_args, pipe_args
        self.render_func, self.GRsetting, self.GRzer = get_render_pipe(pipe)
        self.background = torch.zeros(3 if pipe == "render_dynamics" else 1, device=dev)
        self.cams = [Camera(colmap_id=k, R=c.R, T=c.T, FoVx=c.FoVx, FoVy=c.FoVy, image=gt, gt_alpha_mask=None, image_name=f"train0{k}", uid=k,
                            real_image=gt.clone()) for k, (c, gt) in enumerate(zip(cams, gts_cpu))]
        self.code, self.where = RP.loop_body(body)
        self.tb = RP.NullWriter()
        self.with_distance = with_distance
        if not with_distance:
            optim_args.lambda_current_distance = 0.0     # the FluidNexus body then skips the dense cdist (O(V^2) memory)
            zero = lambda positions, threshold: torch.zeros((), device=positions.device)
        self.ns = dict(gaussians=gm, optim_args=optim_args, random=random, cur_viewpoint_set=self.cams, render_func=self.render_func,
                       pipe_args=pipe_args, background=self.background, GRsetting=self.GRsetting, GRzer=self.GRzer, torch=torch,
                       l1_loss=l1_loss, ssim=ssim, distance_loss=distance_loss if with_distance else zero, l2_loss=l2_loss,
                       tb_writer=self.tb, cur_time_index=1)
        self.l1_loss, self.ssim = l1_loss, ssim
        self.itr = 0
        self.grey = pipe == "render_dynamics"

    def render_gt(self, cam_index, render_xyz_gpu):
        """Ground truth through the reference rasterizer itself (positions given in render units on the GPU)."""
        gm = self.gm
        gm._gpu_leaf = render_xyz_gpu
        try:
            with torch.no_grad():
                pkg = self.render_func(self.cams[cam_index], gm, self.pipe_args, self.background, GRsetting=self.GRsetting, GRzer=self.GRzer,
                                       pos_type="guess_visual_nn", scale=True)
            return pkg["render"].detach()
        finally:
            gm._gpu_leaf = None

    def set_ground_truth(self, gts_cpu):
        for cam, gt in zip(self.cams, gts_cpu):
            cam.original_image = gt.clamp(0.0, 1.0)

    def iteration(self):
        self.itr += 1
        self.ns["itr"] = self.itr
        exec(self.code, self.ns)
        return self.tb.scalars

    def gpu_part(self, gts_gpu, render_xyz_gpu):
        """Only what the reference runs on the GPU for one iteration: per view the stock render pipe + stock image losses +
        backward to the Gaussian means (no host physics, no .item()): the tightest comparison with libfnx's rasterizer + loss
        kernels."""
        gm, oa = self.gm, self.optim_args
        leaf = render_xyz_gpu.detach().clone().requires_grad_(True)
        gm._gpu_leaf = leaf
        try:
            for cam, gt in zip(self.cams, gts_gpu):
                pkg = self.render_func(cam, gm, self.pipe_args, self.background, GRsetting=self.GRsetting, GRzer=self.GRzer,
                                       pos_type="guess_visual_nn", scale=True)
                image, gt_image = pkg["render"], gt
                if self.grey:
                    gt_image = torch.cat([torch.mean(gt_image, dim=0, keepdim=True)] * 3, dim=0)
                    image = torch.cat([torch.mean(image, dim=0, keepdim=True)] * 3, dim=0)
                loss = (1.0 - oa.lambda_dssim) * self.l1_loss(image, gt_image) * oa.lambda_image + oa.lambda_dssim * (1.0 - self.ssim(image, gt_image)) * oa.lambda_image
                loss.backward()
        finally:
            gm._gpu_leaf = None
        return leaf.grad
