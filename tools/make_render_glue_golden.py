"""Run in the build container (needs /root/reference): calls the reference's OWN render glue (renderer/pipe_fluid.py,
pipe_dynamics.py, pipe_background.py) on the CPU with a recording stand-in for the rasterizer classes, and stores what the
glue handed to the rasterizer and what it returned (tests/golden/pyref_render_glue.npz).  The glue modules import the model
classes, which import torch_cluster & co.: stubbed as in tools/make_physics_golden.py; device="cuda" is redirected."""
import math
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/FluidDynamics"
OUT = os.path.join(ROOT, "tests", "golden", "pyref_render_glue.npz")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from fluidnexus_b200 import synthetic as S  # noqa: E402
from make_physics_golden import cuda_as_cpu, install_stubs  # noqa: E402

CASES = {  # name -> (function, kwargs)
    "fluid_visual": ("render_fluid", dict(pos_type="visual", scale=True)),
    "fluid_nn": ("render_fluid", dict(pos_type="guess_visual_nn", scale=True)),
    "fluid_hidden": ("render_fluid", dict(pos_type="hidden", scale=False, scaling_modifier=0.7)),
    "dyn_full": ("render_dynamics", dict(pos_type="guess_visual_nn", scale=True)),
    "dyn_gpf": ("render_dynamics", dict(pos_type="visual", scale=True, gpf_only=True)),
    "dyn_gs": ("render_dynamics", dict(pos_type="visual", gs_only=True)),
    "background": ("render_background", dict()),
}


class FakeModel:
    """The accessors the pipes read, CPU tensors, seeded (shared with the test through build_model)."""
    scale_factor, active_sh_degree = 100.0, 0

    def __init__(self, seed=0):
        f = S.fluid_gaussians(40, 1, seed=seed).torch("cpu")
        self.get_visual_xyz = f["xyz"] * self.scale_factor
        self.get_visual_opacity, self.get_visual_scaling, self.get_visual_rotation, self.get_visual_color = f["opacity"], f["scales"], f["rotations"], f["colors"]
        h = S.fluid_gaussians(25, 1, seed=seed + 1).torch("cpu")
        self.get_xyz_hidden = h["xyz"]
        self.get_opacity_dummy, self.get_scaling_dummy, self.get_rotation_dummy, self.get_color_dummy = h["opacity"], h["scales"], h["rotations"], h["colors"]
        g = S.background_gaussians(30, 3, seed=seed + 2).torch("cpu")
        self.get_gs_xyz, self.get_gs_opacity, self.get_gs_scaling, self.get_gs_rotation, self.get_gs_color = g["xyz"], g["opacity"], g["scales"], g["rotations"], g["colors"]
        self.background = False

    def get_visual_xyz_from_nn(self):
        return self.get_visual_xyz + 0.25

    # render_fluid(pos_type="hidden") and render_background both read get_xyz (different model classes in the reference)
    @property
    def get_xyz(self):
        return self.get_gs_xyz if self.background else self.get_xyz_hidden

    get_opacity = property(lambda s: s.get_gs_opacity)
    get_scaling = property(lambda s: s.get_gs_scaling)
    get_rotation = property(lambda s: s.get_gs_rotation)
    get_color = property(lambda s: s.get_gs_color)


def recording_rasterizer(log):
    from typing import NamedTuple

    class Settings(NamedTuple):
        image_height: int
        image_width: int
        tan_fov_x: float
        tan_fov_y: float
        bg: torch.Tensor
        scale_modifier: float
        view_matrix: torch.Tensor
        proj_matrix: torch.Tensor
        sh_degree: int
        campos: torch.Tensor
        prefiltered: bool

    class Rasterizer:
        def __init__(self, raster_settings):
            self.rs = raster_settings

        def __call__(self, **kw):
            log.clear()
            log.update({k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in kw.items()})
            log["settings"] = self.rs
            P = kw["means3D"].shape[0]
            C = kw["colors_precomp"].shape[1]
            img = torch.full((C, self.rs.image_height, self.rs.image_width), 0.5)
            return img, (torch.arange(P) % 3).int(), torch.full((1, self.rs.image_height, self.rs.image_width), 2.0)
    return Settings, Rasterizer


def run_case(fn, gm, cam, kwargs):
    log = {}
    Settings, Rasterizer = recording_rasterizer(log)
    gm.background = fn.__name__ == "render_background"
    out = fn(cam, gm, None, torch.tensor([0.1, 0.2, 0.3]), GRsetting=Settings, GRzer=Rasterizer, **kwargs)
    rs = log["settings"]
    rec = {f"in_{k}": log[k].numpy() for k in ("means3D", "means2D", "colors_precomp", "opacities", "scales", "rotations")}
    rec.update(in_shs_none=log["shs"] is None, in_cov_none=log["cov3D_precomp"] is None, rs_hw=np.array([rs.image_height, rs.image_width]),
               rs_tan=np.array([rs.tan_fov_x, rs.tan_fov_y]), rs_bg=rs.bg.numpy(), rs_scale_modifier=rs.scale_modifier,
               rs_view=rs.view_matrix.numpy(), rs_proj=rs.proj_matrix.numpy(), rs_sh_degree=rs.sh_degree, rs_campos=rs.campos.numpy(),
               rs_prefiltered=rs.prefiltered, out_keys=np.array(sorted(out.keys())))
    for k in ("render", "radii", "depth", "render_xyz", "raw_render_xyz", "means3D", "opacity", "rotations", "colors_precomp", "scales",
              "visibility_filter"):
        rec[f"out_{k}"] = out[k].detach().numpy()
    return rec


def main():
    install_stubs()
    # renderer/__init__.py also pulls in pipe.py (upstream 3DGS rasterizer, selected by no config) and the background model (plyfile)
    dgr = types.ModuleType("diff_gaussian_rasterization")
    dgr.GaussianRasterizationSettings, dgr.GaussianRasterizer = object, object
    ply = types.ModuleType("plyfile")
    ply.PlyData, ply.PlyElement = object, object
    sys.modules.update({"diff_gaussian_rasterization": dgr, "plyfile": ply})
    sys.path.insert(0, REF)
    with cuda_as_cpu():
        from renderer.pipe_background import render_background
        from renderer.pipe_dynamics import render_dynamics
        from renderer.pipe_fluid import render_fluid
        fns = dict(render_fluid=render_fluid, render_dynamics=render_dynamics, render_background=render_background)
        gm, cam = FakeModel(), S.make_cameras(5, 32, height=24)[1]
        out = {}
        for name, (fn, kw) in CASES.items():
            for k, v in run_case(fns[fn], gm, cam, kw).items():
                out[f"{name}__{k}"] = v
        # ---- the reference's Camera class itself (scene/camera.py:14-110; kornia's meshgrid is only needed for rays) ----
        kornia = types.ModuleType("kornia")
        kornia.create_meshgrid = lambda *a, **k: None
        sys.modules["kornia"] = kornia
        from scene.camera import Camera
        R = np.array([[0.36, 0.48, -0.8], [-0.8, 0.6, 0.0], [0.48, 0.64, 0.6]])
        T = np.array([0.1, -0.2, 1.3])
        img = torch.rand(3, 24, 32, generator=torch.Generator().manual_seed(3)) * 1.4 - 0.2        # exercises the clamp
        c = Camera(colmap_id=1, R=R, T=T, FoVx=0.69, FoVy=0.55, image=img, gt_alpha_mask=None, image_name="view_00", uid=0,
                   real_image=img.clone(), timestamp=0.5)
        out.update(camera__R=R, camera__T=T, camera__image_in=img.numpy(), camera__original_image=c.original_image.numpy(),
                   camera__hw=np.array([c.image_height, c.image_width]), camera__world_view_transform=c.world_view_transform.numpy(),
                   camera__projection_matrix=c.projection_matrix.numpy(), camera__full_proj_transform=c.full_proj_transform.numpy(),
                   camera__camera_center=c.camera_center.numpy(), camera__znear_zfar=np.array([c.z_near, c.z_far]),
                   camera__fov=np.array([c.FoVx, c.FoVy]), camera__timestamp=c.timestamp)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()
