"""GPU: the render glue mirrors (fluidnexus_b200.renderer) against direct rasterizer calls on the attribute sets the
reference's pipes would select (FD/renderer/pipe_fluid.py, pipe_dynamics.py, pipe_background.py), including the
screen-space gradient that the densification statistics read from `viewspace_points.grad`."""
import math

import numpy as np
import pytest
import torch

from fluidnexus_b200 import rasterizer as R
from fluidnexus_b200 import renderer as RD
from fluidnexus_b200 import synthetic as S

pytestmark = pytest.mark.gpu


class FakeModel:
    """The accessors of the reference's GaussianModel that the pipes read."""
    scale_factor = 100.0
    active_sh_degree = 0

    def __init__(self, C, with_gs):
        d = "cuda"
        f = S.fluid_gaussians(700, C, seed=1, log_scale=-4.6).torch(d)
        self.vis_scaled = (f["xyz"] * self.scale_factor).requires_grad_(True)     # scaled units, like _visual_xyz
        self.get_visual_opacity, self.get_visual_scaling = f["opacity"], f["scales"]
        self.get_visual_rotation, self.get_visual_color = f["rotations"], f["colors"].requires_grad_(True)
        h = S.fluid_gaussians(300, C, seed=2).torch(d)
        self.get_xyz = h["xyz"]
        self.get_opacity_dummy, self.get_scaling_dummy = h["opacity"], h["scales"]
        self.get_rotation_dummy, self.get_color_dummy = h["rotations"], h["colors"]
        if with_gs:
            g = S.background_gaussians(900, 3, seed=3).torch(d)
            self.get_gs_xyz, self.get_gs_opacity, self.get_gs_scaling = g["xyz"], g["opacity"], g["scales"]
            self.get_gs_rotation, self.get_gs_color = g["rotations"], g["colors"]

    @property
    def get_visual_xyz(self):
        return self.vis_scaled

    def get_visual_xyz_from_nn(self):
        return self.vis_scaled + 0.01


def _direct(C, cam, bg, xyz, colors, opacity, scales, rotations):
    _, img, radii, depth = R.raster_forward(C, bg, xyz.detach().float().contiguous(), colors.detach().float().contiguous(),
                                            opacity.detach().reshape(-1).contiguous(), scales.contiguous(), rotations.contiguous(), 1.0, None,
                                            cam.world_view_transform, cam.full_proj_transform, math.tan(cam.FoVx * 0.5),
                                            math.tan(cam.FoVy * 0.5), cam.image_height, cam.image_width, speculative=False)
    return img, radii, depth   # single camera: [C,H,W], [P], [1,H,W]


def test_render_fluid_selects_positions_and_attributes(libfnx):
    gm, cam = FakeModel(1, False), S.make_cameras(5, 96, device="cuda")[1]
    bg = torch.zeros(3, device="cuda")
    out = RD.render_fluid(cam, gm, None, bg, pos_type="guess_visual_nn", scale=True)
    img, radii, depth = _direct(1, cam, bg, (gm.vis_scaled + 0.01) / 100.0, gm.get_visual_color, gm.get_visual_opacity,
                                gm.get_visual_scaling, gm.get_visual_rotation)
    assert out["render"].shape == (1, 96, 96) and torch.equal(out["render"], img)
    assert torch.equal(out["radii"], radii) and torch.equal(out["depth"], depth) and torch.equal(out["visibility_filter"], radii > 0)
    assert torch.equal(out["raw_render_xyz"], gm.vis_scaled + 0.01) and out["means3D"] is out["render_xyz"]
    out["render"].sum().backward()
    assert gm.vis_scaled.grad is not None and float(gm.vis_scaled.grad.abs().max()) > 0           # through the /scale_factor
    assert out["viewspace_points"].grad is not None and out["viewspace_points"].grad.shape == (700, 3)
    assert float(out["viewspace_points"].grad[:, :2].abs().max()) > 0 and float(out["viewspace_points"].grad[:, 2].abs().max()) == 0
    hid = RD.render_fluid(cam, gm, None, bg, pos_type="hidden")
    img_h, _, _ = _direct(1, cam, bg, gm.get_xyz, gm.get_color_dummy, gm.get_opacity_dummy, gm.get_scaling_dummy, gm.get_rotation_dummy)
    assert torch.equal(hid["render"], img_h)
    with pytest.raises(ValueError, match="Unknown pos_type"):
        RD.render_fluid(cam, gm, None, bg, pos_type="nope")


def test_render_dynamics_concatenates_and_background_renders(libfnx):
    gm, cam = FakeModel(1, True), S.make_cameras(5, 80, device="cuda")[3]
    bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")
    cat = lambda a, b: torch.cat([a, b], 0)
    grey3 = gm.get_visual_color.detach().repeat(1, 3)
    full = RD.render_dynamics(cam, gm, None, bg, pos_type="visual", scale=True)
    img, radii, _ = _direct(3, cam, bg, cat(gm.vis_scaled / 100.0, gm.get_gs_xyz), cat(grey3, gm.get_gs_color),
                            cat(gm.get_visual_opacity, gm.get_gs_opacity), cat(gm.get_visual_scaling, gm.get_gs_scaling),
                            cat(gm.get_visual_rotation, gm.get_gs_rotation))
    assert full["render"].shape == (3, 80, 80) and torch.equal(full["render"], img) and full["radii"].shape == (1600,)
    assert full["render_xyz"].shape == (700, 3) and full["means3D"].shape == (1600, 3) and full["colors_precomp"].shape == (1600, 3)
    gpf = RD.render_dynamics(cam, gm, None, bg, pos_type="visual", scale=True, gpf_only=True)
    img_f, _, _ = _direct(3, cam, bg, gm.vis_scaled / 100.0, grey3, gm.get_visual_opacity, gm.get_visual_scaling, gm.get_visual_rotation)
    assert torch.equal(gpf["render"], img_f)
    gs = RD.render_dynamics(cam, gm, None, bg, gs_only=True)
    class BackgroundModel:   # gm_background.GaussianModel's accessors
        active_sh_degree = 0
        get_xyz, get_opacity, get_scaling, get_rotation, get_color = gm.get_gs_xyz, gm.get_gs_opacity, gm.get_gs_scaling, gm.get_gs_rotation, gm.get_gs_color
    back = RD.render_background(cam, BackgroundModel(), None, bg)
    img_b, _, _ = _direct(3, cam, bg, gm.get_gs_xyz, gm.get_gs_color, gm.get_gs_opacity, gm.get_gs_scaling, gm.get_gs_rotation)
    assert torch.equal(gs["render"], img_b) and torch.equal(back["render"], img_b)
    assert set(back) == set(full)
