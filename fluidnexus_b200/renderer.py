"""Render glue of the reference on top of libfnx's rasterizer.

Same call signatures, `pos_type` / `scale` / `gpf_only` / `gs_only` semantics and return dictionaries as

    render_fluid        FD/renderer/pipe_fluid.py:8-135        (1-channel fluid scenes, ScalarReal)
    render_dynamics     FD/renderer/pipe_dynamics.py:8-180     (fluid particles ++ frozen background, grey -> RGB)
    render_background   FD/renderer/pipe_background.py:9-95    (static background training)

so the entries can import them from here instead of `renderer.*` -- or keep their own copies and only shadow the
rasterizer modules.  `gm` is duck-typed: any object with the reference GaussianModel's
accessors (`get_visual_xyz`, `get_visual_opacity`, ..., `scale_factor`, `active_sh_degree`).  `GRsetting` / `GRzer` default
to the libfnx drop-ins of the channel count the colours have; passing the reference's own classes works as well.

What differs from the reference glue is only cost: tensors that already are contiguous fp32 are not copied by `.float()`,
and the `[dynamic ; static]` concatenations of render_dynamics are built once per call with a single allocation per
attribute."""
import math

import torch

from . import rasterizer as R

_MODULES = {}


def _raster_classes(channels):
    if channels not in _MODULES:
        _MODULES[channels] = R.make_module(channels)[:2]
    return _MODULES[channels]


_POSITIONS = {  # pos_type -> accessor of the positions (pipe_fluid.py:27-40)
    "guess_visual_nn": lambda gm: gm.get_visual_xyz_from_nn(),
    "guess_visual_hidden": lambda gm: gm.get_visual_xyz_from_hidden_guess(),
    "visual": lambda gm: gm.get_visual_xyz,
    "hidden": lambda gm: gm.get_xyz,
    "rigid": lambda gm: gm.get_rigid_xyz,
    "re_sim_visual": lambda gm: gm.get_re_sim_visual_xyz,
}
_ATTRIBUTES = {  # pos_type -> (opacity, scaling, rotation, colour) accessor names (pipe_fluid.py:79-103); default: visual
    "hidden": ("get_opacity_dummy", "get_scaling_dummy", "get_rotation_dummy", "get_color_dummy"),
    "rigid": ("get_rigid_opacity", "get_rigid_scaling", "get_rigid_rotation", "get_rigid_color"),
    "high": ("get_high_opacity", "get_high_scaling", "get_high_rotation", "get_high_color"),
    "dense": ("get_dense_opacity", "get_dense_scaling", "get_dense_rotation", "get_dense_color"),
}
_VISUAL = ("get_visual_opacity", "get_visual_scaling", "get_visual_rotation", "get_visual_color")


def _positions(gm, pos_type, scale):
    if pos_type not in _POSITIONS:
        raise ValueError(f"Unknown pos_type: {pos_type}")
    raw = _POSITIONS[pos_type](gm)
    return raw, (raw / gm.scale_factor if scale else raw)


def _attributes(gm, pos_type):
    return tuple(getattr(gm, name) for name in _ATTRIBUTES.get(pos_type, _VISUAL))


def _settings(GRsetting, cam, bg_color, scaling_modifier, sh_degree):
    return GRsetting(image_height=int(cam.image_height), image_width=int(cam.image_width), tan_fov_x=math.tan(cam.FoVx * 0.5),
                     tan_fov_y=math.tan(cam.FoVy * 0.5), bg=bg_color.float(), scale_modifier=scaling_modifier,
                     view_matrix=cam.world_view_transform, proj_matrix=cam.full_proj_transform, sh_degree=sh_degree,
                     campos=cam.camera_center, prefiltered=False)


def _rasterize(cam, gm, bg_color, scaling_modifier, GRsetting, GRzer, means3D, opacity, scales, rotations, colors_precomp,
               render_xyz, raw_render_xyz):
    if GRsetting is None or GRzer is None:
        GRsetting, GRzer = _raster_classes(int(colors_precomp.shape[1]))
    # zero tensor that receives the screen-space gradient (densification statistics read its .grad)
    means2D = torch.zeros_like(means3D, dtype=means3D.dtype, requires_grad=True, device=means3D.device) + 0
    try:
        means2D.retain_grad()
    except Exception:
        pass
    rasterizer = GRzer(raster_settings=_settings(GRsetting, cam, bg_color, scaling_modifier, gm.active_sh_degree))
    image, radii, depth = rasterizer(means3D=means3D.float(), means2D=means2D.float(), shs=None, colors_precomp=colors_precomp.float(),
                                     opacities=opacity.float(), scales=scales.float(), rotations=rotations.float(), cov3D_precomp=None)
    return {"render": image, "viewspace_points": means2D, "visibility_filter": radii > 0, "radii": radii, "opacity": opacity,
            "depth": depth, "render_xyz": render_xyz, "raw_render_xyz": raw_render_xyz, "means3D": means3D, "means2D": means2D,
            "rotations": rotations, "colors_precomp": colors_precomp, "scales": scales}


def render_fluid(viewpoint_camera, gm, pipe_args, bg_color, scaling_modifier=1.0, override_color=None, GRsetting=None, GRzer=None,
                 pos_type="visual", scale=False, prev_visual_xyz=None, **kwargs):
    """pipe_fluid.py:8-135.  Background tensor (bg_color) must be on the GPU."""
    raw, xyz = _positions(gm, pos_type, scale)
    opacity, scales, rotations, colors = _attributes(gm, pos_type)
    return _rasterize(viewpoint_camera, gm, bg_color, scaling_modifier, GRsetting, GRzer, xyz, opacity, scales, rotations, colors, xyz, raw)


def render_dynamics(viewpoint_camera, gm, pipe_args, bg_color, scaling_modifier=1.0, override_color=None, GRsetting=None, GRzer=None,
                    pos_type="visual", scale=False, prev_visual_xyz=None, gpf_only=False, gs_only=False, debug=False, **kwargs):
    """pipe_dynamics.py:8-180: the fluid particles ("gpf") are drawn together with the frozen background Gaussians ("gs");
    grey particle colours are repeated to RGB."""
    raw, xyz = _positions(gm, pos_type, scale)
    opacity, scales, rotations, colors = _attributes(gm, pos_type)
    if colors.shape[1] == 1:
        colors = colors.repeat(1, 3)
    if gpf_only:
        means3D = xyz
    elif gs_only:
        means3D, opacity, scales, rotations, colors = gm.get_gs_xyz, gm.get_gs_opacity, gm.get_gs_scaling, gm.get_gs_rotation, gm.get_gs_color
    else:
        means3D = torch.cat([xyz, gm.get_gs_xyz], dim=0)
        opacity = torch.cat([opacity, gm.get_gs_opacity], dim=0)
        scales = torch.cat([scales, gm.get_gs_scaling], dim=0)
        rotations = torch.cat([rotations, gm.get_gs_rotation], dim=0)
        colors = torch.cat([colors, gm.get_gs_color], dim=0)
    return _rasterize(viewpoint_camera, gm, bg_color, scaling_modifier, GRsetting, GRzer, means3D, opacity, scales, rotations, colors, xyz, raw)


def render_background(viewpoint_camera, gm, pipe_args, bg_color, scaling_modifier=1.0, override_color=None, GRsetting=None, GRzer=None,
                      **kwargs):
    """pipe_background.py:9-95: one static set with RGB colours."""
    xyz = gm.get_xyz
    return _rasterize(viewpoint_camera, gm, bg_color, scaling_modifier, GRsetting, GRzer, xyz, gm.get_opacity, gm.get_scaling,
                      gm.get_rotation, gm.get_color, xyz, xyz)
