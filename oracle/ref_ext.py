"""Loader for the compiled, unmodified reference extensions in oracle/_ref (TEST INFRASTRUCTURE ONLY).

The reference Python packages are thin wrappers around `_C.rasterize_gaussians` /
`_C.rasterize_gaussians_backward` (R3/diff_gaussian_rasterization_ch3/__init__.py:72,126) and
`simple_knn._C.distCUDA2` (KNN/ext.cpp).  We bind the compiled modules directly and restate the few lines of
argument shuffling here, so nothing from the reference's source tree is needed at run time (the GPU box has no
/root/reference).  Needs a CUDA device to *run*; loading works anywhere torch does.
"""
import importlib.machinery
import importlib.util
import os

import torch

from . import build_ref

_MODS = {}


def available(key):
    return os.path.exists(build_ref.so_path(key))


def load(key):
    """key in {'ch3','ch1','knn'} -> the pybind module."""
    if key not in _MODS:
        path = build_ref.so_path(key)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `python oracle/build_ref.py {key}` where /root/reference exists")
        name = build_ref.TARGETS[key]["name"]
        loader = importlib.machinery.ExtensionFileLoader(name, path)
        spec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _MODS[key] = mod
    return _MODS[key]


def _e():
    return torch.Tensor([])


class RefRaster:
    """Drives the reference rasterizer exactly like _RasterizeGaussians.forward/backward do."""

    def __init__(self, channels):
        self.C = channels
        self.mod = load("ch3" if channels == 3 else "ch1")

    def forward(self, bg, means3D, colors, opacities, scales, rotations, scale_modifier, view, proj, tan_fov_x, tan_fov_y,
                H, W, campos=None):
        campos = torch.zeros(3, device=means3D.device) if campos is None else campos
        args = (bg, means3D, colors, opacities, scales, rotations, float(scale_modifier), _e(), view, proj, float(tan_fov_x),
                float(tan_fov_y), int(H), int(W), _e(), 0, campos, False)
        num_rendered, color, radii, geom, binning, img, depth = self.mod.rasterize_gaussians(*args)
        self.saved = dict(bg=bg, means3D=means3D, radii=radii, colors=colors, scales=scales, rotations=rotations,
                          scale_modifier=float(scale_modifier), view=view, proj=proj, tfx=float(tan_fov_x),
                          tfy=float(tan_fov_y), campos=campos, geom=geom, R=num_rendered, binning=binning, img=img)
        return dict(color=color, radii=radii, depth=depth, num_rendered=num_rendered, geom=geom, binning=binning, img=img)

    def backward(self, dL_dcolor):
        s = self.saved
        args = (s["bg"], s["means3D"], s["radii"], s["colors"], s["scales"], s["rotations"], s["scale_modifier"], _e(),
                s["view"], s["proj"], s["tfx"], s["tfy"], dL_dcolor, _e(), 0, s["campos"], s["geom"], s["R"], s["binning"],
                s["img"])
        m2, col, op, m3, cov, sh, sc, rot = self.mod.rasterize_gaussians_backward(*args)
        return dict(means2D=m2, colors=col, opacity=op, means3D=m3, cov3D=cov, scales=sc, rotations=rot)


def dist_cuda2(points):
    return load("knn").distCUDA2(points)
