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
                H, W, campos=None, sh=None, degree=0):
        """colors [P,C] or, with sh [P,M,3] + degree + campos, colors=None (the reference's SH colour path)."""
        campos = torch.zeros(3, device=means3D.device) if campos is None else campos
        colors = _e() if sh is not None else colors
        self._sh, self._degree = (_e() if sh is None else sh), int(degree)
        args = (bg, means3D, colors, opacities, scales, rotations, float(scale_modifier), _e(), view, proj, float(tan_fov_x),
                float(tan_fov_y), int(H), int(W), self._sh, self._degree, campos, False)
        num_rendered, color, radii, geom, binning, img, depth = self.mod.rasterize_gaussians(*args)
        self.saved = dict(bg=bg, means3D=means3D, radii=radii, colors=colors, scales=scales, rotations=rotations,
                          scale_modifier=float(scale_modifier), view=view, proj=proj, tfx=float(tan_fov_x),
                          tfy=float(tan_fov_y), campos=campos, geom=geom, R=num_rendered, binning=binning, img=img)
        return dict(color=color, radii=radii, depth=depth, num_rendered=num_rendered, geom=geom, binning=binning, img=img)

    def backward(self, dL_dcolor):
        s = self.saved
        args = (s["bg"], s["means3D"], s["radii"], s["colors"], s["scales"], s["rotations"], s["scale_modifier"], _e(),
                s["view"], s["proj"], s["tfx"], s["tfy"], dL_dcolor, self._sh, self._degree, s["campos"], s["geom"], s["R"], s["binning"],
                s["img"])
        m2, col, op, m3, cov, sh, sc, rot = self.mod.rasterize_gaussians_backward(*args)
        return dict(means2D=m2, colors=col, opacity=op, means3D=m3, cov3D=cov, scales=sc, rotations=rot, sh=sh)


def carve_geom(buf, P):
    """The reference's own per-Gaussian state out of its opaque geomBuffer.  GeometryState::fromChunk layout
    (R3/cuda_rasterizer/rasterizer_impl.cu:144-160): depths f32[P], clamped bool[3P], radii i32[P], means2D f32[2P], cov3D
    f32[6P], conic_opacity f32[4P], ... each aligned to 128 B from the chunk's own address.  Returns numpy
    (depth [P], means2D [P,2], conic_opacity [P,4])."""
    import numpy as np
    base = buf.data_ptr()
    off = 0

    def take(nbytes):
        nonlocal off
        start = ((base + off + 127) // 128) * 128 - base
        off = start + nbytes
        return start

    o_depth = take(4 * P)
    take(3 * P)
    take(4 * P)
    o_m2 = take(8 * P)
    take(24 * P)
    o_co = take(16 * P)
    raw = buf.cpu().numpy()
    f = lambda o, n: raw[o:o + 4 * n].view(np.float32).copy()
    return f(o_depth, P), f(o_m2, 2 * P).reshape(P, 2), f(o_co, 4 * P).reshape(P, 4)


def dist_cuda2(points):
    return load("knn").distCUDA2(points)
