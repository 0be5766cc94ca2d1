"""ctypes front-end to oracle/raster_ref.c.  TEST INFRASTRUCTURE ONLY (see that file's header).

Only tests/, bench.py's cpu_baseline/reference legs and __graft_entry__.smoke() may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build():
    """Compile the C restatement (gcc, seconds)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _lib(kind):
    if kind not in _LIBS:
        path = os.path.join(_HERE, "_build", f"liboracle_{kind}.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.fnx_oracle_forward.restype = C.c_void_p
        lib.fnx_oracle_forward.argtypes = [C.c_int] * 4 + [C.c_void_p] * 5 + [C.c_float] + [C.c_void_p] * 4 + [
            C.c_float, C.c_float] + [C.c_void_p] * 4
        lib.fnx_oracle_backward.restype = None
        lib.fnx_oracle_backward.argtypes = [C.c_void_p] * 10
        lib.fnx_oracle_free.argtypes = [C.c_void_p]
        lib.fnx_oracle_get_geom.argtypes = [C.c_void_p] * 5
        lib.fnx_oracle_get_image_state.argtypes = [C.c_void_p] * 3
        lib.fnx_oracle_mark_visible.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIBS[kind] = lib
    return _LIBS[kind]


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RasterOracle:
    """One forward (+ optional backward) of the restated rasterizer.

    Arguments mirror `_C.rasterize_gaussians` (R3/rasterize_points.h:18-37): view/proj are the
    16 floats of the row-major tensors that hold the *transposed* matrices, exactly as the
    reference's Camera stores them (FD/scene/camera.py:90-108).
    """

    def __init__(self, kind="f32"):
        self.lib = _lib(kind)
        self.h = None

    def forward(self, bg, means3D, colors, opacities, scales, rotations, scale_modifier, view, proj, tan_fov_x,
                tan_fov_y, H, W, cov3D_precomp=None):
        means3D = _f32(means3D).reshape(-1, 3)
        P = means3D.shape[0]
        colors = _f32(colors).reshape(P, -1)
        Cn = colors.shape[1]
        bg = _f32(bg).reshape(-1)
        if bg.shape[0] < Cn:
            raise ValueError("bg too short")
        opac = _f32(opacities).reshape(-1)
        scales = _f32(scales)
        rotations = _f32(rotations)
        cov = _f32(cov3D_precomp)
        view = _f32(view).reshape(16)
        proj = _f32(proj).reshape(16)
        color = np.zeros((Cn, H, W), np.float32)
        depth = np.zeros((1, H, W), np.float32)
        radii = np.zeros((P,), np.int32)
        nr = C.c_int64(0)
        self.free()
        self.h = self.lib.fnx_oracle_forward(P, Cn, W, H, _p(bg), _p(means3D), _p(colors), _p(opac), _p(scales),
                                             C.c_float(scale_modifier), _p(rotations), _p(cov), _p(view), _p(proj),
                                             C.c_float(tan_fov_x), C.c_float(tan_fov_y), _p(color), _p(depth),
                                             _p(radii), C.cast(C.byref(nr), C.c_void_p))
        self.P, self.Cn, self.H, self.W = P, Cn, H, W
        return dict(color=color, depth=depth, radii=radii, num_rendered=int(nr.value))

    def geom(self):
        P = self.P
        xy = np.zeros((P, 2)); depth = np.zeros((P,)); co = np.zeros((P, 4)); tt = np.zeros((P,), np.uint32)
        self.lib.fnx_oracle_get_geom(self.h, _p(xy), _p(depth), _p(co), _p(tt))
        return dict(xy=xy, depth=depth, conic_opacity=co, tiles_touched=tt)

    def image_state(self):
        fT = np.zeros((self.H, self.W)); nc = np.zeros((self.H, self.W), np.uint32)
        self.lib.fnx_oracle_get_image_state(self.h, _p(fT), _p(nc))
        return dict(final_T=fT, n_contrib=nc)

    def backward(self, dL_dpix):
        P, Cn = self.P, self.Cn
        dpix = _f32(dL_dpix).reshape(Cn, self.H, self.W)
        out = dict(means2D=np.zeros((P, 3)), conic=np.zeros((P, 4)), opacity=np.zeros((P, 1)),
                   colors=np.zeros((P, Cn)), means3D=np.zeros((P, 3)), cov3D=np.zeros((P, 6)),
                   scales=np.zeros((P, 3)), rotations=np.zeros((P, 4)))
        self.lib.fnx_oracle_backward(self.h, _p(dpix), _p(out["means2D"]), _p(out["conic"]), _p(out["opacity"]),
                                     _p(out["colors"]), _p(out["means3D"]), _p(out["cov3D"]), _p(out["scales"]),
                                     _p(out["rotations"]))
        return out

    def free(self):
        if self.h:
            self.lib.fnx_oracle_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def mark_visible(means3D, view, kind="f32"):
    means3D = _f32(means3D).reshape(-1, 3)
    out = np.zeros((means3D.shape[0],), np.uint8)
    v = _f32(view).reshape(16)
    _lib(kind).fnx_oracle_mark_visible(means3D.shape[0], _p(means3D), _p(v), _p(out))
    return out.astype(bool)
