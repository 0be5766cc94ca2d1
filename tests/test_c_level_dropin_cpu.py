"""CPU: host logic of the `_C`-level drop-in's backward (fluidnexus_b200/rasterizer.py:make_C) -- how it rebuilds the scratch handle
and the argument block from what the reference's wrapper hands back (R3/diff_gaussian_rasterization_ch3/__init__.py:96-126) -- with
the library call replaced by a recorder (no kernels run here; the GPU twin is
tests/test_raster_gpu.py::test_reference_wrapper_package_runs_unchanged_on_the_C_level_dropin)."""
import contextlib
import ctypes as C
import types

import pytest
import torch

from fluidnexus_b200 import _lib as L
from fluidnexus_b200 import rasterizer as R


class _Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*args):
            self.calls.append((name, args))
            return L.FNX_OK
        return fn


@pytest.mark.parametrize("C_,use_sh", [(3, False), (1, False), (3, True)])
def test_backward_rebuilds_scratch_and_args_from_the_buffers(monkeypatch, C_, use_sh):
    rec = _Recorder()
    monkeypatch.setattr(L, "lib", lambda: rec)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: types.SimpleNamespace(cuda_stream=0))
    P, H, W, Rn = 11, 24, 40, 1234
    _, bwd, _ = R.make_C(C_)
    e = torch.Tensor([])
    means, radii = torch.rand(P, 3), torch.ones(P, dtype=torch.int32)
    colors = e if use_sh else torch.rand(P, C_)
    sh = torch.rand(P, 16, 3) if use_sh else e
    scales, rots = torch.rand(P, 3), torch.rand(P, 4)
    view, proj, bg, campos = torch.eye(4), torch.eye(4), torch.zeros(C_), torch.zeros(3)
    geom, binning, img = (torch.zeros(n, dtype=torch.uint8) for n in (1000, 2000, 3000))
    dL = torch.rand(C_, H, W)
    out = bwd(bg, means, radii, colors, scales, rots, 1.5, e, view, proj, 0.4, 0.3, dL, sh, 2 if use_sh else 0, campos, geom, Rn, binning, img)
    # the reference's return order (rasterize_points.cu:193): means2D, colors, opacity, means3D, cov3D, sh, scales, rotations
    assert [tuple(t.shape) for t in out] == [(P, 3), (P, C_), (P, 1), (P, 3), (P, 6), (P, 16, 3) if use_sh else (P, 0, 3), (P, 3), (P, 4)]
    (name, args), = rec.calls
    assert name == f"fnx_raster_backward_ch{C_}"
    a, sc, nr, radii_ptr, dpix_ptr, gr, stream = args
    a, sc, gr = a._obj, sc._obj, gr._obj
    assert (a.P, a.V, a.C, a.W, a.H) == (P, 1, C_, W, H) and nr == Rn and radii_ptr == radii.data_ptr()
    assert a.means3D == means.data_ptr() and a.scales == scales.data_ptr() and a.rotations == rots.data_ptr() and a.cov3D_precomp is None
    assert a.opacities is None                # the reference's backward is not handed them either (include/fnx.h: may be NULL here)
    assert abs(a.tan_fov_x - 0.4) < 1e-7 and abs(a.tan_fov_y - 0.3) < 1e-7 and a.scale_modifier == 1.5 and a.flags == 0
    assert (a.colors is None) == use_sh and (a.sh is None) != use_sh
    if use_sh:
        assert (a.sh_degree, a.sh_coeffs) == (2, 16) and a.campos is not None and gr.dL_dsh == out[5].data_ptr() and gr.dL_dcolors is None
    else:
        assert gr.dL_dcolors == out[1].data_ptr() and gr.dL_dsh is None
    assert (sc.geom, sc.geom_bytes, sc.binning, sc.binning_bytes, sc.image, sc.image_bytes) == (
        geom.data_ptr(), 1000, binning.data_ptr(), 2000, img.data_ptr(), 3000)
    assert sc.binning_capacity == Rn          # the forward sized the binning buffer exactly: capacity == instance count
    assert (gr.dL_dmeans2D, gr.dL_dopacity, gr.dL_dmeans3D, gr.dL_dcov3D, gr.dL_dscales, gr.dL_drotations) == tuple(
        out[i].data_ptr() for i in (0, 2, 3, 4, 6, 7))
    assert dpix_ptr == dL.data_ptr() and stream == 0


def test_backward_with_precomputed_covariance_returns_zero_scale_and_rotation_gradients(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(L, "lib", lambda: rec)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: types.SimpleNamespace(cuda_stream=0))
    P = 5
    _, bwd, _ = R.make_C(3)
    e = torch.Tensor([])
    buf = lambda: torch.zeros(64, dtype=torch.uint8)
    out = bwd(torch.zeros(3), torch.rand(P, 3), torch.ones(P, dtype=torch.int32), torch.rand(P, 3), e, e, 1.0, torch.rand(P, 6), torch.eye(4),
              torch.eye(4), 0.5, 0.5, torch.rand(3, 16, 16), e, 0, torch.zeros(3), buf(), 7, buf(), buf())
    (_, args), = rec.calls
    a, gr = args[0]._obj, args[5]._obj
    assert a.scales is None and a.rotations is None and a.cov3D_precomp is not None
    assert gr.dL_dscales is None and gr.dL_drotations is None
    assert not out[6].any() and not out[7].any()          # rasterize_points.cu:150-158: zero tensors for the path not taken
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        bwd(torch.zeros(3), torch.rand(P * 3), torch.ones(P, dtype=torch.int32), torch.rand(P, 3), e, e, 1.0, torch.rand(P, 6), torch.eye(4),
            torch.eye(4), 0.5, 0.5, torch.rand(3, 16, 16), e, 0, torch.zeros(3), buf(), 7, buf(), buf())
