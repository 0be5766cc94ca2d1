"""CPU: the C ABI as a C program sees it.

1. include/fnx.h compiles as plain C11 (gcc -fsyntax-only -Wall -Werror -pedantic) and as C++17: no C++-isms, no torch / CUDA
   types in a signature.
2. Every struct that crosses the boundary has the same size and field offsets in ctypes (fluidnexus_b200/_lib.py) as in C -- a
   mismatch would otherwise only show up as garbage arguments on the GPU.
3. Error behaviour of the entry points (SURVEY.md 8(b) "Errors"): bad arguments are rejected with FNX_ERR_INVALID and a message
   from fnx_last_error() BEFORE any CUDA call is made, so this runs without a GPU (no compute calls).  Messages that the reference
   raises itself keep its wording (rasterizer_impl.cu:226-228).
"""
import ctypes as C
import os
import subprocess

import pytest

from conftest import ROOT
from fluidnexus_b200 import _lib as L

HEADER = os.path.join(ROOT, "include", "fnx.h")
STRUCTS = {"fnx_raster_args": L.RasterArgs, "fnx_raster_scratch": L.RasterScratch, "fnx_raster_grads": L.RasterGrads,
           "fnx_gs_state": L.GsState, "fnx_gs_grads": L.GsGrads, "fnx_gs_hparams": L.GsHparams, "fnx_gs_level_two": L.GsLevelTwo}


@pytest.mark.parametrize("lang", ["c", "c++"])
def test_header_is_plain_c(lang, tmp_path):
    src = tmp_path / ("t.c" if lang == "c" else "t.cpp")
    src.write_text('#include "fnx.h"\nint main(void) { fnx_raster_args a; (void)a; return 0; }\n')
    std = "-std=c11" if lang == "c" else "-std=c++17"
    r = subprocess.run(["gcc" if lang == "c" else "g++", std, "-fsyntax-only", "-Wall", "-Werror", "-pedantic", "-I", os.path.dirname(HEADER), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_ctypes_structs_have_the_c_layout(tmp_path):
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fnx.h"', 'int main(void) {']
    for cname, ct in STRUCTS.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines))
    subprocess.run(["gcc", "-std=c11", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in STRUCTS.items():
        assert int(got[cname]) == C.sizeof(ct), f"sizeof({cname}): C {got[cname]} vs ctypes {C.sizeof(ct)}"
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"offsetof({cname}, {fname})"
    # and the header has no field the ctypes mirror lacks: a struct's fields tile it without a hole at the end larger than its alignment
    for cname, ct in STRUCTS.items():
        last, ftype = ct._fields_[-1]
        assert getattr(ct, last).offset + C.sizeof(ftype) + C.alignment(ct) > C.sizeof(ct), cname


def _args(**kw):
    a = L.RasterArgs()
    a.P, a.V, a.C, a.W, a.H = 10, 1, 3, 64, 64
    for p in ("means3D", "colors", "opacities", "scales", "rotations", "view_matrix", "proj_matrix", "bg"):
        setattr(a, p, 0x1000)   # never dereferenced: every call below is rejected by argument validation
    a.tan_fov_x = a.tan_fov_y = 0.5
    a.scale_modifier = 1.0
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def _null_alloc():
    return L.ALLOC_FN(lambda _ctx, _n: None)


def _fwd(lib, fn, a, out=0x1000, radii=0x1000):
    nr, sc, cb = C.c_int64(0), L.RasterScratch(), _null_alloc()
    return getattr(lib, fn)(C.byref(a), cb, None, cb, None, cb, None, out, out, radii, C.byref(nr), C.byref(sc), None)


@pytest.mark.parametrize("fn,kw,needle", [
    ("fnx_raster_forward_ch3", dict(C=1), "needs C == 3"),
    ("fnx_raster_forward_ch1", dict(C=3), "needs C == 1"),
    ("fnx_raster_forward", dict(C=2), "C must be 1 or 3"),
    ("fnx_raster_forward", dict(P=-1), "bad sizes"),
    ("fnx_raster_forward", dict(V=0), "bad sizes"),
    ("fnx_raster_forward", dict(W=0), "bad sizes"),
    ("fnx_raster_forward", dict(means3D=None), "means3D"),
    ("fnx_raster_forward", dict(scales=None), "scales+rotations or cov3D_precomp"),
    ("fnx_raster_forward", dict(view_matrix=None), "view_matrix"),
    # the reference's own message (rasterizer_impl.cu:226-228)
    ("fnx_raster_forward_ch1", dict(C=1, colors=None, sh=0x1000, sh_coeffs=16, campos=0x1000), "For non-RGB, provide precomputed Gaussian colors!"),
    ("fnx_raster_forward", dict(sh=0x1000, sh_coeffs=16, campos=0x1000), "exactly one of sh / colors"),
    ("fnx_raster_forward", dict(colors=None, sh=0x1000, sh_degree=4, sh_coeffs=16, campos=0x1000), "sh_degree"),
    ("fnx_raster_forward", dict(colors=None, sh=0x1000, sh_degree=2, sh_coeffs=4, campos=0x1000), "sh_degree"),
    ("fnx_raster_forward", dict(colors=None, sh=0x1000, sh_degree=1, sh_coeffs=4), "campos"),
    ("fnx_raster_forward", dict(flags=L.FNX_NO_HOST_SYNC), "instance_capacity_hint"),
    ("fnx_raster_forward", dict(P=1 << 30, V=4), "P*V too large"),
])
def test_forward_rejects_bad_arguments_before_touching_cuda(libfnx, fn, kw, needle):
    rc = _fwd(libfnx, fn, _args(**kw))
    msg = libfnx.fnx_last_error().decode()
    assert rc == L.FNX_ERR_INVALID, (rc, msg)
    assert needle in msg, msg


def test_forward_needs_outputs_and_allocators(libfnx):
    assert _fwd(libfnx, "fnx_raster_forward", _args(), out=None) == L.FNX_ERR_INVALID
    assert "out_color" in libfnx.fnx_last_error().decode()
    assert _fwd(libfnx, "fnx_raster_forward", _args(), radii=None) == L.FNX_ERR_INVALID
    assert "radii" in libfnx.fnx_last_error().decode()
    a, nr, sc = _args(), C.c_int64(0), L.RasterScratch()
    null_fn = C.cast(None, L.ALLOC_FN)
    rc = libfnx.fnx_raster_forward(C.byref(a), null_fn, None, null_fn, None, null_fn, None, 0x1000, 0x1000, 0x1000, C.byref(nr), C.byref(sc), None)
    assert rc == L.FNX_ERR_INVALID and "allocators" in libfnx.fnx_last_error().decode()
    assert libfnx.fnx_raster_forward(None, null_fn, None, null_fn, None, null_fn, None, 0x1000, 0x1000, 0x1000, C.byref(nr), C.byref(sc), None) == L.FNX_ERR_INVALID


def test_backward_and_helpers_reject_bad_arguments(libfnx):
    a, sc, gr = _args(), L.RasterScratch(), L.RasterGrads()
    assert libfnx.fnx_raster_backward(C.byref(a), C.byref(sc), 0, 0x1000, 0x1000, C.byref(gr), None) == L.FNX_ERR_INVALID
    assert "scratch buffers missing" in libfnx.fnx_last_error().decode()
    assert libfnx.fnx_raster_backward_ch1(C.byref(a), C.byref(sc), 0, 0x1000, 0x1000, C.byref(gr), None) == L.FNX_ERR_INVALID
    assert "needs C == 1" in libfnx.fnx_last_error().decode()
    # opacities: needed by the forward, not by the backward (the reference's backward has no such argument, rasterize_points.h:39-59)
    assert _fwd(libfnx, "fnx_raster_forward", _args(opacities=None)) == L.FNX_ERR_INVALID and "opacities" in libfnx.fnx_last_error().decode()
    b = _args(opacities=None)
    assert libfnx.fnx_raster_backward(C.byref(b), C.byref(sc), 0, 0x1000, 0x1000, C.byref(gr), None) == L.FNX_ERR_INVALID
    assert "scratch buffers missing" in libfnx.fnx_last_error().decode()       # i.e. the argument block itself was accepted
    p = C.c_void_p()
    assert libfnx.fnx_raster_overflow_flag(C.byref(sc), C.byref(p)) == L.FNX_ERR_INVALID
    nr = C.c_int64(0)
    assert libfnx.fnx_raster_check(C.byref(sc), C.byref(nr), None) == L.FNX_ERR_INVALID
    assert "no forward to check" in libfnx.fnx_last_error().decode()
    assert libfnx.fnx_mark_visible(0, None, None, None, None, None) == L.FNX_OK      # P == 0 short-circuits (rasterize_points.cu:81)
    assert libfnx.fnx_mark_visible(-1, 0x1000, 0x1000, 0x1000, 0x1000, None) == L.FNX_ERR_INVALID
    assert libfnx.fnx_mark_visible(5, None, 0x1000, 0x1000, 0x1000, None) == L.FNX_ERR_INVALID
    # physics / loss / optimiser entry points
    assert libfnx.fnx_grid_build(0x1000, 10, C.c_float(0.0), 0x1000, None) == L.FNX_ERR_INVALID          # cell <= 0
    assert libfnx.fnx_grid_build(None, -1, C.c_float(1.0), 0x1000, None) == L.FNX_ERR_INVALID
    assert libfnx.fnx_radius_count(0x1000, 10, C.c_float(1.0), 0x1000, 10, C.c_float(2.0), 32, 0x1000, 0x1000, None) == L.FNX_ERR_INVALID
    assert "r <= cell" in libfnx.fnx_last_error().decode()
    assert libfnx.fnx_pair_distance_loss(0x1000, 0x1000, 10, C.c_float(1.0), C.c_float(2.0), C.c_float(1.0), 0x1000, 0x1000, None) == L.FNX_ERR_INVALID
    assert "threshold <= cell" in libfnx.fnx_last_error().decode()
    assert libfnx.fnx_image_loss(0, 3, 64, 64, 0x1000, 0x1000, 0, C.c_float(0.2), C.c_float(1.0), 0x1000, 0x1000, 0x1000, 0x1000, None) == L.FNX_ERR_INVALID
    assert libfnx.fnx_adam_step(10, 0x1000, 0x1000, 0x1000, 0x1000, C.c_float(1e-3), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-15),
                                C.c_float(0.0), 0, None) == L.FNX_ERR_INVALID                              # step counts from 1
    assert libfnx.fnx_scatter_min(10, None, None, 4, 0x1000, 0x1000, None) == L.FNX_ERR_INVALID
    hp = L.GsHparams()
    hp.step = 0
    st, gg = L.GsState(), L.GsGrads()
    assert libfnx.fnx_gs_update(10, 3, C.byref(st), C.byref(gg), C.byref(hp), None, None, None) == L.FNX_ERR_INVALID
    assert "step counts from 1" in libfnx.fnx_last_error().decode()
