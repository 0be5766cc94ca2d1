"""CPU: pins for the rasterizer oracle (oracle/raster_ref.c).

1. closed-form spot checks: a single isotropic Gaussian centred on a pixel gives
   alpha = min(.99, o) there and colour = alpha*c + (1-alpha)*bg; tile counts; median depth; culling.
2. fp32 build vs fp64 twin.
3. finite-difference check of the analytic backward (fp64 twin).
4. golden fixtures produced by the COMPILED REFERENCE on a B200 (tests/golden/ref_*.npz, tools/make_golden.py).
"""
import glob
import math
import os

import numpy as np
import pytest

import scenes
from fluidnexus_b200 import synthetic as S
from oracle.raster_oracle import RasterOracle, mark_visible

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _single(opacity, color, scale=0.01, size=65, C=3):
    # odd image size => the plume centre projects exactly onto pixel centre (32,32)
    cam = S.make_cameras(5, size)[2]
    gs = S.GaussianSet(xyz=S.PLUME_CENTER[None].copy(), scales=np.full((1, 3), scale), rotations=np.array([[1., 0, 0, 0]]),
                       opacity=np.array([[opacity]]), colors=np.asarray(color, float).reshape(1, C))
    return gs, cam


@pytest.mark.parametrize("kind", ["f32", "f64"])
def test_single_gaussian_closed_form(oracle_built, kind):
    gs, cam = _single(0.6, [0.2, 0.5, 0.9])
    bg = np.array([0.1, 0.3, 0.7], np.float32)
    o = RasterOracle(kind)
    out = o.forward(**S.raster_inputs(gs, cam, bg))
    g = o.geom()
    assert np.allclose(g["xy"], [[32.0, 32.0]], atol=1e-4)
    assert abs(g["depth"][0] - 1.0) < 1e-5  # cameras sit at radius 1 from the plume centre
    expect = 0.6 * np.array([0.2, 0.5, 0.9]) + 0.4 * bg
    assert np.allclose(out["color"][:, 32, 32], expect, atol=1e-6)
    # focal = W / (2 tan(fov/2)); sigma_px^2 = (focal*scale/depth)^2 + 0.3 ; radius = ceil(3 sigma)
    focal = 65 / (2 * math.tan(0.69 / 2))
    sigma = math.sqrt((focal * 0.01) ** 2 + 0.3)
    assert out["radii"][0] == math.ceil(3 * sigma)
    # far corner is pure background, depth default 15, centre depth = 1 only if alpha crosses .5
    assert np.allclose(out["color"][:, 0, 0], bg)
    assert out["depth"][0, 0, 0] == 15.0
    assert abs(out["depth"][0, 32, 32] - 1.0) < 1e-5


def test_alpha_clamp_and_median_depth(oracle_built):
    gs, cam = _single(1.0, [1.0], C=1)
    out = RasterOracle("f32").forward(**S.raster_inputs(gs, cam, np.array([0.0], np.float32)))
    assert abs(out["color"][0, 32, 32] - 0.99) < 1e-6  # alpha clamp .99
    gs, cam = _single(0.4, [1.0], C=1)
    out = RasterOracle("f32").forward(**S.raster_inputs(gs, cam, np.array([0.0], np.float32)))
    assert out["depth"][0, 32, 32] == 15.0  # T never crosses .5 -> default depth


def test_low_opacity_never_blends(oracle_built):
    gs, cam = _single(1.0 / 300.0, [1.0], C=1)
    out = RasterOracle("f32").forward(**S.raster_inputs(gs, cam, np.array([0.25], np.float32)))
    assert np.all(out["color"] == 0.25)
    assert out["radii"][0] > 0  # still "visible" for the reference


def test_near_cull_and_mark_visible(oracle_built):
    cam = S.make_cameras(5, 64)[2]
    eye = cam.camera_center.numpy().astype(np.float64)
    fwd = (S.PLUME_CENTER - eye) / np.linalg.norm(S.PLUME_CENTER - eye)
    pts = np.stack([eye + fwd * 0.19, eye + fwd * 0.21, eye - fwd * 1.0, eye + fwd * 1.0])
    vis = mark_visible(pts, cam.world_view_transform.numpy())
    assert vis.tolist() == [False, True, False, True]
    gs = S.GaussianSet(xyz=pts, scales=np.full((4, 3), 0.005), rotations=np.tile([1., 0, 0, 0], (4, 1)),
                       opacity=np.full((4, 1), 0.5), colors=np.full((4, 1), 1.0))
    out = RasterOracle("f32").forward(**S.raster_inputs(gs, cam))
    assert (out["radii"] > 0).tolist() == [False, True, False, True]


@pytest.mark.parametrize("name", list(scenes.SCENES))
def test_f32_matches_f64_twin(oracle_built, name):
    gs, cam, bg, inp = scenes.build(name)
    o32, o64 = RasterOracle("f32"), RasterOracle("f64")
    a, b = o32.forward(**inp), o64.forward(**inp)
    assert np.abs(a["color"] - b["color"]).max() < 2e-5
    assert (a["radii"] != b["radii"]).mean() < 1e-3
    dL = scenes.dL_dpix(name, a["color"].shape)
    ga, gb = o32.backward(dL), o64.backward(dL)
    for k in ga:
        rel = np.linalg.norm(ga[k] - gb[k]) / (np.linalg.norm(gb[k]) + 1e-30)
        assert rel < 2e-4, (k, rel)


def test_backward_finite_differences(oracle_built):
    """fp64 twin: directional central differences of L = sum(dL * image) match the analytic backward.

    The renderer is only piecewise smooth (1/255 alpha cut, 3-sigma tile rectangles, depth order), so the
    Gaussians are separated in depth (no reordering under the perturbation), dL is smooth, and the tolerance
    allows for the small jumps of pixels crossing the alpha cut."""
    rng = np.random.default_rng(0)
    gs = S.random_gaussians(40, 3, seed=7, spread=0.05, log_scale=(-3.6, -3.0))
    gs.opacity[:] = rng.uniform(0.5, 0.95, gs.opacity.shape)
    cam = S.make_cameras(5, 48)[2]
    eye = cam.camera_center.numpy().astype(np.float64)
    fwd = S.PLUME_CENTER - eye
    fwd /= np.linalg.norm(fwd)
    lat = gs.xyz - S.PLUME_CENTER
    lat -= np.outer(lat @ fwd, fwd)
    gs.xyz = eye + np.outer(0.8 + 0.01 * np.arange(40), fwd) + lat
    inp = S.raster_inputs(gs, cam, np.array([0.2, 0.4, 0.6], np.float32))
    o = RasterOracle("f64")
    o.forward(**inp)
    yy, xx = np.mgrid[0:48, 0:48]
    dL = np.stack([np.sin(xx / 9.0 + c) + np.cos(yy / 7.0) for c in range(3)]).astype(np.float32)
    g = o.backward(dL)

    def loss(**over):
        d = dict(inp)
        d.update(over)
        return float((RasterOracle("f64").forward(**d)["color"].astype(np.float64) * dL).sum())

    checks = [("means3D", "means3D", 3e-4, 0.06, 0.03), ("colors", "colors", 1e-2, 1e-4, 1e-6),
              ("opacities", "opacity", 2e-3, 0.03, 0.01), ("scales", "scales", 2e-4, 0.08, 0.03),
              ("rotations", "rotations", 2e-3, 0.08, 0.03)]
    for key, gkey, h, rtol, atol in checks:
        arr = inp[key]
        for _ in range(4):
            d = rng.normal(size=arr.shape).astype(np.float32)
            p, m = (arr + h * d).astype(np.float32), (arr - h * d).astype(np.float32)
            num = loss(**{key: p}) - loss(**{key: m})
            ana = float((g[gkey].reshape(arr.shape) * (p.astype(np.float64) - m.astype(np.float64))).sum())
            assert abs(num - ana) <= rtol * max(abs(num), abs(ana)) + atol, (key, num, ana)


# ------------------------------------------------------------------------------------------------------------
# golden fixtures from the compiled reference
# ------------------------------------------------------------------------------------------------------------
_gold = sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz")))


@pytest.mark.skipif(not _gold, reason="no reference fixtures yet (tools/make_golden.py needs the GPU box)")
@pytest.mark.parametrize("path", _gold, ids=[os.path.basename(p) for p in _gold])
def test_oracle_matches_compiled_reference(oracle_built, path):
    z = np.load(path)
    name = str(z["scene"])
    gs, cam, bg, inp = scenes.build(name)
    o = RasterOracle("f32")
    out = o.forward(**inp)
    assert np.abs(out["color"] - z["color"]).max() < 1e-4
    assert (out["radii"] == z["radii"]).mean() > 0.999
    assert int(out["num_rendered"]) == pytest.approx(int(z["num_rendered"]), rel=2e-3)
    dmis = (np.abs(out["depth"] - z["depth"]) > 1e-5).mean()  # CPU (unfused) depths differ by an ulp
    assert dmis < 2e-3, dmis
    dL = scenes.dL_dpix(name, out["color"].shape)
    g = RasterOracle("f64")
    g.forward(**inp)
    gr = g.backward(dL)
    for k in ("means2D", "colors", "opacity", "means3D", "scales", "rotations"):
        ref = z["g_" + k].astype(np.float64).reshape(gr[k].shape)
        rel = np.linalg.norm(gr[k] - ref) / (np.linalg.norm(ref) + 1e-30)
        assert rel < 1e-3, (k, rel)
