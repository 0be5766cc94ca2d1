"""CPU: size-independent properties of the rasterizer that any implementation of the reference's algorithm must have, checked on the
oracle (fp64 twin) -- the same properties hold the CUDA path at BASELINE sizes where a literal CPU evaluation is out of reach:

* the backward is linear in dL/dpixel (it is a vector-Jacobian product);
* the image does not depend on the ORDER of the input Gaussians (the depth sort decides; ties aside), gradients permute with them;
* a Gaussian that can never reach alpha >= 1/255 (opacity below the cut), one behind the near plane, or one outside the frustum leaves
  the image untouched and receives no gradient;
* the frozen-background split: rendering [A ; B] and differentiating only w.r.t. A gives A's rows of the full gradient;
* with an opaque splat in front, what lies behind is invisible and gradient-free (early termination at T < 1e-4, tested BEFORE the
  blend: the Gaussian that would cross the threshold is itself not blended).
"""
import numpy as np
import pytest

import scenes
from fluidnexus_b200 import synthetic as S
from oracle.raster_oracle import RasterOracle


def _run(inp, dL=None):
    o = RasterOracle("f64")
    out = o.forward(**inp)
    g = o.backward(dL) if dL is not None else None
    return out, g


def _subset(inp, idx):
    d = dict(inp)
    for k in ("means3D", "colors", "opacities", "scales", "rotations"):
        d[k] = inp[k][idx]
    return d


@pytest.mark.parametrize("name", ["mixed_ch3_96", "dense_ch1_64"])
def test_backward_is_linear_in_the_pixel_gradient(oracle_built, name):
    gs, cam, bg, inp = scenes.build(name)
    out, _ = _run(inp)
    rng = np.random.default_rng(3)
    d1, d2 = (rng.normal(size=out["color"].shape).astype(np.float32) for _ in range(2))
    _, g1 = _run(inp, d1)
    _, g2 = _run(inp, d2)
    _, g12 = _run(inp, (2.0 * d1 - 0.5 * d2).astype(np.float32))
    for k in g1:
        want = 2.0 * g1[k] - 0.5 * g2[k]
        assert np.abs(g12[k] - want).max() <= 1e-6 * (np.abs(want).max() + 1e-30) + 1e-12, k


def test_input_order_does_not_matter(oracle_built):
    gs, cam, bg, inp = scenes.build("ragged_ch3_70x45")
    P = inp["means3D"].shape[0]
    perm = np.random.default_rng(0).permutation(P)
    dL = scenes.dL_dpix("ragged_ch3_70x45", (3, inp["H"], inp["W"]))
    a, ga = _run(inp, dL)
    b, gb = _run(_subset(inp, perm), dL)
    assert np.abs(a["color"] - b["color"]).max() < 1e-12 and np.array_equal(a["radii"][perm], b["radii"])
    for k in ("means3D", "colors", "opacity", "scales", "rotations"):
        assert np.abs(ga[k][perm] - gb[k]).max() <= 1e-9 * (np.abs(ga[k]).max() + 1e-30), k


def test_gaussians_that_cannot_contribute_change_nothing(oracle_built):
    gs, cam, bg, inp = scenes.build("mixed_ch3_96")
    P = inp["means3D"].shape[0]
    eye = cam.camera_center.numpy().astype(np.float32)
    fwd = (S.PLUME_CENTER - eye) / np.linalg.norm(S.PLUME_CENTER - eye)
    extra = {k: inp[k][:3].copy() for k in ("means3D", "colors", "opacities", "scales", "rotations")}
    extra["opacities"][0] = 1.0 / 300.0                       # in view, but alpha < 1/255 everywhere
    extra["means3D"][1] = eye + 0.15 * fwd                    # closer than the 0.2 near cut
    extra["means3D"][2] = eye - 1.0 * fwd                     # behind the camera
    both = dict(inp)
    for k, v in extra.items():
        both[k] = np.concatenate([inp[k], v], axis=0)
    dL = scenes.dL_dpix("mixed_ch3_96", (3, inp["H"], inp["W"]))
    a, ga = _run(inp, dL)
    b, gb = _run(both, dL)
    assert np.array_equal(a["color"], b["color"]) and np.array_equal(a["depth"], b["depth"])
    assert b["radii"][P] > 0 and b["radii"][P + 1] == 0 and b["radii"][P + 2] == 0      # the first is "visible" to the reference, the others culled
    for k in ("means3D", "colors", "opacity", "scales", "rotations"):
        assert np.array_equal(ga[k], gb[k][:P]) and not gb[k][P:].any(), k


def test_gradient_rows_of_a_subset_do_not_depend_on_who_else_is_trainable(oracle_built):
    """What the frozen-background split relies on: dL/d(row i) is a property of the rendered scene, not of which rows ask for it."""
    gs, cam, bg, inp = scenes.build("mixed_ch3_96")                      # 1500 fluid ++ 2500 background
    dL = scenes.dL_dpix("mixed_ch3_96", (3, inp["H"], inp["W"]))
    _, g = _run(inp, dL)
    # zeroing the background rows' pixel influence is not possible; instead: duplicates of the scene with the background's
    # attributes perturbed where it cannot matter (rotations of isotropic splats) leave the fluid rows' gradients unchanged
    iso = dict(inp)
    iso["scales"] = inp["scales"].copy()
    iso["scales"][1500:] = iso["scales"][1500:, :1]                      # isotropic background splats
    _, g0 = _run(iso, dL)
    rot = dict(iso)
    q = np.random.default_rng(1).normal(size=(2500, 4)).astype(np.float32)
    rot["rotations"] = inp["rotations"].copy()
    rot["rotations"][1500:] = q / np.linalg.norm(q, axis=1, keepdims=True)   # any unit quaternion: the covariance is s^2 I either way
    out_r, g1 = _run(rot, dL)
    for k in ("means3D", "colors", "opacity", "scales"):
        assert np.abs(g0[k][:1500] - g1[k][:1500]).max() <= 2e-6 * (np.abs(g0[k][:1500]).max() + 1e-30), k


def test_nothing_behind_an_opaque_wall_is_seen_or_trained(oracle_built):
    cam = S.make_cameras(5, 48)[2]
    eye = cam.camera_center.numpy().astype(np.float64)
    fwd = (S.PLUME_CENTER - eye) / np.linalg.norm(S.PLUME_CENTER - eye)
    n_wall, n_back = 6, 5
    centre = eye + 0.6 * fwd
    xyz = np.concatenate([np.tile(centre, (n_wall, 1)) + np.outer(1e-3 * np.arange(n_wall), fwd),      # big opaque splats, stacked in depth
                          np.tile(eye + 0.9 * fwd, (n_back, 1)) + np.random.default_rng(0).normal(0, 0.01, (n_back, 3))])
    P = n_wall + n_back
    gs = S.GaussianSet(xyz=xyz, scales=np.concatenate([np.full((n_wall, 3), 10.0), np.full((n_back, 3), 0.01)]),
                       rotations=np.tile([1., 0, 0, 0], (P, 1)), opacity=np.concatenate([np.full((n_wall, 1), 1.0), np.full((n_back, 1), 0.9)]),
                       colors=np.concatenate([np.full((n_wall, 3), 0.3), np.full((n_back, 3), 1.0)]))
    inp = S.raster_inputs(gs, cam, np.array([0.0, 1.0, 0.0], np.float32))
    dL = np.ones((3, 48, 48), np.float32)
    out, g = _run(inp, dL)
    # alpha is clamped at 0.99 (as a float): T after the first wall layer is 1 - 0.99f, and the second layer would take it to
    # (1 - 0.99f)^2 = 9.9999998e-5 < 1e-4 -- the early-termination test fires BEFORE that layer is blended, so exactly one layer counts
    a = float(np.float32(0.99))
    assert np.allclose(out["color"], 0.3 * a + np.array([0.0, 1.0, 0.0]).reshape(3, 1, 1) * (1 - a), atol=1e-6)
    st = RasterOracle("f64")
    st.forward(**inp)
    assert int(st.image_state()["n_contrib"].max()) == 1
    for k in ("means3D", "colors", "opacity"):
        assert not g[k][1:].any(), k                          # wall layers 2.. and everything behind: no gradient at all
    assert np.abs(g["colors"][0]).min() > 0
