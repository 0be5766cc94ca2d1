"""GPU: parity of libfnx's rasterizer (through the C ABI) against the oracle and the reference fixtures.

Tolerances (fp32 path): rendered pixels max|d| < 1e-3 (north_star); gradients vs the fp64 oracle
rel-L2 < 2e-4 per tensor (the reference's own atomics make its gradients run-to-run noisy at ~1e-6..1e-5).
"""
import glob
import os

import numpy as np
import pytest
import torch

import scenes
from fluidnexus_b200 import rasterizer as R
from fluidnexus_b200 import synthetic as S
from oracle import ref_ext
from oracle.raster_oracle import RasterOracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PIX_TOL = 1e-3
GRAD_TOL = 2e-4


def _t(x, dev="cuda"):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def run_fnx(inp, exact_rect=False, speculative=False, views=None):
    C = inp["colors"].shape[1]
    view, proj = _t(inp["view"]), _t(inp["proj"])
    if views is not None:
        view, proj = views
    return R.raster_forward(C, _t(inp["bg"]), _t(inp["means3D"]), _t(inp["colors"]), _t(inp["opacities"]),
                            _t(inp["scales"]), _t(inp["rotations"]), inp["scale_modifier"], None, view, proj,
                            inp["tan_fov_x"], inp["tan_fov_y"], inp["H"], inp["W"], exact_rect=exact_rect,
                            speculative=speculative)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


@pytest.mark.parametrize("exact_rect", [False, True])
@pytest.mark.parametrize("name", list(scenes.SCENES))
def test_forward_matches_oracle(libfnx, oracle_built, name, exact_rect):
    gs, cam, bg, inp = scenes.build(name)
    o = RasterOracle("f32")
    ref = o.forward(**inp)
    ctx, color, radii, depth = run_fnx(inp, exact_rect=exact_rect)
    color, radii, depth = color.cpu().numpy(), radii.cpu().numpy(), depth.cpu().numpy()
    assert color.shape == ref["color"].shape and depth.shape == ref["depth"].shape
    assert np.abs(color - ref["color"]).max() < PIX_TOL
    assert (radii == ref["radii"]).mean() > 0.999
    # median depth flips only where T crosses .5 within rounding (CPU depths are unfused: compare with a tolerance)
    assert (np.abs(depth - ref["depth"]) > 1e-5).mean() < 2e-3
    if exact_rect:
        assert abs(ctx.num_rendered - ref["num_rendered"]) <= max(2, 2e-3 * ref["num_rendered"])
        st = R.read_image_state(ctx)
        ist = o.image_state()
        assert (st["n_contrib"][0].cpu().numpy() != ist["n_contrib"]).mean() < 2e-3
        assert np.abs(st["final_T"][0].cpu().numpy() - ist["final_T"]).max() < 1e-4
    else:
        assert ctx.num_rendered <= ref["num_rendered"]
    g = R.read_geom(ctx)
    og = o.geom()
    vis = ref["radii"] > 0
    assert np.abs(g["xy"][0].cpu().numpy()[vis] - og["xy"][vis]).max() < 1e-2
    assert np.abs(g["depth"][0].cpu().numpy()[vis] - og["depth"][vis]).max() < 1e-5


@pytest.mark.parametrize("name", list(scenes.SCENES))
def test_culled_and_exact_rect_are_identical(libfnx, name):
    """The opacity-aware tile culling must not change a single bit of the image or the gradients' inputs."""
    gs, cam, bg, inp = scenes.build(name)
    c1, col1, rad1, d1 = run_fnx(inp, exact_rect=False)
    c2, col2, rad2, d2 = run_fnx(inp, exact_rect=True)
    assert torch.equal(col1, col2) and torch.equal(rad1, rad2) and torch.equal(d1, d2)
    assert torch.equal(R.read_image_state(c1)["final_T"], R.read_image_state(c2)["final_T"])
    assert c1.num_rendered <= c2.num_rendered


@pytest.mark.parametrize("name", list(scenes.SCENES))
def test_backward_matches_fp64_oracle(libfnx, oracle_built, name):
    gs, cam, bg, inp = scenes.build(name)
    o = RasterOracle("f64")
    ref = o.forward(**inp)
    dL = scenes.dL_dpix(name, ref["color"].shape)
    gref = o.backward(dL)
    for exact_rect in (False, True):
        ctx, color, radii, depth = run_fnx(inp, exact_rect=exact_rect)
        g = R.raster_backward(ctx, _t(dL))
        for k in ("means2D", "colors", "opacity", "means3D", "scales", "rotations", "cov3D"):
            r = rel(g[k].cpu().numpy().reshape(gref[k].shape), gref[k])
            assert r < GRAD_TOL, (k, exact_rect, r)
        assert torch.all(g["means2D"][..., 2] == 0)


def test_speculative_capacity_path_matches_exact(libfnx):
    """Second call uses the capacity hint (no blocking size query); a shrunken hint forces the overflow retry."""
    gs, cam, bg, inp = scenes.build("mixed_ch3_96")
    c0, col0, _, _ = run_fnx(inp, speculative=False)
    c1, col1, _, _ = run_fnx(inp, speculative=True)   # first speculative call: no hint yet -> exact
    c2, col2, _, _ = run_fnx(inp, speculative=True)   # uses the hint
    assert torch.equal(col0, col1) and torch.equal(col0, col2)
    assert c2.num_rendered == c0.num_rendered
    for k in list(R._capacity_hint):
        R._capacity_hint[k] = 64  # far too small -> library must grow and redo
    c3, col3, _, _ = run_fnx(inp, speculative=True)
    assert torch.equal(col0, col3) and c3.num_rendered == c0.num_rendered
    dL = _t(scenes.dL_dpix("mixed_ch3_96", tuple(col0.shape)))
    g0, g3 = R.raster_backward(c0, dL), R.raster_backward(c3, dL)
    assert rel(g3["means3D"].cpu().numpy(), g0["means3D"].cpu().numpy()) < 1e-5


def test_batched_views_equal_single_views(libfnx):
    gs = S.cat_sets(S.fluid_gaussians(1500, 3, seed=10), S.background_gaussians(2500, 3, seed=11))
    cams = S.make_cameras(5, 96, height=80)
    inps = [S.raster_inputs(gs, c, np.array([0.1, 0.2, 0.3], np.float32)) for c in cams]
    views = torch.stack([_t(i["view"]) for i in inps])
    projs = torch.stack([_t(i["proj"]) for i in inps])
    ctxb, colb, radb, depb = run_fnx(inps[0], views=(views, projs))
    assert colb.shape == (5, 3, 80, 96)
    dL = torch.randn(colb.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    gb = R.raster_backward(ctxb, dL)
    acc = None
    for v, inp in enumerate(inps):
        ctx, col, rad, dep = run_fnx(inp)
        assert torch.equal(col, colb[v]) and torch.equal(rad, radb[v]) and torch.equal(dep, depb[v])
        g = R.raster_backward(ctx, dL[v])
        assert rel(gb["means2D"][v].cpu().numpy(), g["means2D"].cpu().numpy()) < 1e-5
        acc = {k: g[k].clone() if acc is None else acc[k] + g[k] for k in g if k != "means2D"}
    for k in acc:
        assert rel(gb[k].cpu().numpy(), acc[k].cpu().numpy()) < 1e-5, k


def test_edge_cases(libfnx):
    gs, cam, bg, inp = scenes.build("behind_ch3_48")
    # P == 0 -> zero outputs (rasterize_points.cu:81)
    e = dict(inp)
    for k in ("means3D", "colors", "opacities", "scales", "rotations"):
        e[k] = inp[k][:0]
    ctx, col, rad, dep = run_fnx(e)
    assert col.abs().max() == 0 and dep.abs().max() == 0 and rad.numel() == 0 and ctx.num_rendered == 0
    # everything culled -> background + depth 15
    f = dict(inp)
    f["means3D"] = inp["means3D"] * 0 + np.array([0.34, 0.3, 50.0], np.float32)  # far behind all cameras
    ctx, col, rad, dep = run_fnx(f)
    assert ctx.num_rendered == 0 and torch.all(rad == 0) and torch.all(dep == 15.0)
    assert torch.allclose(col, _t(inp["bg"]).view(3, 1, 1).expand_as(col))
    g = R.raster_backward(ctx, torch.ones_like(col))
    assert all(torch.all(v == 0) for v in g.values())
    # shape error -> RuntimeError like AT_ERROR (rasterize_points.cu:56-58)
    with pytest.raises(RuntimeError):
        R.raster_forward(3, _t(inp["bg"]), _t(inp["means3D"]).reshape(-1), _t(inp["colors"]), _t(inp["opacities"]),
                         _t(inp["scales"]), _t(inp["rotations"]), 1.0, None, _t(inp["view"]), _t(inp["proj"]),
                         inp["tan_fov_x"], inp["tan_fov_y"], inp["H"], inp["W"])


def test_autograd_module_interface(libfnx):
    """The reference's class interface: settings NamedTuple, nn.Module call, 3 outputs, grads on all inputs."""
    from fluidnexus_b200 import install_compat
    install_compat()
    from diff_gaussian_rasterization_ch3 import GaussianRasterizationSettings, GaussianRasterizer
    gs, cam, bg, inp = scenes.build("mixed_ch3_96")
    t = {k: _t(inp[k]).requires_grad_(True) for k in ("means3D", "colors", "opacities", "scales", "rotations")}
    rs = GaussianRasterizationSettings(
        image_height=inp["H"], image_width=inp["W"], tan_fov_x=inp["tan_fov_x"], tan_fov_y=inp["tan_fov_y"],
        bg=_t(inp["bg"]), scale_modifier=1.0, view_matrix=_t(inp["view"]), proj_matrix=_t(inp["proj"]), sh_degree=0,
        campos=torch.zeros(3, device="cuda"), prefiltered=False)
    rz = GaussianRasterizer(raster_settings=rs)
    means2D = torch.zeros_like(t["means3D"], requires_grad=True) + 0
    means2D.retain_grad()
    color, radii, depth = rz(means3D=t["means3D"], means2D=means2D, shs=None, colors_precomp=t["colors"],
                             opacities=t["opacities"], scales=t["scales"], rotations=t["rotations"], cov3D_precomp=None)
    assert color.shape == (3, inp["H"], inp["W"]) and radii.dtype == torch.int32 and depth.shape == (1, inp["H"], inp["W"])
    (color * _t(scenes.dL_dpix("mixed_ch3_96", tuple(color.shape)))).sum().backward()
    for k, v in t.items():
        assert v.grad is not None and v.grad.shape == v.shape and torch.isfinite(v.grad).all(), k
    assert means2D.grad is not None and means2D.grad.abs().sum() > 0
    vis = rz.mark_visible(t["means3D"].detach())
    assert vis.dtype == torch.bool and vis.shape == (t["means3D"].shape[0],)
    with pytest.raises(Exception):
        rz(means3D=t["means3D"], means2D=means2D, opacities=t["opacities"])  # neither SH nor colours


_gold = sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz")))


@pytest.mark.skipif(not _gold, reason="no reference fixtures yet")
@pytest.mark.parametrize("path", _gold, ids=[os.path.basename(p) for p in _gold])
def test_matches_compiled_reference_fixture(libfnx, path):
    z = np.load(path)
    name = str(z["scene"])
    gs, cam, bg, inp = scenes.build(name)
    ctx, color, radii, depth = run_fnx(inp)
    assert np.abs(color.cpu().numpy() - z["color"]).max() < PIX_TOL
    assert np.array_equal(radii.cpu().numpy(), z["radii"])
    assert (depth.cpu().numpy() != z["depth"]).mean() < 1e-3
    dL = scenes.dL_dpix(name, tuple(color.shape))
    g = R.raster_backward(ctx, _t(dL))
    for k in ("means2D", "colors", "opacity", "means3D", "scales", "rotations"):
        r = rel(g[k].cpu().numpy().reshape(z["g_" + k].shape), z["g_" + k])
        assert r < 1e-3, (k, r)


def test_frozen_range_gives_identical_gradients_for_the_trainable_rows(libfnx):
    """grad_begin/grad_end: frozen Gaussians still occlude; the trainable rows get exactly the gradients of a full
    backward (to the rounding of float atomics), the frozen rows read zero."""
    gs, cam, bg, inp = scenes.build("mixed_ch3_96")
    P = inp["means3D"].shape[0]
    nb = 1500  # the first 1500 (fluid) trainable, the background frozen
    args = (3, _t(inp["bg"]), _t(inp["means3D"]), _t(inp["colors"]), _t(inp["opacities"]), _t(inp["scales"]), _t(inp["rotations"]),
            1.0, None, _t(inp["view"]), _t(inp["proj"]), inp["tan_fov_x"], inp["tan_fov_y"], inp["H"], inp["W"])
    c_all, col_all, _, _ = R.raster_forward(*args, speculative=False)
    c_frz, col_frz, _, _ = R.raster_forward(*args, speculative=False, grad_range=(0, nb))
    assert torch.equal(col_all, col_frz)
    dL = _t(scenes.dL_dpix("mixed_ch3_96", tuple(col_all.shape)))
    g_all, g_frz = R.raster_backward(c_all, dL), R.raster_backward(c_frz, dL)
    for k in ("means3D", "means2D", "colors", "opacity", "scales", "rotations", "cov3D"):
        a, f = g_all[k].reshape(P, -1), g_frz[k].reshape(P, -1)
        assert rel(f[:nb].cpu().numpy(), a[:nb].cpu().numpy()) < 1e-5, k
        assert torch.all(f[nb:] == 0), k


@pytest.mark.parametrize("prepare,tile_cache,bucket,dense", [(False, False, False, False), (True, False, False, False),
                                                             (True, True, False, False), (False, False, True, False),
                                                             (True, True, True, False), (True, True, True, True)])
def test_merged_static_dynamic_streams_equal_concatenated_set(libfnx, prepare, tile_cache, bucket, dense):
    """MergedRasterWorkspace (static stream binned once, truncated to the depth the static-only blend reaches, tiles
    without dynamic instances kept from the static-only render, backward resumed from the forward's snapshot at the
    last dynamic record) against one plain forward/backward over [dynamic ; static] -- over several iterations in
    which the dynamic set moves across tiles, leaves the image and comes back.  bucket: per-tile buckets sorted in shared
    memory instead of two global radix sorts; dense: > 2048 dynamic instances in one tile (the rank-sort path)."""
    dev = torch.device("cuda")
    V, size, nd = (3, 128, 900) if not dense else (2, 64, 9000)
    cams = S.make_cameras(5, size, device=dev)[:V]
    view = torch.stack([c.world_view_transform.float() for c in cams]).contiguous()
    proj = torch.stack([c.full_proj_transform.float() for c in cams]).contiguous()
    import math
    tfx, tfy = math.tan(cams[0].FoVx * 0.5), math.tan(cams[0].FoVy * 0.5)
    dyn_np, sta_np = S.fluid_gaussians(nd, 3, seed=60, log_scale=-4.8), S.background_gaussians(5000, 3, seed=61)
    if dense:  # squeeze the plume into a few tiles and make it nearly transparent so that nothing saturates early
        c = dyn_np.xyz.mean(0, keepdims=True)
        dyn_np.xyz = (c + (dyn_np.xyz - c) * 0.25).astype(np.float32)
        dyn_np.opacity = (dyn_np.opacity * 0.1).astype(np.float32)
    d, s = dyn_np.torch(dev), sta_np.torch(dev)
    key = dict(means3D="xyz", colors="colors", opacities="opacity", scales="scales", rotations="rotations")
    dyn = {k: d[v].reshape(-1).contiguous() if k == "opacities" else d[v].contiguous() for k, v in key.items()}
    sta = {k: s[v].reshape(-1).contiguous() if k == "opacities" else s[v].contiguous() for k, v in key.items()}
    bg = torch.tensor([0.05, 0.1, 0.2], device=dev)
    ws = R.MergedRasterWorkspace(dev, nd, V, size, size, bg, dyn, sta, view, proj, tfx, tfy, margin=3.0, static_prepare=prepare,
                                 static_tile_cache=tile_cache, bucket_binning=bucket)
    base = dyn["means3D"].clone()
    gen = torch.Generator(device="cpu").manual_seed(5)
    offsets = [(0.0, 0.0, 0.0), (0.08, 0.0, 0.0), (-0.1, 0.05, 0.02), (5.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.02, -0.04, 0.0)]
    seen_static_only = 0
    for it, off in enumerate(offsets):
        dyn["means3D"].copy_(base + torch.tensor(off, device=dev))
        img = ws.forward(dyn["means3D"], dyn["colors"], dyn["opacities"], dyn["scales"], dyn["rotations"])
        dL = (torch.randn(V, 3, size, size, generator=gen)).to(dev)
        g = ws.backward(dL)["means3D"].clone()
        cat = lambda k: torch.cat([dyn[k], sta[k]], 0).contiguous()
        ctx, ref_img, _, ref_depth = R.raster_forward(3, bg, cat("means3D"), cat("colors"), cat("opacities"), cat("scales"),
                                                      cat("rotations"), 1.0, None, view, proj, tfx, tfy, size, size, speculative=False)
        assert torch.equal(img, ref_img), f"iteration {it}: image differs, max|d| = {(img - ref_img).abs().max().item()}"
        assert torch.equal(ws.depth, ref_depth), f"iteration {it}: depth differs"
        gref = R.raster_backward(ctx, dL)["means3D"][:nd]
        r = rel(g.cpu().numpy(), gref.cpu().numpy()) if float(gref.abs().max()) > 0 else float(g.abs().max())
        assert r < 2e-5, (it, r)
        ts = ws.tile_state()
        seen_static_only += int((ts["tile_src"] != 0).sum())
        assert np.all(ts["tile_dyn_last"][ts["tile_src"] != 0] == 0)
        assert np.all(ts["tile_dyn_last"][ts["tile_src"] == 0] >= 1)
    assert seen_static_only > 0
    if dense:
        assert int((ts["end"] - ts["begin"]).max()) > 2048


def test_bucket_binning_equals_sorted_binning(libfnx, monkeypatch):
    """FNX_BUCKET_BINNING (per-tile buckets sorted in shared memory) against the two-radix-sort pipeline: identical images,
    depth, radii, instance counts; gradients to the rounding of float atomics.  `dense` scenes put > 8192 instances into
    one tile (the rank-sort path)."""
    for name, dense in (("mixed_ch3_96", False), ("fluid_ch1_112", False), ("ragged_ch3_70x45", False), ("dense_ch1_64", True)):
        gs, cam, bg, inp = scenes.build(name)
        if dense:  # 12x the Gaussians squeezed into the same few tiles
            rng = np.random.default_rng(1)
            for k in ("means3D", "colors", "opacities", "scales", "rotations"):
                inp[k] = np.concatenate([inp[k]] * 12, 0)
            inp["means3D"] = (inp["means3D"] + rng.normal(0, 0.002, inp["means3D"].shape)).astype(np.float32)
            inp["opacities"] = (inp["opacities"] * 0.05).astype(np.float32)
        res = {}
        for flag in (False, True):
            monkeypatch.setattr(R, "BUCKET_BINNING", flag)
            ctx, color, radii, depth = run_fnx(inp)
            g = R.raster_backward(ctx, _t(scenes.dL_dpix(name, tuple(color.shape))))
            res[flag] = (color, radii, depth, ctx.num_rendered, g)
        a, b = res[False], res[True]
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and a[3] == b[3], name
        for k in ("means3D", "means2D", "colors", "opacity", "scales", "rotations"):
            assert rel(b[4][k].cpu().numpy(), a[4][k].cpu().numpy()) < 2e-5, (name, k)


def test_tile_order_is_a_work_sorted_permutation_and_changes_nothing(libfnx):
    """fnx_raster_tile_order: a permutation of the (view, tile) units, longest tile first (8-record bins); forward images and
    backward gradients with the order are those without it (bit-identical images; gradients to float-atomic noise)."""
    gs = S.cat_sets(S.fluid_gaussians(3000, 3, seed=10), S.background_gaussians(3000, 3, seed=11))
    cams = S.make_cameras(5, 128)
    dev = torch.device("cuda")
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)
    inp = S.raster_inputs(gs, cams[0], np.array([0.1, 0.2, 0.3], np.float32))
    view = torch.stack([c.world_view_transform for c in cams]).to(dev).contiguous()
    proj = torch.stack([c.full_proj_transform for c in cams]).to(dev).contiguous()
    ws = R.RasterWorkspace(dev, 3, gs.P, 5, 128, 128, 400_000, want=("means3D", "colors", "opacity"))
    args = (t(inp["bg"]), t(inp["means3D"]), t(inp["colors"]), t(inp["opacities"]).reshape(-1).contiguous(), t(inp["scales"]), t(inp["rotations"]), 1.0,
            view, proj, inp["tan_fov_x"], inp["tan_fov_y"])
    img0 = ws.forward(*args).clone()
    dL = torch.randn(img0.shape, device=dev, generator=torch.Generator("cuda").manual_seed(3))
    g0 = {k: v.clone() for k, v in ws.backward(dL).items()}
    ws.update_tile_order()
    torch.cuda.synchronize()
    order = ws.tile_order.cpu().numpy()
    n = 5 * 8 * 8
    assert sorted(order.tolist()) == list(range(n))
    st = torch.zeros((n, 4), dtype=torch.int32, device=dev)
    import ctypes as C
    from fluidnexus_b200 import _lib as L
    L.check(L.lib().fnx_raster_read_tiles(C.byref(ws.scratch), 128, 128, 5, 0, None, st.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream))
    work = st.max(dim=1).values.cpu().numpy()[order]
    bins = (work + 7) // 8
    assert (np.diff(bins) <= 0).all() and work[0] == work.max() and work.max() > 8 * work[work > 0].min()
    img1 = ws.forward(*args)
    assert torch.equal(img1, img0)
    g1 = ws.backward(dL)
    for k in g0:
        assert float((g1[k] - g0[k]).norm() / g0[k].norm()) < 1e-5, k


@pytest.mark.parametrize("subset", [[1, 3], [4, 0, 2], [2]])
def test_one_static_stream_serves_every_camera_subset(libfnx, subset):
    """the static stream of a frame is built ONCE for all cameras (StaticStream); MergedRasterWorkspaces for subsets of
    the cameras (in any order -- the reference draws random.sample(view_set, batch)) share it through static_view_map and give
    exactly the image / depth / gradient of one plain forward over [dynamic ; static] with those cameras."""
    import math
    dev = torch.device("cuda")
    size, nd = 128, 900
    cams = S.make_cameras(5, size, device=dev)
    view_all = torch.stack([c.world_view_transform.float() for c in cams]).contiguous()
    proj_all = torch.stack([c.full_proj_transform.float() for c in cams]).contiguous()
    tfx, tfy = math.tan(cams[0].FoVx * 0.5), math.tan(cams[0].FoVy * 0.5)
    d, s = S.fluid_gaussians(nd, 3, seed=60, log_scale=-4.8).torch(dev), S.background_gaussians(5000, 3, seed=61).torch(dev)
    key = dict(means3D="xyz", colors="colors", opacities="opacity", scales="scales", rotations="rotations")
    dyn = {k: d[v].reshape(-1).contiguous() if k == "opacities" else d[v].contiguous() for k, v in key.items()}
    sta = {k: s[v].reshape(-1).contiguous() if k == "opacities" else s[v].contiguous() for k, v in key.items()}
    bg = torch.tensor([0.05, 0.1, 0.2], device=dev)
    stream = R.StaticStream(dev, 5, size, size, bg, sta, view_all, proj_all, tfx, tfy)
    idx = torch.tensor(subset, device=dev)
    view, proj = view_all[idx].contiguous(), proj_all[idx].contiguous()
    ws = R.MergedRasterWorkspace(dev, nd, len(subset), size, size, bg, dyn, None, view, proj, tfx, tfy, margin=3.0, static_stream=stream,
                                 view_ids=subset)
    base = dyn["means3D"].clone()
    gen = torch.Generator(device="cpu").manual_seed(5)
    for it, off in enumerate([(0.0, 0.0, 0.0), (0.08, 0.0, 0.0), (5.0, 0.0, 0.0), (0.02, -0.04, 0.0)]):
        dyn["means3D"].copy_(base + torch.tensor(off, device=dev))
        img = ws.forward(dyn["means3D"], dyn["colors"], dyn["opacities"], dyn["scales"], dyn["rotations"])
        dL = torch.randn(len(subset), 3, size, size, generator=gen).to(dev)
        g = ws.backward(dL)["means3D"].clone()
        cat = lambda k: torch.cat([dyn[k], sta[k]], 0).contiguous()
        ctx, ref_img, _, ref_depth = R.raster_forward(3, bg, cat("means3D"), cat("colors"), cat("opacities"), cat("scales"), cat("rotations"), 1.0,
                                                      None, view, proj, tfx, tfy, size, size, speculative=False)
        assert torch.equal(img, ref_img), (it, float((img - ref_img).abs().max()))
        assert torch.equal(ws.depth, ref_depth), it
        gref = R.raster_backward(ctx, dL)["means3D"][:nd]
        r = rel(g.cpu().numpy(), gref.cpu().numpy()) if float(gref.abs().max()) > 0 else float(g.abs().max())
        assert r < 2e-5, (it, r)
    assert (ws.tile_state()["tile_src"] != 0).sum() > 0


@pytest.mark.skipif(not ref_ext.available("ch3"), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("degree,M", [(0, 1), (1, 4), (2, 16), (3, 16)])
def test_sh_colour_path_matches_compiled_reference(libfnx, degree, M):
    """The reference module's other colour input (shs + sh_degree + campos instead of colors_precomp; dead on FluidNexus' pipes,
    R3/cuda_rasterizer/forward.cu:20-67, backward.cu:20-132) through the drop-in GaussianRasterizer: image / depth / radii and the
    gradients w.r.t. means3D, shs, opacity, scales, rotations against the compiled reference run live, incl. clamped channels."""
    import math
    from fluidnexus_b200 import rasterizer as RZ
    dev = torch.device("cuda")
    gs = S.random_gaussians(2500, 3, seed=70, spread=0.2, log_scale=(-4.6, -3.2))
    cam = S.make_cameras(5, 112, device=dev)[1]
    rng = np.random.default_rng(degree)
    sh = rng.normal(0, 0.6, (gs.P, M, 3)).astype(np.float32)          # strong enough to clamp some channels at zero
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)
    means, op, sc, rot, shs = t(gs.xyz).requires_grad_(True), t(gs.opacity).requires_grad_(True), t(gs.scales).requires_grad_(True), \
        t(gs.rotations).requires_grad_(True), t(sh).requires_grad_(True)
    Settings, Rasterizer, _, _ = RZ.make_module(3)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    rs = Settings(image_height=112, image_width=112, tan_fov_x=math.tan(cam.FoVx * 0.5), tan_fov_y=math.tan(cam.FoVy * 0.5), bg=bg,
                  scale_modifier=1.0, view_matrix=cam.world_view_transform, proj_matrix=cam.full_proj_transform, sh_degree=degree,
                  campos=cam.camera_center, prefiltered=False)
    img, radii, depth = Rasterizer(rs)(means3D=means, means2D=torch.zeros_like(means, requires_grad=True), opacities=op, shs=shs,
                                       colors_precomp=None, scales=sc, rotations=rot, cov3D_precomp=None)
    dL = torch.randn(img.shape, device=dev, generator=torch.Generator("cuda").manual_seed(11))
    (img * dL).sum().backward()
    rr = ref_ext.RefRaster(3)
    ro = rr.forward(bg, means.detach(), None, op.detach(), sc.detach(), rot.detach(), 1.0, cam.world_view_transform, cam.full_proj_transform,
                    rs.tan_fov_x, rs.tan_fov_y, 112, 112, campos=cam.camera_center, sh=shs.detach(), degree=degree)
    rg = rr.backward(dL.contiguous())
    assert torch.equal(radii, ro["radii"]) and torch.equal(depth, ro["depth"])
    assert float((img - ro["color"]).abs().max()) < 1e-5
    rel_ = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
    nb = (degree + 1) ** 2
    assert rel_(shs.grad[:, :nb], rg["sh"][:, :nb]) < 1e-4 and float(shs.grad[:, nb:].abs().max() if nb < M else 0.0) == 0.0
    for k, mine in (("means3D", means.grad), ("opacity", op.grad), ("scales", sc.grad), ("rotations", rot.grad)):
        assert rel_(mine.reshape(rg[k].shape), rg[k]) < 1e-4, (k, rel_(mine.reshape(rg[k].shape), rg[k]))


@pytest.mark.parametrize("C_", [3, 1])
def test_reference_wrapper_package_runs_unchanged_on_the_C_level_dropin(libfnx, C_):
    """The level at which the reference itself binds native code: its wrapper package (`diff_gaussian_rasterization_chN/__init__.py`,
    unmodified, staged bytecode) on top of libfnx's `_C` drop-in (rasterize_gaussians / rasterize_gaussians_backward / mark_visible
    with the pybind signatures of R3/rasterize_points.h:18-64, buffers handed back to the backward like the reference's).  Same
    kernels as the package-level drop-in: images / depth / radii must be bit-identical, gradients equal up to the order of the
    float reductions."""
    from oracle import ref_python
    if not ref_python.staged():
        pytest.skip("staged reference Python (oracle/_ref) not present")
    from fluidnexus_b200 import install_compat
    install_compat()
    import importlib
    pkg = f"diff_gaussian_rasterization_ch{C_}"
    wrap = ref_python.reference_wrapper_on(pkg, importlib.import_module(pkg + "._C"), f"refwrap_ch{C_}")
    drop = importlib.import_module(pkg)
    assert wrap.GaussianRasterizer is not drop.GaussianRasterizer
    name = "mixed_ch3_96" if C_ == 3 else "fluid_ch1_112"
    gs, cam, bg, inp = scenes.build(name)
    dL = _t(scenes.dL_dpix(name, (C_, inp["H"], inp["W"])))
    res = {}
    for tag, m in (("wrap", wrap), ("drop", drop)):
        t = {k: _t(inp[k]).requires_grad_(True) for k in ("means3D", "colors", "opacities", "scales", "rotations")}
        rs = m.GaussianRasterizationSettings(
            image_height=inp["H"], image_width=inp["W"], tan_fov_x=inp["tan_fov_x"], tan_fov_y=inp["tan_fov_y"], bg=_t(inp["bg"]),
            scale_modifier=1.0, view_matrix=_t(inp["view"]), proj_matrix=_t(inp["proj"]), sh_degree=0, campos=torch.zeros(3, device="cuda"),
            prefiltered=False)
        rz = m.GaussianRasterizer(raster_settings=rs)
        means2D = torch.zeros_like(t["means3D"], requires_grad=True) + 0
        means2D.retain_grad()
        color, radii, depth = rz(means3D=t["means3D"], means2D=means2D, shs=None, colors_precomp=t["colors"], opacities=t["opacities"],
                                 scales=t["scales"], rotations=t["rotations"], cov3D_precomp=None)
        (color * dL).sum().backward()
        res[tag] = dict(color=color.detach(), radii=radii, depth=depth, vis=rz.mark_visible(t["means3D"].detach()), means2D=means2D.grad,
                        **{k: v.grad for k, v in t.items()})
    w, d = res["wrap"], res["drop"]
    assert torch.equal(w["color"], d["color"]) and torch.equal(w["radii"], d["radii"]) and torch.equal(w["depth"], d["depth"])
    assert torch.equal(w["vis"], d["vis"]) and w["vis"].dtype == torch.bool
    for k in ("means3D", "means2D", "colors", "opacities", "scales", "rotations"):
        assert w[k] is not None and w[k].shape == d[k].shape, k
        assert rel(w[k].cpu().numpy(), d[k].cpu().numpy()) < 2e-5, k
    # and against the oracle, like every other path
    o = RasterOracle("f64")
    ref = o.forward(**inp)
    assert np.abs(w["color"].cpu().numpy() - ref["color"]).max() < PIX_TOL
    gref = o.backward(dL.cpu().numpy())
    assert rel(w["means3D"].cpu().numpy(), gref["means3D"]) < GRAD_TOL and rel(w["rotations"].cpu().numpy(), gref["rotations"]) < GRAD_TOL
    # P == 0 short-circuit through the reference's wrapper (rasterize_points.cu:81,160): zero image, empty buffers, empty grads
    C3 = importlib.import_module(pkg + "._C")
    e = lambda *s: torch.zeros(s, device="cuda")
    nr, col, rad, geo, binb, img, dep = C3.rasterize_gaussians(_t(inp["bg"]), e(0, 3), e(0, C_), e(0, 1), e(0, 3), e(0, 4), 1.0, torch.Tensor([]),
                                                               _t(inp["view"]), _t(inp["proj"]), inp["tan_fov_x"], inp["tan_fov_y"], inp["H"],
                                                               inp["W"], torch.Tensor([]), 0, e(3), False)
    assert nr == 0 and col.abs().max() == 0 and rad.numel() == 0 and geo.numel() == 0
    g8 = C3.rasterize_gaussians_backward(_t(inp["bg"]), e(0, 3), rad, e(0, C_), e(0, 3), e(0, 4), 1.0, torch.Tensor([]), _t(inp["view"]),
                                         _t(inp["proj"]), inp["tan_fov_x"], inp["tan_fov_y"], dL, torch.Tensor([]), 0, e(3), geo, nr, binb, img)
    assert len(g8) == 8 and all(x.numel() == 0 for x in g8)
