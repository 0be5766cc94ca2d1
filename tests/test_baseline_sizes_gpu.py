"""GPU: parity at BASELINE.json's sizes, inside the driver-run suite (VERDICT r1 "weak" #2).

Rasterizer: configs 2-5 (c2 = 50k Gaussians 5x400x400 ch3; c3 = 150k ch1 512x512; c4 = 200k = 20k fluid + 180k frozen ch3; c5 =
300k = 30k + 270k incl. the ball, ch3), all five views in ONE batched libfnx call, against the COMPILED, UNMODIFIED REFERENCE
(oracle/_ref/ch{1,3}/*.so, run live on the same GPU, one view at a time as the reference does).  Gates:
  * per-Gaussian preprocess state (radius, screen xy, view depth, conic + opacity): BIT-equal to the reference's geomBuffer;
  * rendered pixels and median depth: BIT-equal to the reference's (north_star's gate is max|delta| < 1e-3).  This holds because
    libfnx rounds `power` the way the reference's SASS does and evaluates the reference's accurate expf on every kept pair
    (raster.cu:pair_alpha); with one fused product the other way round, single pixels differed by up to 1.35e-3 at these sizes;
  * gradients (summed over the five views) rel-L2 < 1e-4 per tensor, with the reference's own run-to-run difference (its
    backward sums with float atomics) printed next to it.
The static + dynamic stream path (MergedRasterWorkspace) is held to the same reference on c4 / c5.
Physics: P1-P3 + gradient at N = 28 000 hidden / V = 20 000 visual particles against the fp64 oracle (max_num_neighbors
binding and not), and BASELINE config 1 (32^3 lattice, P2 + 0.1 P3 + 0.1 P4).
"""
import numpy as np
import pytest
import torch

from fluidnexus_b200 import physics as P
from fluidnexus_b200 import rasterizer as R
from fluidnexus_b200 import synthetic as S
from oracle import pbf_ref as O
from oracle import ref_ext

pytestmark = pytest.mark.gpu

SIZES = {
    #      nf       nb      C  size
    "c2": (50_000, 0, 3, 400),
    "c3": (150_000, 0, 1, 512),
    "c4": (20_000, 180_000, 3, 512),
    "c5": (30_000, 270_000, 3, 512),
}


def _sets(name):
    nf, nb, C, size = SIZES[name]
    fluid = S.fluid_gaussians(nf, C, seed=0)
    bg = None
    if nb:
        bg = S.background_gaussians(nb, C, seed=1) if name != "c5" else S.cat_sets(S.background_gaussians(nb - 30_000, C, seed=1),
                                                                                  S.ball_gaussians(30_000, C, seed=4))
    return fluid, bg, C, size


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


def _reference_views(C, gs, cams, dL):
    """The compiled reference, one view at a time.  Returns per-view outputs and the gradients summed over the views (twice:
    its run-to-run noise)."""
    rr = ref_ext.RefRaster(C)
    inp0 = S.raster_inputs(gs, cams[0], np.zeros(C, np.float32))
    common = (_t(inp0["bg"]), _t(inp0["means3D"]), _t(inp0["colors"]), _t(inp0["opacities"]), _t(inp0["scales"]), _t(inp0["rotations"]))
    outs, sums = [], [None, None]
    for v, cam in enumerate(cams):
        view, proj = cam.world_view_transform.cuda(), cam.full_proj_transform.cuda()
        ro = rr.forward(*common, 1.0, view, proj, inp0["tan_fov_x"], inp0["tan_fov_y"], inp0["H"], inp0["W"])
        depth, m2, co = ref_ext.carve_geom(ro["geom"], gs.P)
        outs.append(dict(color=ro["color"].clone(), depth=ro["depth"].clone(), radii=ro["radii"].clone(), gdepth=depth, xy=m2, conic=co,
                         R=ro["num_rendered"]))
        for rep in range(2):
            g = rr.backward(dL[v].contiguous())
            if sums[rep] is None:
                sums[rep] = {k: t.clone() for k, t in g.items() if k != "means2D"}
            else:
                for k in sums[rep]:
                    sums[rep][k] += g[k]
    return common, inp0, outs, sums


@pytest.mark.skipif(not (ref_ext.available("ch3") and ref_ext.available("ch1")), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("name", list(SIZES))
def test_rasterizer_matches_compiled_reference_at_baseline_size(libfnx, name):
    fluid, bg, C, size = _sets(name)
    gs = fluid if bg is None else S.cat_sets(fluid, bg)
    cams = S.make_cameras(5, size)
    gen = torch.Generator("cuda").manual_seed(7)
    dL = torch.randn((5, C, size, size), device="cuda", generator=gen)
    common, inp0, ref, sums = _reference_views(C, gs, cams, dL)
    view_all = torch.stack([c.world_view_transform for c in cams]).cuda().contiguous()
    proj_all = torch.stack([c.full_proj_transform for c in cams]).cuda().contiguous()
    ctx, col, rad, dep = R.raster_forward(C, common[0], common[1], common[2], common[3], common[4], common[5], 1.0, None, view_all, proj_all,
                                          inp0["tan_fov_x"], inp0["tan_fov_y"], size, size, speculative=False)
    g = R.read_geom(ctx)
    report, depth_mismatch = [], 0
    for v in range(5):
        r = ref[v]
        vis = (r["radii"] > 0).cpu().numpy()
        assert int((rad[v] != r["radii"]).sum()) == 0, (name, v, "radii")
        u = lambda a: np.ascontiguousarray(a).view(np.uint32)
        assert int((u(g["xy"][v].cpu().numpy()[vis]) != u(r["xy"][vis])).sum()) == 0, (name, v, "screen xy bits")
        assert int((u(g["depth"][v].cpu().numpy()[vis]) != u(r["gdepth"][vis])).sum()) == 0, (name, v, "depth bits")
        assert int((u(g["conic_opacity"][v].cpu().numpy()[vis]) != u(r["conic"][vis])).sum()) == 0, (name, v, "conic bits")
        d = float((col[v] - r["color"]).abs().max())
        assert d == 0.0 and torch.equal(col[v], r["color"]), (name, v, d)
        nd = int((dep[v] != r["depth"]).sum())
        assert nd == 0, (name, v, "median depth", nd)
        depth_mismatch += nd
        assert ctx.num_rendered <= sum(x["R"] for x in ref)          # opacity-aware tile culling never adds instances
        report.append(d)
    gf = R.raster_backward(ctx, dL)
    for k in ("means3D", "colors", "opacity", "scales", "rotations"):
        noise = _rel(sums[1][k], sums[0][k])
        err = _rel(gf[k].reshape(sums[0][k].shape), sums[0][k])
        print(f"{name} grad {k:9s} rel-L2 vs reference {err:.2e} (reference run-to-run {noise:.2e})")
        assert err < 1e-4, (name, k, err, noise)
    print(f"{name}: image max|d| per view {['%.1e' % x for x in report]}, instances fnx {ctx.num_rendered} vs reference {sum(x['R'] for x in ref)}, "
          f"median-depth pixels that differ: {depth_mismatch} of {5 * size * size}")
    if bg is None:
        return
    # ---- static + dynamic streams: the frozen set binned once, fluid rows re-binned and merged per tile ----
    V = fluid.P
    sl = lambda t, a, b: t[a:b].contiguous()
    means, colors, opac, scales, rots = common[1], common[2], common[3].reshape(-1).contiguous(), common[4], common[5]
    dyn = dict(means3D=sl(means, 0, V), colors=sl(colors, 0, V), opacities=sl(opac, 0, V), scales=sl(scales, 0, V), rotations=sl(rots, 0, V))
    sta = dict(means3D=sl(means, V, gs.P), colors=sl(colors, V, gs.P), opacities=sl(opac, V, gs.P), scales=sl(scales, V, gs.P),
               rotations=sl(rots, V, gs.P))
    ws = R.MergedRasterWorkspace(torch.device("cuda"), V, 5, size, size, common[0], dyn, sta, view_all, proj_all, inp0["tan_fov_x"], inp0["tan_fov_y"])
    ws.forward(dyn["means3D"], dyn["colors"], dyn["opacities"], dyn["scales"], dyn["rotations"])
    for v in range(5):
        assert torch.equal(ws.color[v], ref[v]["color"]), (name, "merged", v, float((ws.color[v] - ref[v]["color"]).abs().max()))
        assert torch.equal(ws.depth[v], ref[v]["depth"]), (name, "merged depth", v)
    gm = ws.backward(dL)["means3D"]
    torch.cuda.synchronize()
    assert not ws.overflowed()
    err = _rel(gm, sums[0]["means3D"][:V])
    print(f"{name} merged streams: dL/dmeans3D (fluid rows) rel-L2 vs reference {err:.2e}")
    assert err < 1e-4, (name, "merged means3D", err)


def _state(hp, nvis, seed=0):
    rng = np.random.default_rng(seed)
    vis = hp.xyz[rng.choice(hp.N, nvis, replace=False)] + rng.uniform(-0.4, 0.4, (nvis, 3))
    d = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32).double()     # float32-rounded, as the device sees them
    return dict(xyz=d(hp.xyz), estimate_xyz=d(hp.estimate_xyz), buoyancy=d(hp.buoyancy), force=d(hp.force), imass=d(hp.imass), visual_xyz=d(vis))


@pytest.mark.parametrize("K,bmax", [(100, 0.0), (100, 0.8), (30, 0.8)])
def test_physics_terms_at_baseline_size(libfnx, K, bmax):
    """P1-P3 + gradient at N = 28 000 / V = 20 000 (bench workloads) vs the fp64 oracle; K = 30 makes max_num_neighbors bind
    (a lattice particle has ~46 neighbours within H)."""
    prm = O.PBFParams(KNN_K=K, buoyancy_max_y=bmax, p0=1.5)
    hp = S.hidden_lattice(28_000, seed=200, buoyancy=(0.0, 1.96, 0.0) if bmax > 0 else (0.0, 0.0, 0.0))
    hp.force[:] = np.random.default_rng(1).normal(0, 3, hp.force.shape)
    st64 = _state(hp, 20_000)
    e64 = (st64["estimate_xyz"].float() / 100).double().requires_grad_(True)
    w = torch.tensor(np.random.default_rng(2).normal(size=(20_000, 3)))
    o1 = O.visual_xyz_from_nn(prm, e64, st64["xyz"], st64["visual_xyz"])
    pr = O.gas_constraints_from_exyz_nn(prm, e64, st64["imass"])
    pn = O.gas_constraints_from_vel_nn_guess(prm, e64, st64["xyz"], st64["buoyancy"], st64["force"], st64["imass"])
    loss = (o1 * w).sum() * 1e-3 + O.l2_loss(pr, torch.ones_like(pr)) + 0.1 * O.l2_loss(pn, torch.ones_like(pn))
    loss.backward()
    st = {k: v.float().cuda() for k, v in st64.items()}
    e = e64.detach().float().cuda().requires_grad_(True)
    X = e * 100.0
    a1 = P.visual_advect(X, st["xyz"], st["visual_xyz"], prm.H, prm.secs, K)
    got = P.density_ratio(X, st["imass"], prm.H, prm.p0, K)
    Y = O.guess_hidden_particles_from_nn(prm, e, st["xyz"], st["buoyancy"], st["force"])
    gotn = P.density_ratio(Y, st["imass"], prm.H, prm.p0, K)
    l = (a1 * w.float().cuda()).sum() * 1e-3 + O.l2_loss(got, torch.ones_like(got)) + 0.1 * O.l2_loss(gotn, torch.ones_like(gotn))
    l.backward()
    assert _rel(a1.detach().cpu(), o1.detach()) < 1e-6
    assert _rel(got.detach().cpu(), pr.detach()) < 1e-5
    binds = K < 46
    if not binds:
        assert _rel(gotn.detach().cpu(), pn.detach()) < 1e-5
        assert abs(float(l) - float(loss)) < 1e-4 * abs(float(loss))
        assert _rel(e.grad.cpu(), e64.grad) < 1e-4
    else:
        # With the cap binding the neighbour set is "the first K hits in index order": a neighbour that sits within fp32 rounding of
        # the radius H changes WHICH later neighbours are kept, so the next-tick density (positions Y computed in fp32 on the device,
        # in fp64 by the oracle) is discontinuous in the last bits of Y.  The kernel itself is held to the oracle on IDENTICAL
        # positions; the composite only has to agree on all but a few particles.
        pn_same = O._density_ratio(prm, Y.detach().cpu().double(), st64["imass"])
        assert _rel(gotn.detach().cpu(), pn_same) < 1e-5
        bad = ((gotn.detach().cpu().double() - pn.detach()).abs() > 1e-4 * pn.detach().abs()).double().mean()
        assert float(bad) < 2e-3, float(bad)
        assert abs(float(l) - float(loss)) < 1e-3 * abs(float(loss))
        assert _rel(e.grad.cpu(), e64.grad) < 2e-2
    # the neighbour lists themselves, exactly (index-order cap)
    ref_e = O.radius_graph(X.detach().cpu(), prm.H, loop=True, max_num_neighbors=K)
    got_e = P.radius_graph(X.detach(), prm.H, loop=True, max_num_neighbors=K).cpu()
    assert torch.equal(got_e, ref_e)


def test_config1_lattice_32cubed(libfnx):
    """BASELINE config 1: 32^3 = 32 768-particle lattice, loss = P2 + 0.1 P3 + 0.1 P4, value and gradient."""
    prm = O.PBFParams()
    hp = S.cube_lattice(32, seed=3)
    d = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32).double()
    xyz, est, buo, force, imass = d(hp.xyz), d(hp.estimate_xyz), d(hp.buoyancy), d(hp.force), d(hp.imass)
    e64 = ((est.float() / 100) + 1e-4 * torch.tensor(np.random.default_rng(0).normal(size=est.shape)).float()).double().requires_grad_(True)

    def loss_of(e, f):
        pr = f["p2"](e)
        pn = f["p3"](e)
        return O.l2_loss(pr, torch.ones_like(pr)) + 0.1 * O.l2_loss(pn, torch.ones_like(pn)) + 0.1 * O.l2_loss(e * 100.0, f["est"])
    ref = loss_of(e64, dict(p2=lambda e: O.gas_constraints_from_exyz_nn(prm, e, imass),
                            p3=lambda e: O.gas_constraints_from_vel_nn_guess(prm, e, xyz, buo, force, imass), est=est))
    ref.backward()
    c = lambda t: t.float().cuda()
    e = e64.detach().float().cuda().requires_grad_(True)
    got = loss_of(e, dict(p2=lambda e_: P.density_ratio(e_ * 100.0, c(imass), prm.H, prm.p0, prm.KNN_K),
                          p3=lambda e_: P.density_ratio(O.guess_hidden_particles_from_nn(prm, e_, c(xyz), c(buo), c(force)), c(imass), prm.H, prm.p0,
                                                        prm.KNN_K), est=c(est)))
    got.backward()
    assert abs(float(got) - float(ref)) < 1e-4 * abs(float(ref))
    assert _rel(e.grad.cpu(), e64.grad) < 1e-4
