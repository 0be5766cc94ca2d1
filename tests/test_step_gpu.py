"""GPU: one fused training step (fluidnexus_b200.step) against the literal CPU restatement of the reference loop
(oracle/step_ref.py: per-view python loop, fp64 oracle rasterizer + torch CPU physics + torch.optim.Adam).

Tolerances: loss scalars rel 1e-4 (SURVEY.md 8(d) parity gate); averaged gradient rel-L2 2e-3 -- dominated by
pixels that sit on the rasterizer's 1/255 alpha cut differently in fp32 and fp64; updated parameters after Adam
within lr*1e-2 of the oracle's."""
import math

import numpy as np
import pytest
import torch

from fluidnexus_b200 import synthetic as S
from fluidnexus_b200.step import FrameState, PhysicalStep, StepParams
from oracle import pbf_ref as O
from oracle import step_ref

pytestmark = pytest.mark.gpu


def _scene(C, with_bg, size=64, N=1500, V=600, seed=0):
    hp = S.hidden_lattice(N, seed=seed + 1, buoyancy=(0.0, 1.96, 0.0))
    rng = np.random.default_rng(seed)
    fluid = S.fluid_gaussians(V, C, seed=seed + 2, log_scale=-4.6)
    # visual particles live inside the hidden lattice so that P1 has neighbours; fluid Gaussians sit on them
    vis = hp.xyz[rng.choice(hp.N, V, replace=False)] + rng.uniform(-0.3, 0.3, (V, 3))
    fluid.xyz = vis / 100.0
    bg = S.background_gaussians(400, C, seed=seed + 3) if with_bg else None
    cams = S.make_cameras(5, size)
    return hp, vis, fluid, bg, cams


def _cam_dict(cam):
    return dict(view=cam.world_view_transform.numpy(), proj=cam.full_proj_transform.numpy(),
                tan_fov_x=math.tan(cam.FoVx / 2), tan_fov_y=math.tan(cam.FoVy / 2), H=cam.image_height, W=cam.image_width)


@pytest.mark.parametrize("C,with_bg,grey,bmax,views", [(3, True, True, 0.0, [0, 2, 4]), (1, False, False, 0.8, [1, 3])])
def test_fused_step_matches_reference_loop(libfnx, oracle_built, C, with_bg, grey, bmax, views):
    hp, vis, fluid, bg, cams = _scene(C, with_bg)
    prm = StepParams(p0=1.5, buoyancy_max_y=bmax, grey=grey, distance_threshold_visual=0.004, lr=1.6e-4)
    oprm = O.PBFParams(p0=1.5, buoyancy_max_y=bmax, distance_threshold_visual=0.004, lr=1.6e-4)
    rng = np.random.default_rng(5)
    gts = [np.clip(0.3 + 0.3 * rng.random((C, cams[0].image_height, cams[0].image_width)), 0, 1).astype(np.float32) for _ in views]

    # ---- oracle ----
    d = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    st = dict(xyz=d(hp.xyz), estimate_xyz=d(hp.estimate_xyz), buoyancy=d(hp.buoyancy), force=d(hp.force), imass=d(hp.imass),
              visual_xyz=d(vis))
    # float32-rounded copies, as the device sees them
    st = {k: v.float().double() for k, v in st.items()}
    allg = fluid if bg is None else S.cat_sets(fluid, bg)
    gauss = dict(scales=d(allg.scales).float().double(), rotations=d(allg.rotations).float().double(),
                 opacity=d(allg.opacity).float().double().reshape(-1), colors=d(allg.colors).float().double(),
                 bg_xyz=None if bg is None else d(bg.xyz).float().double(), bg=np.zeros(C, np.float32))
    fr = FrameState(hp, vis, fluid, bg, prm=prm)
    e0 = fr.e.cpu().double()  # training_setup_current(): estimate_xyz / scale_factor, as rounded on the device
    logs, g_ref, e_ref, _ = step_ref.reference_step(oprm, st, gauss, [_cam_dict(cams[v]) for v in views], [d(g) for g in gts], e0,
                                                   grey=grey)

    # ---- fused ----
    ps = PhysicalStep(cams, C, prm)
    out = ps.step(fr, views, torch.tensor(np.stack(gts)).cuda())
    torch.cuda.synchronize()
    for k, v in (("gas", out["gas"]), ("next_gas", out["next_gas"]), ("exyz", out["exyz"]), ("dist", out["dist"])):
        assert abs(float(v) - logs[0][k]) <= 1e-4 * abs(logs[0][k]) + 1e-9, (k, float(v), logs[0][k])
    for i in range(len(views)):
        assert abs(float(out["l1"][i]) - logs[i]["l1"]) < 1e-4 * logs[i]["l1"]
        assert abs(float(out["ssim"][i]) - logs[i]["ssim"]) < 1e-4  # both are the mean SSIM
    tot_ref = np.mean([l["total"] for l in logs])
    assert abs(float(ps.total_loss(out)) - tot_ref) < 1e-4 * abs(tot_ref)
    g = out["grad"].cpu().double()
    r = (g - g_ref).norm() / g_ref.norm()
    assert r < 2e-3, r
    assert (fr.e.cpu().double() - e_ref).abs().max() < prm.lr * 1e-2


def test_step_is_deterministic_and_descends(libfnx):
    hp, vis, fluid, bg, cams = _scene(3, True, seed=3)
    prm = StepParams(grey=True, distance_threshold_visual=0.004)
    ps = PhysicalStep(cams, 3, prm)
    gt = torch.rand(5, 3, 64, 64, generator=torch.Generator().manual_seed(0)).cuda() * 0.5
    runs = []
    for rep in range(2):
        fr = FrameState(hp, vis, fluid, bg, prm=prm)
        losses = []
        for it in range(8):
            out = ps.step(fr, [0, 1, 2, 3, 4], gt)
            losses.append(float(ps.total_loss(out)))
        runs.append((losses, fr.e.clone()))
    # physics part is atomics-free; the rasterizer's backward uses float atomics, so allow rounding-level drift
    assert np.allclose(runs[0][0], runs[1][0], rtol=1e-5)
    assert (runs[0][1] - runs[1][1]).abs().max() < 1e-6
    assert runs[0][0][-1] < runs[0][0][0]


def test_cuda_graph_replay_equals_eager(libfnx):
    """The captured iteration (one host call per iteration, no host sync) must follow the eager one bit for bit in the
    atomics-free physics and to rounding level in the rasterizer gradient; pinned-host ground truth is accepted."""
    hp, vis, fluid, bg, cams = _scene(3, True, seed=4)
    prm = StepParams(grey=True, distance_threshold_visual=0.004)
    ps = PhysicalStep(cams, 3, prm)
    gt = (torch.rand(5, 3, 64, 64, generator=torch.Generator().manual_seed(1)) * 0.5)
    gt_pinned = gt.pin_memory()
    res = {}
    for mode in ("eager", "graph"):
        fr = FrameState(hp, vis, fluid, bg, prm=prm)
        losses = []
        for it in range(6):
            out = ps.step(fr, [0, 1, 2, 3, 4], gt.cuda() if mode == "eager" else gt_pinned, graph=(mode == "graph"))
            losses.append(float(ps.total_loss(out)))
        res[mode] = (losses, fr.e.clone(), int(fr.step_dev.item()))
    assert res["eager"][2] == res["graph"][2] == 6
    assert np.allclose(res["eager"][0], res["graph"][0], rtol=1e-5)
    assert (res["eager"][1] - res["graph"][1]).abs().max() < 1e-6


def test_view_sharded_gradients_sum_to_the_full_step(libfnx):
    """What the all-reduce adds up: a rank rendering views {0,1,2} and owning the frame's physics terms plus a rank
    rendering views {3,4} without them give the full 5-view gradient (P1's backward is linear in dL/dmeans3D)."""
    hp, vis, fluid, bg, cams = _scene(3, True, seed=6)
    prm = StepParams(grey=True, distance_threshold_visual=0.004)
    ps = PhysicalStep(cams, 3, prm)
    gt = torch.rand(5, 3, 64, 64, generator=torch.Generator().manual_seed(2)).cuda() * 0.5
    full = ps.step(FrameState(hp, vis, fluid, bg, prm=prm), [0, 1, 2, 3, 4], gt, update=False)["grad"].clone()
    a = ps.step(FrameState(hp, vis, fluid, bg, prm=prm), [0, 1, 2], gt[:3], update=False, batch=5, physics=True)["grad"].clone()
    b = ps.step(FrameState(hp, vis, fluid, bg, prm=prm), [3, 4], gt[3:], update=False, batch=5, physics=False)["grad"].clone()
    r = ((a + b) - full).norm() / full.norm()
    assert r < 1e-5, r


def test_static_background_cache_equals_full_rebinning(libfnx):
    """Binning the frozen background once and merging it with the per-iteration fluid stream must give the same image
    bit for bit and the same parameter trajectory as re-binning everything every iteration."""
    hp, vis, fluid, bg, cams = _scene(3, True, seed=8, size=96)
    prm = StepParams(grey=True, distance_threshold_visual=0.004)
    gt = torch.rand(5, 3, 96, 96, generator=torch.Generator().manual_seed(3)).cuda() * 0.5
    res = {}
    for cache in (False, True):
        ps = PhysicalStep(cams, 3, prm, static_cache=cache)
        fr = FrameState(hp, vis, fluid, bg, prm=prm)
        imgs, losses = [], []
        for it in range(4):
            out = ps.step(fr, [0, 1, 2, 3, 4], gt)
            imgs.append(out["images"].clone())
            losses.append(float(ps.total_loss(out)))
        res[cache] = (imgs, losses, fr.e.clone())
    assert torch.equal(res[False][0][0], res[True][0][0])           # first iteration: identical inputs -> identical image
    for a, b in zip(res[False][0], res[True][0]):                   # later ones: parameters agree to ~1e-7, images to rounding
        assert (a - b).abs().max() < 1e-4
    assert np.allclose(res[False][1], res[True][1], rtol=1e-5)
    assert (res[False][2] - res[True][2]).abs().max() < 1e-6


def test_frame_lanes_equal_sequential_frames(libfnx):
    """Independent frames dealt to two CUDA streams (FrameLanes) must give what running them one after the other gives:
    same losses, same parameters after a few iterations (captured graphs, pinned host ground truth)."""
    from fluidnexus_b200.parallel import FrameLanes
    prm = StepParams(grey=True, distance_threshold_visual=0.004)
    scenes_ = [_scene(3, True, seed=20 + k) for k in range(3)]
    cams = scenes_[0][4]
    gts = [(torch.rand(5, 3, 64, 64, generator=torch.Generator().manual_seed(30 + k)) * 0.5).pin_memory() for k in range(3)]
    res = {}
    for n in (1, 2):
        lanes = FrameLanes(lambda k: PhysicalStep(cams, 3, prm), n, "cuda")
        frs = [FrameState(hp, vis, fluid, bg, prm=prm) for (hp, vis, fluid, bg, _) in scenes_]
        losses = []
        for it in range(4):
            # the returned loss tensors live in per-lane scratch that the lane's next frame overwrites: reduce them to
            # the iteration's scalar on the lane's own stream, right behind the step
            outs = lanes.run(list(range(3)), lambda step, k: step.total_loss(step.step(frs[k], [0, 1, 2, 3, 4], gts[k], graph=True)).clone())
            torch.cuda.synchronize()
            losses.append([float(o) for o in outs])
        res[n] = (np.array(losses), [fr.e.clone() for fr in frs])
    assert np.allclose(res[1][0], res[2][0], rtol=1e-5)
    for a, b in zip(res[1][1], res[2][1]):
        assert (a - b).abs().max() < 1e-6


def test_view_subsets_share_the_static_stream_and_any_order_gives_the_same_step(libfnx):
    """one static stream per frame whatever cameras an iteration draws; the same cameras in another order reuse the
    workspace (and give the same update)."""
    hp, vis, fluid, bg, cams = _scene(3, True, seed=5)
    prm = StepParams(grey=True, distance_threshold_visual=0.004)
    ps = PhysicalStep(cams, 3, prm)
    gt = torch.rand(5, 3, 64, 64, device="cuda") * 0.5
    fa, fb = FrameState(hp, vis, fluid, bg, prm=prm), FrameState(hp, vis, fluid, bg, prm=prm)
    for views in ([0, 2, 4], [1, 3], [2, 0, 4], [3]):
        idx = torch.tensor(views, device="cuda")
        ps.step(fa, views, gt[idx])
        ps.step(fb, sorted(views), gt[torch.tensor(sorted(views), device="cuda")])
    torch.cuda.synchronize()
    assert len(fa.ws) == 3 and fa.static_stream is not None and all(w.static is fa.static_stream for w in fa.ws.values())
    assert float((fa.e - fb.e).abs().max()) < 1e-7


def test_overflowed_iteration_is_voided_on_the_device_and_recovered(libfnx):
    """when the fluid spreads over many more tiles than the workspace was sized for, the forward overflows its instance
    capacity and renders only the background.  The update of that iteration must not be applied (Adam is gated by the device-side
    overflow flag: parameters, moments and step count untouched), and the next step re-sizes (re-captures) and carries on."""
    hp, vis, fluid, bg, cams = _scene(3, True, seed=6)
    prm = StepParams(grey=True, distance_threshold_visual=0.004)
    gt = torch.rand(5, 3, 64, 64, device="cuda") * 0.5
    views = [0, 1, 2, 3, 4]
    for graph in (False, True):
        ps = PhysicalStep(cams, 3, prm)
        ps.capacity_slack = 256                      # (the default slack of 64k instances exceeds what this tiny scene can produce)
        fr = FrameState(hp, vis, fluid, bg, prm=prm)
        ps.step(fr, views, gt, graph=graph)
        ps.step(fr, views, gt, graph=graph)
        torch.cuda.synchronize()
        ws = fr.ws[tuple(views)]
        assert not ws.overflowed() and int(fr.step_dev) == 2
        # blow the splats up: every fluid Gaussian now touches every tile -> far beyond the capacity
        scales0 = fr.scales[:fr.V].clone()
        fr.scales[:fr.V] *= 60.0
        e0, m0, v0 = fr.e.clone(), fr.m.clone(), fr.v.clone()
        ps.step(fr, views, gt, graph=graph)
        torch.cuda.synchronize()
        assert ws.overflowed()
        assert torch.equal(fr.e, e0) and torch.equal(fr.m, m0) and torch.equal(fr.v, v0) and int(fr.step_dev) == 2
        fr.scales[:fr.V] = scales0
        ps.step(fr, views, gt, graph=graph)          # notices the overflow, rebuilds the workspace (graph: re-captures), steps
        torch.cuda.synchronize()
        assert fr.skipped_iterations == 1 and not fr.ws[tuple(views)].overflowed() and fr.ws[tuple(views)] is not ws
        assert int(fr.step_dev) == 3 and float((fr.e - e0).abs().max()) > 0


@pytest.mark.parametrize("graph", [False, True])
def test_ground_truth_cache_uploads_once_and_follows_edits(libfnx, graph):
    """step(..., cache_gt=True): a host ground-truth tensor handed in again is served from its device copy,
    an in-place edit of it is seen, and another tensor of the same shape gets its own copy -- the results must equal the uncached
    path's at every step."""
    hp, vis, fluid, bg, cams = _scene(3, True, seed=9)
    prm = StepParams(grey=True, distance_threshold_visual=0.004)
    views = [0, 2, 3]
    gen = torch.Generator().manual_seed(1)
    gt_a = (torch.rand(3, 3, 64, 64, generator=gen) * 0.5).pin_memory()
    gt_b = (torch.rand(3, 3, 64, 64, generator=gen) * 0.5).pin_memory()
    res = {}
    for cached in (False, True):
        ps = PhysicalStep(cams, 3, prm)
        fr = FrameState(hp, vis, fluid, bg, prm=prm)
        a, b = gt_a.clone().pin_memory(), gt_b.clone().pin_memory()
        l1 = []
        for it, gt in enumerate((a, a, a, b, a)):
            if it == 2:
                torch.cuda.synchronize()                       # (the asynchronous uploads of the earlier steps have read `a`)
                a.mul_(0.25)                                   # in-place edit: the cache must upload the new content
            out = ps.step(fr, views, gt, graph=graph, cache_gt=cached)
            l1.append(out["l1"].clone())
        torch.cuda.synchronize()
        res[cached] = (torch.stack(l1).cpu(), fr.e.clone(), len(ps._gt_cache))
    assert res[False][2] == 0 and res[True][2] == 2            # two host tensors seen -> two device copies
    assert torch.allclose(res[False][0], res[True][0], rtol=1e-5, atol=1e-8)
    assert float((res[False][0][1] - res[False][0][2]).abs().max()) > 1e-3     # the edit really changed the loss
    assert (res[False][1] - res[True][1]).abs().max() < 1e-6
