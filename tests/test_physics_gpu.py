"""GPU: parity of libfnx's neighbour search and fused physics terms (through the C ABI) against oracle/pbf_ref.py.

Index results are compared exactly; fp32 sums against the fp64-evaluated oracle with rel-L2 < 1e-5 (values) and
< 1e-4 (gradients)."""
import numpy as np
import pytest
import torch

from fluidnexus_b200 import physics as P
from fluidnexus_b200 import synthetic as S
from oracle import pbf_ref as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def _state(hp, nvis, seed=0, dtype=torch.float32):
    rng = np.random.default_rng(seed)
    vis = hp.xyz[rng.choice(hp.N, nvis, replace=False)] + rng.uniform(-0.4, 0.4, (nvis, 3))
    return dict(xyz=torch.tensor(hp.xyz, dtype=dtype), estimate_xyz=torch.tensor(hp.estimate_xyz, dtype=dtype),
                buoyancy=torch.tensor(hp.buoyancy, dtype=dtype), force=torch.tensor(hp.force, dtype=dtype),
                imass=torch.tensor(hp.imass, dtype=dtype), visual_xyz=torch.tensor(vis, dtype=dtype))


@pytest.mark.parametrize("K", [100, 7])
@pytest.mark.parametrize("n", [1, 300, 4096])
def test_radius_matches_oracle_exactly(libfnx, n, K):
    rng = np.random.default_rng(n + K)
    x = torch.tensor(rng.uniform(0, 10, (n, 3)), dtype=torch.float32)
    y = torch.tensor(rng.uniform(0, 10, (max(1, n // 2), 3)), dtype=torch.float32)
    ref = O.radius(x, y, 2.0, K)
    got = P.radius(x.cuda(), y.cuda(), 2.0, max_num_neighbors=K).cpu()
    assert got.dtype == torch.long and torch.equal(got, ref)
    for loop in (True, False):
        rg, gg = O.radius_graph(x, 2.0, loop=loop, max_num_neighbors=K), P.radius_graph(x.cuda(), 2.0, loop=loop, max_num_neighbors=K).cpu()
        assert torch.equal(gg, rg)


def test_radius_edge_cases(libfnx):
    e = P.radius(torch.zeros(0, 3).cuda(), torch.zeros(4, 3).cuda(), 1.0)
    assert e.shape == (2, 0)
    x = torch.zeros(50, 3).cuda()  # all coincident: every pair is within r, cap binds -> first K indices
    e = P.radius(x, x[:2], 0.5, max_num_neighbors=8).cpu()
    assert e[1].view(2, 8).tolist() == [list(range(8))] * 2
    far = torch.tensor([[1e4, -3e3, 7.0], [-1e4, 5.0, 2.0]]).cuda()  # hashing: huge, negative coordinates
    e = P.radius(far, far, 1.0).cpu()
    assert e.tolist() == [[0, 1], [0, 1]]


def test_scatter_min_and_knn(libfnx):
    src = torch.tensor([5.0, -1.0, 3.0, -1.0, 9.0, 2.0]).cuda()
    idx = torch.tensor([0, 2, 0, 2, 4, 4]).cuda()
    out, arg = P.scatter_min(src, idx, dim=0, dim_size=6)
    ro, ra = O.scatter_min(src.cpu(), idx.cpu(), dim_size=6)
    assert torch.equal(out.cpu(), ro) and torch.equal(arg.cpu(), ra)
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.uniform(0, 1, (5000, 3)), rng.uniform(50, 50.001, (3, 3)), [[500.0, 0, 0]]]).astype(np.float32)
    got = P.distCUDA2(torch.tensor(pts).cuda()).cpu().numpy()
    assert rel(got, O.knn3_mean_dist2(pts)) < 1e-5


@pytest.mark.parametrize("shape", ["planar", "line", "outliers", "coincident"])
def test_distcuda2_degenerate_clouds(libfnx, shape):
    """planar / linear clouds and clouds with far outliers (an initial point cloud on a wall; a few stray points) must
    neither change the result nor send every query into the exhaustive fallback (the call has to stay fast)."""
    import time
    rng = np.random.default_rng(4)
    n = 60_000
    if shape == "planar":
        pts = np.stack([rng.uniform(0, 3, n), rng.uniform(0, 2, n), np.full(n, 0.7)], 1)
    elif shape == "line":
        pts = np.stack([rng.uniform(0, 50, n), np.full(n, 1.0), np.full(n, -2.0)], 1)
    elif shape == "outliers":
        pts = np.concatenate([rng.uniform(0, 1, (n - 5, 3)), rng.uniform(1e3, 1e4, (5, 3))])
    else:
        pts = np.concatenate([np.zeros((n - 3, 3)), rng.uniform(0, 1, (3, 3))])
    pts = pts.astype(np.float32)
    P.distCUDA2(torch.tensor(pts[:100]).cuda())
    torch.cuda.synchronize()
    t0 = time.time()
    got = P.distCUDA2(torch.tensor(pts).cuda()).cpu().numpy()
    dt = time.time() - t0
    sub = rng.choice(n, 300, replace=False)            # brute-force check on a sample
    d2 = ((pts[sub, None, :].astype(np.float64) - pts[None, :, :].astype(np.float64)) ** 2).sum(-1)
    d2[np.arange(300), sub] = np.inf
    ref = np.sort(d2, axis=1)[:, :3].mean(1)
    assert np.allclose(got[sub], ref, rtol=1e-4, atol=1e-12), shape
    assert dt < 5.0, (shape, dt)


@pytest.mark.parametrize("K,bmax", [(100, 0.0), (100, 0.8), (20, 0.8)])
def test_density_and_next_tick_match_oracle(libfnx, K, bmax):
    prm = O.PBFParams(KNN_K=K, buoyancy_max_y=bmax, p0=1.5)
    hp = S.hidden_lattice(3000, seed=5, buoyancy=(0.0, 1.96, 0.0))
    hp.force[:] = np.random.default_rng(1).normal(0, 3, hp.force.shape)
    st64 = _state(hp, 500, dtype=torch.float64)
    e64 = (st64["estimate_xyz"] / 100).clone().requires_grad_(True)
    pr = O.gas_constraints_from_exyz_nn(prm, e64, st64["imass"])
    pn = O.gas_constraints_from_vel_nn_guess(prm, e64, st64["xyz"], st64["buoyancy"], st64["force"], st64["imass"])
    loss = O.l2_loss(pr, torch.ones_like(pr)) + 0.1 * O.l2_loss(pn, torch.ones_like(pn))
    loss.backward()

    st = {k: v.float().cuda() for k, v in st64.items()}
    e = (st["estimate_xyz"] / 100).clone().requires_grad_(True)
    X = e * 100.0
    got = P.density_ratio(X, st["imass"], prm.H, prm.p0, K)
    assert rel(got.detach().cpu().numpy(), pr.detach().numpy()) < 1e-5
    Y = O.guess_hidden_particles_from_nn(prm, e, st["xyz"], st["buoyancy"], st["force"])  # plain torch ops on GPU
    gotn = P.density_ratio(Y, st["imass"], prm.H, prm.p0, K)
    assert rel(gotn.detach().cpu().numpy(), pn.detach().numpy()) < 1e-5
    l = O.l2_loss(got, torch.ones_like(got)) + 0.1 * O.l2_loss(gotn, torch.ones_like(gotn))
    l.backward()
    assert abs(float(l) - float(loss)) < 1e-5 * abs(float(loss))
    assert rel(e.grad.cpu().numpy(), e64.grad.numpy()) < 1e-4


@pytest.mark.parametrize("K", [100, 12])
def test_visual_advect_matches_oracle(libfnx, K):
    prm = O.PBFParams(KNN_K=K)
    hp = S.hidden_lattice(3000, seed=6)
    st64 = _state(hp, 700, seed=3, dtype=torch.float64)
    e64 = (st64["estimate_xyz"] / 100).clone().requires_grad_(True)
    out64 = O.visual_xyz_from_nn(prm, e64, st64["xyz"], st64["visual_xyz"])
    w = torch.tensor(np.random.default_rng(2).normal(size=out64.shape))
    (out64 * w).sum().backward()
    st = {k: v.float().cuda() for k, v in st64.items()}
    e = (st["estimate_xyz"] / 100).clone().requires_grad_(True)
    out = P.visual_advect(e * 100.0, st["xyz"], st["visual_xyz"], prm.H, prm.secs, K)
    assert rel(out.detach().cpu().numpy(), out64.detach().numpy()) < 1e-6
    (out * w.float().cuda()).sum().backward()
    assert rel(e.grad.cpu().numpy(), e64.grad.numpy()) < 1e-4


def test_visual_advect_isolated_particles(libfnx):
    """visual particles with no hidden neighbour stay put and send no gradient (sum_p6 clamp, gm_fluid.py:1327)."""
    hp = S.cube_lattice(5, seed=2)
    xyz = torch.tensor(hp.xyz, dtype=torch.float32).cuda()
    X = (xyz + 0.5).requires_grad_(True)
    vis = torch.tensor([[1e3, 1e3, 1e3], [hp.xyz[0, 0], hp.xyz[0, 1], hp.xyz[0, 2]]], dtype=torch.float32).cuda()
    out = P.visual_advect(X, xyz, vis, 2.0, 0.033, 100)
    assert torch.equal(out[0], vis[0]) and not torch.equal(out[1], vis[1])
    out[0].sum().backward()
    assert X.grad.abs().max() == 0


@pytest.mark.parametrize("thr", [0.004, 0.0005])
def test_pair_distance_loss_matches_oracle(libfnx, thr):
    rng = np.random.default_rng(7)
    p = rng.uniform(0, 0.02, (1500, 3))
    p[10] = p[11]  # an exact duplicate: value thr^2 (x2), zero gradient
    p64 = torch.tensor(p, dtype=torch.float64, requires_grad=True)
    ref = O.distance_loss(p64, thr)
    ref.backward()
    pg = torch.tensor(p, dtype=torch.float32).cuda().requires_grad_(True)
    got = P.pair_distance_loss(pg, thr)
    got.backward()
    assert abs(float(got) - float(ref)) < 2e-4 * abs(float(ref)) + 1e-12
    assert rel(pg.grad.cpu().numpy(), p64.grad.numpy()) < 2e-4


def test_adam_matches_torch(libfnx):
    import ctypes as C
    from fluidnexus_b200 import _lib as L
    torch.manual_seed(0)
    p0 = torch.randn(1000, 3)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([{"params": [ref], "lr": 1.6e-4}], lr=0.0, eps=1e-15)
    p = p0.clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(1000, 3) * (10.0 ** (step - 3))
        ref.grad = g.clone() * 0.2
        opt.step()
        gc = g.cuda()
        L.check(L.lib().fnx_adam_step(p.numel(), p.data_ptr(), gc.data_ptr(), m.data_ptr(), v.data_ptr(), 0.2, 1.6e-4, 0.9, 0.999,
                                      1e-15, step, torch.cuda.current_stream().cuda_stream))
        assert (p.cpu() - ref.detach()).abs().max() < 2e-7


def test_compat_imports_resolve_to_libfnx(libfnx):
    from fluidnexus_b200 import install_compat
    install_compat()
    from simple_knn._C import distCUDA2
    from torch_cluster import radius, radius_graph
    from torch_scatter import scatter_min
    assert radius is P.radius and radius_graph is P.radius_graph and scatter_min is P.scatter_min and distCUDA2 is P.distCUDA2


@pytest.mark.parametrize("K", [100, 12])
def test_counted_kernels_equal_count_then_gather(libfnx, K):
    """fnx_pbf_density_fwd_counted / fnx_visual_advect_fwd_counted (neighbour count + cut-off + sums in one walk) against
    the fnx_radius_count -> fnx_pbf_density_fwd / fnx_visual_advect_fwd sequences, cap not binding (K = 100) and
    binding (K = 12): identical cut-offs, bitwise-identical sums."""
    from fluidnexus_b200 import _lib as L
    lib = L.lib()
    hp = S.hidden_lattice(3000, seed=9)
    st = {k: v.cuda() for k, v in _state(hp, 700, seed=4).items()}
    X = (st["estimate_xyz"]).contiguous()
    N, V, H = X.size(0), st["visual_xyz"].size(0), 2.0
    stream = torch.cuda.current_stream().cuda_stream
    g = P.Grid(X, H)
    _, kth = g.count(X, H, K)
    _, kthV = g.count(st["visual_xyz"], H, K)
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device="cuda")
    p_ref, p_got, kth_got, flag = z(N), z(N), z(N, dt=torch.int32), z(1, dt=torch.int32)
    im = st["imass"].reshape(-1).contiguous()
    L.check(lib.fnx_pbf_density_fwd(g.buf.data_ptr(), X.data_ptr(), N, im.data_ptr(), kth.data_ptr(), H, 1.5, p_ref.data_ptr(), stream))
    L.check(lib.fnx_pbf_density_fwd_counted(g.buf.data_ptr(), X.data_ptr(), N, im.data_ptr(), K, H, 1.5, kth_got.data_ptr(),
                                            p_got.data_ptr(), flag.data_ptr(), stream))
    assert torch.equal(kth_got, kth) and torch.equal(p_got, p_ref)
    assert int(flag.item()) == (1 if K == 12 else 0)
    vis, xyz = st["visual_xyz"].contiguous(), st["xyz"].contiguous()
    o_ref, n_ref, d_ref, o_got, n_got, d_got = z(V, 3), z(V, 3), z(V), z(V, 3), z(V, 3), z(V)
    kthV_got = z(V, dt=torch.int32)
    L.check(lib.fnx_visual_advect_fwd(g.buf.data_ptr(), X.data_ptr(), xyz.data_ptr(), N, vis.data_ptr(), V, kthV.data_ptr(), H, 0.033, 100.0,
                                      o_ref.data_ptr(), n_ref.data_ptr(), d_ref.data_ptr(), stream))
    L.check(lib.fnx_visual_advect_fwd_counted(g.buf.data_ptr(), X.data_ptr(), xyz.data_ptr(), N, vis.data_ptr(), V, K, H, 0.033, 100.0,
                                              o_got.data_ptr(), n_got.data_ptr(), d_got.data_ptr(), kthV_got.data_ptr(), stream))
    assert torch.equal(kthV_got, kthV)
    assert torch.equal(o_got, o_ref) and torch.equal(n_got, n_ref) and torch.equal(d_got, d_ref)
