"""CPU: pins for the physics oracle (oracle/pbf_ref.py).  The neighbour search is PARITY-UNPINNED against
torch_cluster (not installable here); these tests anchor the restatement on closed forms and internal consistency."""
import math

import numpy as np
import pytest
import torch

from fluidnexus_b200 import synthetic as S
from oracle import pbf_ref as O


def test_radius_semantics_first_k_in_index_order():
    x = torch.tensor([[0.0, 0, 0], [0.5, 0, 0], [1.0, 0, 0], [1.9, 0, 0], [2.0, 0, 0], [5.0, 0, 0]])
    y = torch.tensor([[0.0, 0, 0], [5.0, 0, 0]])
    e = O.radius(x, y, 2.0, max_num_neighbors=10)
    # strict < r: the point at exactly distance 2.0 is excluded
    assert e.tolist() == [[0, 0, 0, 0, 1], [0, 1, 2, 3, 5]]
    e = O.radius(x, y, 2.0, max_num_neighbors=2)  # cap keeps the first two BY INDEX, not the nearest
    assert e.tolist() == [[0, 0, 1], [0, 1, 5]]
    g = O.radius_graph(x, 2.0, loop=True, max_num_neighbors=10)
    assert g.shape[0] == 2 and (g[0] == g[1]).sum() == 6   # self loops kept; row = neighbour, col = query
    g2 = O.radius_graph(x, 2.0, loop=False, max_num_neighbors=10)
    assert (g2[0] == g2[1]).sum() == 0 and g2.shape[1] == g.shape[1] - 6


def test_two_particle_density_closed_form():
    prm = O.PBFParams(H=2.0, p0=1.5)
    d = 0.7
    e = torch.tensor([[0.0, 0, 0], [d / 100.0, 0, 0]])
    pr = O.gas_constraints_from_exyz_nn(prm, e, torch.ones(2, 1))
    t1 = 315.0 / (64 * math.pi * 2.0 ** 9)
    expect = (t1 * 4.0 ** 3 + t1 * (4.0 - d * d) ** 3) / 1.5
    assert torch.allclose(pr, torch.full((2, 1), expect), rtol=1e-5)


def test_advect_uniform_velocity_moves_visual_by_secs_u():
    """If every hidden particle has the same velocity u, the poly6-weighted mean is u and visual moves by secs*u."""
    prm = O.PBFParams(secs=0.033)
    hp = S.cube_lattice(8, seed=1)
    xyz = torch.from_numpy(hp.xyz).float()
    u = torch.tensor([1.0, 30.0, -2.0])
    est = (xyz + prm.secs * u) / 100.0
    vis = xyz[:50] + 0.3
    out = O.visual_xyz_from_nn(prm, est, xyz, vis)
    inside = ((out - vis) - prm.secs * u).abs().max()
    assert inside < 2e-4


def test_distance_loss_matches_reference_formula_small():
    torch.manual_seed(0)
    p = torch.rand(40, 3) * 0.01
    thr = 0.004
    d = torch.cdist(p.double(), p.double(), p=2, compute_mode="donot_use_mm_for_euclid_dist")
    mask = d < thr
    mask.fill_diagonal_(False)
    ref = ((thr - d) * mask.double()).clamp(min=0).pow(2).sum()
    assert abs(O.distance_loss(p.double(), thr) - ref) < 1e-12


def test_guess_next_tick_formula():
    prm = O.PBFParams(secs=0.033, buoyancy_max_y=0.8)
    e = torch.tensor([[0.3, 0.4, -0.2]])
    xyz = torch.tensor([[29.0, 39.5, -20.5]])
    b = torch.tensor([[0.0, 1.96, 0.0]])
    f = torch.tensor([[0.5, 0.0, 0.0]])
    y = O.guess_hidden_particles_from_nn(prm, e, xyz, b, f)
    coeff = 1 - 0.4 / 0.8
    v = (e * 100 - xyz) / 0.033 + b * coeff * 0.033 + 0.033 * f
    assert torch.allclose(y, e * 100 + 0.033 * v)


def test_knn3_oracle():
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3], [10, 10, 10]], float)
    out = O.knn3_mean_dist2(pts)
    assert abs(out[0] - (1 + 4 + 9) / 3) < 1e-9


def test_physics_loss_terms_backward_runs():
    prm = O.PBFParams()
    hp = S.cube_lattice(6, seed=4)
    st = dict(xyz=torch.from_numpy(hp.xyz).float(), estimate_xyz=torch.from_numpy(hp.estimate_xyz).float(),
              buoyancy=torch.from_numpy(hp.buoyancy).float(), force=torch.from_numpy(hp.force).float(),
              imass=torch.from_numpy(hp.imass).float(), visual_xyz=torch.from_numpy(hp.xyz[:30] + 0.2).float())
    e = (st["estimate_xyz"] / 100).clone().requires_grad_(True)
    total, terms, render_xyz, p, pn = O.physics_loss_terms(prm, e, st)
    total.backward()
    assert torch.isfinite(e.grad).all() and e.grad.abs().sum() > 0
    assert p.shape == (216, 1) and render_xyz.shape == (30, 3)


def test_kdtree_radius_path_equals_literal_scan():
    """oracle.radius switches to a k-d tree candidate search for large inputs; both paths must give the same edges,
    including under a binding max_num_neighbors cap."""
    rng = np.random.default_rng(3)
    x = torch.tensor(rng.uniform(0, 6, (900, 3)), dtype=torch.float32)
    y = torch.tensor(rng.uniform(0, 6, (700, 3)), dtype=torch.float32)   # 630k pairs > literal-scan limit
    for K in (100, 5):
        got = O.radius(x, y, 2.0, K)
        rows, cols = [], []
        xn, yn = x.numpy(), y.numpy()
        for c in range(yn.shape[0]):
            d = ((xn - yn[c]) ** 2).sum(1, dtype=np.float32)
            hit = np.nonzero(d < np.float32(4.0))[0][:K]
            rows += [c] * len(hit)
            cols += hit.tolist()
        assert got.tolist() == [rows, cols]


def test_solver_two_particle_closed_form():
    """project_gas_constraints on two particles d apart (gm_fluid.py:896-996): pi = poly6(0) + poly6(d^2), one spiky
    gradient each (equal and opposite), lambda and the position correction in closed form; confirm / still-particle
    rule of gm_fluid.py:1160-1175."""
    import math
    sp = O.SolverParams(p0=1.5, k=10.0)
    d = 0.7
    st = dict(xyz=torch.tensor([[0.0, 0.0, 0.0], [d, 0.0, 0.0]], dtype=torch.float64), velocity=torch.ones(2, 3, dtype=torch.float64),
              force=torch.zeros(2, 3, dtype=torch.float64), buoyancy=torch.zeros(2, 3, dtype=torch.float64),
              imass=torch.ones(2, 1, dtype=torch.float64), counts=torch.ones(2, 1, dtype=torch.float64),
              visual_xyz=torch.zeros(0, 3, dtype=torch.float64))
    st["estimate_xyz"] = st["xyz"].clone()
    p_ratio, lam = O.solver_project_gas_constraints(sp, st)
    p6 = lambda r2: sp.poly6_term1 * (sp.H2 - r2) ** 3
    pr = (p6(0.0) + p6(d * d)) / sp.p0
    assert abs(float(p_ratio[0]) - pr) < 1e-12 and abs(float(p_ratio[1]) - pr) < 1e-12
    rlen = math.sqrt(d * d + sp.EPSILON)
    g = sp.spiky_grad_term1 * (sp.H - rlen) ** 2 * (d / (rlen + sp.EPSILON))       # |spiky|, pointing away from the neighbour...
    lam_ref = -(pr - 1.0) / (2.0 * (g / sp.p0) ** 2 + sp.RELAXATION)               # grad_dot + gr_dot, one neighbour
    assert abs(float(lam[0]) - lam_ref) < 1e-9 * abs(lam_ref)
    corr = -sp.K_P * (p6(d * d) / sp.lamb_corr_denom) ** sp.E_P
    # particle 0: diff = x0 - x1 = -d  =>  spiky = +g along x;  neighbours_len = 2 (self loop + 1), counts = 1
    dx0 = (2.0 * lam_ref + corr) * g / sp.p0 / 3.0
    assert abs(float(st["estimate_xyz"][0, 0]) - dx0) < 1e-9 * abs(dx0)
    assert abs(float(st["estimate_xyz"][1, 0]) - (d - dx0)) < 1e-9
    assert torch.allclose(st["force"], torch.ones(2, 3, dtype=torch.float64) * (1.0 - pr) * -sp.k)
    st["estimate_xyz"][1] = st["xyz"][1]          # particle 1 did not move
    O.solver_confirm_guess_hidden_particles(sp, st)
    assert torch.all(st["velocity"][1] == 0) and abs(float(st["velocity"][0, 0]) - dx0 / sp.secs) < 1e-9
