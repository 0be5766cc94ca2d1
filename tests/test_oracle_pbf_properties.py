"""CPU: invariances of the particle-physics terms (oracle/pbf_ref.py, fp64) that hold at any size -- the properties behind the
symmetric-gather gradient kernels (no atomics: dL/dX_i gathers what the scatter form would have added from both edge directions):

* translation / rotation invariance of the density ratio and the pair-distance loss => their gradients sum to zero (no net force)
  and carry no net torque;
* the radius graph is symmetric while the neighbour cap does not bind, so the density gradient is antisymmetric pair by pair;
* P1 moves visual particles by secs x the local velocity: a rigid translation of the hidden estimate translates them all alike;
* the next-tick map is affine in the trainable positions (closed form, with and without the buoyancy height term).
"""
import numpy as np
import torch

from oracle import pbf_ref as O


def _cloud(n=400, seed=0, span=10.0):
    return torch.tensor(np.random.default_rng(seed).uniform(0, span, (n, 3)), dtype=torch.float64)


def _rotation(seed=0):
    q, _ = np.linalg.qr(np.random.default_rng(seed).normal(size=(3, 3)))
    return torch.tensor(q * np.sign(np.linalg.det(q)), dtype=torch.float64)


def test_density_ratio_is_invariant_and_its_gradient_carries_no_net_force_or_torque():
    prm = O.PBFParams(p0=1.5)
    X = _cloud()
    imass = torch.tensor(np.random.default_rng(1).uniform(0.8, 1.2, (X.shape[0], 1)), dtype=torch.float64)
    e = (X / O.SCALE_FACTOR).clone().requires_grad_(True)
    p = O.gas_constraints_from_exyz_nn(prm, e, imass)
    R, t = _rotation(3), torch.tensor([3.0, -7.0, 11.0], dtype=torch.float64)
    p_moved = O.gas_constraints_from_exyz_nn(prm, ((X @ R.T) + t) / O.SCALE_FACTOR, imass)
    assert torch.allclose(p, p_moved, rtol=1e-9, atol=1e-12)
    w = torch.tensor(np.random.default_rng(2).normal(size=p.shape))
    (p * w).sum().backward()
    g = e.grad
    assert float(g.abs().max()) > 0
    assert float(g.sum(0).abs().max()) < 1e-9 * float(g.abs().sum())                                  # translation invariance
    assert float(torch.cross(X, g, dim=1).sum(0).abs().max()) < 1e-9 * float((X.norm(dim=1) * g.norm(dim=1)).sum())   # rotation invariance


def test_radius_graph_is_symmetric_until_the_cap_binds():
    X = _cloud(300, seed=4, span=8.0)
    e = O.radius_graph(X, 2.0, loop=False, max_num_neighbors=10_000)
    pairs = set(map(tuple, e.T.tolist()))
    assert pairs and all((j, i) in pairs for i, j in pairs)
    deg = torch.bincount(e[1], minlength=300)
    K = int(deg.max()) // 2                                                # now the cap binds for the crowded particles
    capped = O.radius_graph(X, 2.0, loop=False, max_num_neighbors=K)
    cpairs = set(map(tuple, capped.T.tolist()))
    # radius_graph(loop=False) searches with K + 1 and drops the self pair afterwards (torch-cluster 1.6.3 radius_graph.py): a particle
    # whose first K + 1 hits in index order do not include itself -- many lower-indexed neighbours -- keeps K + 1 of them
    cdeg = torch.bincount(capped[1], minlength=300)
    assert cpairs < pairs and int(cdeg.max()) == K + 1 and int(cdeg[:5].max()) <= K
    assert any((j, i) not in cpairs for i, j in cpairs)                    # index-order truncation is not symmetric


def test_pair_distance_loss_is_invariant_with_zero_net_force():
    P = (_cloud(200, seed=5, span=0.05)).clone().requires_grad_(True)
    thr = 0.006
    loss = O.distance_loss(P, thr)
    assert float(loss.detach()) > 0
    loss.backward()
    assert float(P.grad.sum(0).abs().max()) < 1e-9 * float(P.grad.abs().sum())
    R, t = _rotation(6), torch.tensor([0.3, 0.1, -0.2], dtype=torch.float64)
    assert abs(float(O.distance_loss(P.detach() @ R.T + t, thr)) - float(loss)) < 1e-9 * float(loss)
    far = torch.tensor(np.mgrid[0:4, 0:4, 0:4].reshape(3, -1).T * 0.01, dtype=torch.float64)       # every pair >= 0.01 apart
    assert float(O.distance_loss(far, thr)) == 0.0


def test_visual_particles_follow_a_rigid_translation_of_the_estimate():
    prm = O.PBFParams()
    xyz = _cloud(500, seed=7, span=12.0)
    vis = xyz[:80] + 0.3
    shift = torch.tensor([0.02, 0.05, -0.01], dtype=torch.float64)         # scaled units
    est = (xyz + shift) / O.SCALE_FACTOR
    out = O.visual_xyz_from_nn(prm, est, xyz, vis)
    # every hidden particle has velocity shift / secs, so the poly6-weighted mean is exactly that and v' = v + secs * u = v + shift
    has_nb = ((vis[:, None, :] - (xyz + shift)[None]) ** 2).sum(-1).min(1).values < prm.H2
    assert bool(has_nb.all()) and torch.allclose(out, vis + shift, rtol=0, atol=1e-12)


def test_next_tick_map_is_affine_in_the_trainable_positions():
    prm = O.PBFParams(buoyancy_max_y=0.0)
    rng = np.random.default_rng(8)
    t = lambda *s: torch.tensor(rng.normal(size=s))
    xyz, b, f = t(50, 3) * 5, t(50, 3), t(50, 3)
    e1, e2 = t(50, 3) * 0.05, t(50, 3) * 0.05
    Y = lambda e: O.guess_hidden_particles_from_nn(prm, e, xyz, b, f)
    mid = Y(0.3 * e1 + 0.7 * e2)
    assert torch.allclose(mid, 0.3 * Y(e1) + 0.7 * Y(e2), rtol=1e-12, atol=1e-12)
    # Y = 2 X - xyz + secs^2 (b + F): one tick of the estimated velocity (gm_fluid.py:846-862)
    assert torch.allclose(Y(e1), 2 * e1 * O.SCALE_FACTOR - xyz + prm.secs ** 2 * (b + f), rtol=1e-12, atol=1e-12)
    # with the height term the buoyancy fades linearly with the UNSCALED height (gm_fluid.py:847-850): still affine in e
    prm_h = O.PBFParams(buoyancy_max_y=0.8)
    Yh = O.guess_hidden_particles_from_nn(prm_h, e1, xyz, b, f)
    assert torch.allclose(Yh, 2 * e1 * O.SCALE_FACTOR - xyz + prm.secs ** 2 * (b * (1 - e1[:, 1:2] / 0.8) + f), rtol=1e-12, atol=1e-12)
