"""CPU: oracle/pbf_ref.py against the outputs of the REFERENCE'S OWN physics methods.  tools/make_physics_golden.py ran
gm_fluid.GaussianModel.get_visual_xyz_from_nn / get_gas_constraints_from_exyz_nn / get_guess_hidden_particles_from_nn /
get_gas_constraints_from_vel_nn_guess / project_gas_constraints / update_visual_particles / remove_invalid_particles from
/root/reference on the CPU in fp64, with torch_cluster / torch_scatter replaced by the oracle's radius / radius_graph /
scatter_min (the one part that cannot be pinned here).  Everything around the neighbour search -- kernels, row / col
conventions, scatter sums, the autograd chain, the solver update, the pruning rule -- is therefore checked against the
reference code itself, not against a reading of it."""
import os

import numpy as np
import pytest
import torch

from oracle import pbf_ref as O

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "pyref_physics.npz"))
TAGS = ["smoke", "scalar", "capped"]


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def _t(tag, k):
    return torch.tensor(Z[f"{tag}_{k}"], dtype=torch.float64)


@pytest.mark.parametrize("tag", TAGS)
def test_differentiable_terms_and_their_gradient(tag):
    prm = O.PBFParams(KNN_K=int(Z[f"{tag}_K"]), p0=float(Z[f"{tag}_p0"]), buoyancy_max_y=float(Z[f"{tag}_bmax"]))
    xyz, est, buoy, force, imass, vis = (_t(tag, k) for k in ("xyz", "estimate_xyz", "buoyancy", "force", "imass", "visual_xyz"))
    e = (est / 100.0).clone().requires_grad_(True)
    p1 = O.visual_xyz_from_nn(prm, e, xyz, vis)
    p2 = O.gas_constraints_from_exyz_nn(prm, e, imass)
    Y = O.guess_hidden_particles_from_nn(prm, e, xyz, buoy, force)
    p3 = O.gas_constraints_from_vel_nn_guess(prm, e, xyz, buoy, force, imass)
    loss = (p1 * _t(tag, "w_vis")).sum() + ((p2 - 1.0) ** 2).mean() + 0.1 * ((p3 - 1.0) ** 2).mean()
    loss.backward()
    assert rel(p1.detach().numpy(), Z[f"{tag}_P1"]) < 1e-12
    assert rel(p2.detach().numpy(), Z[f"{tag}_P2"]) < 1e-12
    assert rel(Y.detach().numpy(), Z[f"{tag}_Y"]) < 1e-13
    assert rel(p3.detach().numpy(), Z[f"{tag}_P3"]) < 1e-12
    assert abs(loss.item() - float(Z[f"{tag}_loss"])) < 1e-12 * abs(float(Z[f"{tag}_loss"]))
    assert rel(e.grad.numpy(), Z[f"{tag}_grad"]) < 1e-10


@pytest.mark.parametrize("tag", TAGS)
def test_solver_iteration_visual_update_and_pruning(tag):
    sp = O.SolverParams(KNN_K=int(Z[f"{tag}_K"]), p0=float(Z[f"{tag}_p0"]), k=3.0, buoyancy_max_y=float(Z[f"{tag}_bmax"]), min_neighbors=20)
    st = {k: _t(tag, k) for k in ("xyz", "estimate_xyz", "buoyancy", "force", "velocity", "imass", "counts", "visual_xyz")}
    p_ratio, lambdas = O.solver_project_gas_constraints(sp, st)
    assert rel(st["estimate_xyz"].numpy(), Z[f"{tag}_proj_estimate_xyz"]) < 1e-12
    assert rel(st["force"].numpy(), Z[f"{tag}_proj_force"]) < 1e-10
    assert abs(float(lambdas.mean()) - float(Z[f"{tag}_proj_lambda_mean"])) < 1e-9 * abs(float(Z[f"{tag}_proj_lambda_mean"]))
    assert abs(float(p_ratio.mean()) - float(Z[f"{tag}_proj_pratio_mean"])) < 1e-9 * abs(float(Z[f"{tag}_proj_pratio_mean"]))
    O.solver_update_visual_particles(sp, st)
    assert rel(st["visual_xyz"].numpy(), Z[f"{tag}_visual_after_update"]) < 1e-12
    deg = O.solver_neighbor_degree(sp, st["xyz"])
    assert int((deg >= sp.min_neighbors).sum()) == int(Z[f"{tag}_kept_after_prune"]) and st["xyz"].shape[0] == int(Z[f"{tag}_n0"])
