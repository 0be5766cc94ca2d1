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


@pytest.mark.parametrize("tag", ["guess_plain", "guess_bmax_wind", "guess_stable"])
def test_guess_and_confirm_equal_the_reference(tag):
    """guess_hidden_particles / confirm_guess_hidden_particles were run by the reference in its own fp32 (they hard-wire
    torch.float and device="cuda", redirected to the CPU by the generator script); the fp64 oracle agrees to fp32 rounding."""
    wind = bool(Z[f"{tag}_wind"])
    sp = O.SolverParams(buoyancy_max_y=float(Z[f"{tag}_bmax"]), buoyancy_decay_rate=float(Z[f"{tag}_decay"]), alpha=-0.2,
                        wind_force=(0.3, 0.0, 0.1), wind_power=2.0)
    sp.wind_force_max = 0.3
    st = {k: _t(tag, k) for k in ("xyz", "velocity", "force", "buoyancy")}
    st["estimate_xyz"], st["counts"] = st["xyz"].clone(), torch.ones(st["xyz"].shape[0], 1, dtype=torch.float64)
    O.solver_guess_hidden_particles(sp, st, stable=bool(Z[f"{tag}_stable"]), use_wind=wind)
    for k in ("velocity", "force", "buoyancy", "estimate_xyz", "counts"):
        ref = Z[f"{tag}_after_{k}"]
        if np.abs(ref).max() == 0:
            assert float(st[k].abs().max()) == 0, k
        else:
            assert rel(st[k].numpy(), ref) < 2e-6, (k, rel(st[k].numpy(), ref))
    st["estimate_xyz"] = torch.tensor(Z[f"{tag}_moved"], dtype=torch.float64)
    st["xyz"] = _t(tag, "xyz")
    O.solver_confirm_guess_hidden_particles(sp if not bool(Z[f"{tag}_stable"]) else sp, st)
    assert rel(st["xyz"].numpy(), Z[f"{tag}_confirm_xyz"]) < 1e-7
    v_ref = Z[f"{tag}_confirm_velocity"]
    still = np.all(v_ref == 0, axis=1)
    assert still[::3].all() and not still[1::3].all()                       # the particles left in place, and only they, stop
    assert float(st["velocity"][torch.tensor(still)].abs().max()) == 0
    # (e - x)/secs in fp32 loses ~ |x| * eps / |e - x| relative accuracy: compare with the matching absolute tolerance
    err = np.abs(st["velocity"].numpy() - v_ref).max()
    assert err < 3e-5 * np.abs(Z[f"{tag}_xyz"]).max() / 0.033, err


def test_gradient_cache_and_adam_configuration_equal_the_reference():
    """P6 / P7 (gm_fluid.py:330-355, 401-430): per-view gradients summed, times 1/batch, torch.optim.Adam with lr =
    position_lr_init * spatial_lr_scale and eps = 1e-15; update_learning_rate_current never changes the rate.  This is what
    oracle/step_ref.py inlines and what the flat-bucket all-reduce + fnx_adam_step reproduce."""
    assert float(Z["plumb_lr"]) == pytest.approx(1.6e-4) and float(Z["plumb_lr_after"]) == pytest.approx(1.6e-4) and float(Z["plumb_eps"]) == 1e-15
    e = torch.tensor(Z["plumb_e0"]).clone().requires_grad_(True)
    opt = torch.optim.Adam([{"params": [e], "lr": 1.6e-4, "name": "estimate_xyz_nn"}], lr=0.0, eps=1e-15)     # as oracle/step_ref.py builds it
    vg = Z["plumb_view_grads"]
    for it in range(2):
        cache = torch.zeros_like(e)
        for v in range(3):
            cache += torch.tensor(vg[it, v])
        e.grad = cache * (1.0 / 3)
        assert np.array_equal(e.grad.numpy(), Z[f"plumb_batch_grad{it}"])
        opt.step()
        assert np.allclose(e.detach().numpy(), Z[f"plumb_e{it + 1}"], rtol=0, atol=1e-9)
    # the first Adam step moves every coordinate by ~lr against the sign of its gradient
    d = Z["plumb_e1"] - Z["plumb_e0"]
    assert np.allclose(np.abs(d), 1.6e-4, rtol=1e-3) and np.all(np.sign(d) == -np.sign(Z["plumb_batch_grad0"]))


def test_the_dynamics_model_class_shares_the_physics():
    """The FluidNexus scenes (smoke, ball) train gm_dynamics.GaussianModel, ScalarReal gm_fluid.GaussianModel; the generator ran
    P1-P3 of both classes on the same state and found them bit-identical, so one set of fixtures covers both."""
    assert bool(Z["gm_dynamics_bit_identical"])
