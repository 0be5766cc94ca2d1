"""CPU: the solver-tick oracle (oracle/pbf_ref.py:solver_*) against invariants of the reference algorithm, and the
no-fallback rule of the product class."""
import numpy as np
import pytest
import torch

from fluidnexus_b200 import synthetic as S
from oracle import pbf_ref as O


def _state(N=400, V=60, seed=0):
    hp = S.hidden_lattice(N, seed=seed)
    rng = np.random.default_rng(seed)
    f = lambda a: torch.tensor(a, dtype=torch.float64)
    return dict(xyz=f(hp.xyz), estimate_xyz=f(hp.xyz), velocity=f(rng.normal(0, 5, hp.xyz.shape)), force=torch.zeros(hp.N, 3, dtype=torch.float64),
                buoyancy=torch.zeros(hp.N, 3, dtype=torch.float64), imass=torch.ones(hp.N, 1, dtype=torch.float64),
                counts=torch.zeros(hp.N, 1, dtype=torch.float64), visual_xyz=f(hp.xyz[:V] + 0.2))


def test_guess_then_confirm_without_solver_is_plain_euler():
    """guess_hidden_particles + confirm_guess_hidden_particles with no solver iteration: x += secs * v', v' = v + g*alpha*secs
    (gm_fluid.py:809-844, 1160-1175), forces cleared, counts cleared."""
    sp, st = O.SolverParams(alpha=-0.2), _state()
    x0, v0 = st["xyz"].clone(), st["velocity"].clone()
    st["force"] += 1.0
    O.solver_guess_hidden_particles(sp, st)
    v1 = v0 + torch.tensor([0.0, -9.8 * -0.2, 0.0]) * sp.secs + sp.secs * 1.0
    assert torch.allclose(st["velocity"], v1) and float(st["force"].abs().max()) == 0 and float(st["counts"].abs().max()) == 0
    O.solver_confirm_guess_hidden_particles(sp, st)
    assert torch.allclose(st["xyz"], x0 + sp.secs * v1) and torch.allclose(st["velocity"], v1)


def test_projection_moves_momentum_free_when_cap_does_not_bind():
    """With a symmetric neighbour graph (K not binding) and equal neighbour counts the pairwise corrections
    (lambda_i + lambda_j + s_corr) * spiky are antisymmetric: two isolated particles move by equal and opposite amounts."""
    sp = O.SolverParams()
    st = _state(N=2)
    st["xyz"] = torch.tensor([[0.0, 0.0, 0.0], [0.9, 0.2, -0.1]], dtype=torch.float64)
    st["estimate_xyz"] = st["xyz"].clone()
    st["velocity"], st["force"], st["buoyancy"] = (torch.zeros(2, 3, dtype=torch.float64) for _ in range(3))
    st["imass"], st["counts"] = torch.ones(2, 1, dtype=torch.float64), torch.zeros(2, 1, dtype=torch.float64)
    before = st["estimate_xyz"].clone()
    O.solver_project_gas_constraints(sp, st)
    d = st["estimate_xyz"] - before
    assert torch.allclose(d[0], -d[1], atol=1e-12) and float(d.abs().max()) > 0


def test_update_visual_moves_with_the_local_velocity():
    sp, st = O.SolverParams(), _state()
    st["velocity"] = torch.ones_like(st["velocity"]) * torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)   # uniform flow
    v0 = st["visual_xyz"].clone()
    O.solver_update_visual_particles(sp, st)
    assert torch.allclose(st["visual_xyz"] - v0, torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64) * sp.secs)


def test_neighbor_degree_is_capped_like_torch_cluster():
    sp = O.SolverParams(H=500.0)                     # every particle sees every other: the default cap of 32 binds
    xyz = torch.tensor(S.hidden_lattice(60, seed=1).xyz, dtype=torch.float64)
    deg = O.solver_neighbor_degree(sp, xyz)
    # a query keeps its first 33 hits in index order (self included, then dropped): low indices are everybody's neighbour
    assert int(deg[0]) == 59 and int(deg[-1]) < 32 and int(deg.sum()) <= 60 * 33


def test_solver_class_refuses_cpu():
    from fluidnexus_b200.solver import PBFSolver
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PBFSolver(np.zeros((4, 3), np.float32), device="cpu")


def test_predict_runs_the_future_simulation_loop_in_the_reference_order():
    """PBFSolver.predict = the per-frame sequence of future_simulation.py:118-175 (kernel-calling methods replaced by recorders: no
    GPU here), with the p0 decay schedule of :120 and wind from `wind_since` on."""
    import torch
    from fluidnexus_b200.solver import PBFSolver
    sol = object.__new__(PBFSolver)
    sol.dev, sol.scale_factor, sol.p0 = torch.device("cpu"), 100.0, 2.0
    sol._visual_xyz = torch.tensor([[0.0, -5.0, 0.0], [0.0, 3.0, 0.0], [1.0, -1.7, 0.0]])      # y = -5 lies below -0.017 * 100
    sol._visual_color = torch.zeros((0, 1))
    log = []
    for name in ("remove_invalid_particles", "emit_new_particles", "project_gas_constraints", "confirm_guess_hidden_particles",
                 "update_visual_particles"):
        setattr(sol, name, (lambda n: lambda *a, **k: log.append(n))(name))
    sol.guess_hidden_particles = lambda stable=False, use_wind=False: log.append(("guess", use_wind))
    sol.prepare_future_visual_particles_for_rendering = lambda lvl2=False: log.append(("prepare", lvl2))
    frames = []
    sol.predict(3, first_frame_index=120, solver_iterations_future=2, p0_future=1.5, decay_frames_future_p0=4, wind_since=121,
                use_level_two_in_future=True, on_frame=lambda s, f: frames.append((f, s.p0, s._visual_xyz.shape[0])))
    per_frame = ["remove_invalid_particles", "emit_new_particles", ("guess", None), "project_gas_constraints", "project_gas_constraints",
                 "confirm_guess_hidden_particles", "update_visual_particles", ("prepare", True)]
    assert len(log) == 3 * len(per_frame)
    for t in range(3):
        got = log[t * len(per_frame):(t + 1) * len(per_frame)]
        want = [("guess", 120 + t >= 121) if x == ("guess", None) else x for x in per_frame]
        assert got == want, (t, got)
    # p0: 2.0 -> 1.5 linearly over 4 frames; the bottom particle is dropped before the first frame only
    assert [f for f, _, _ in frames] == [120, 121, 122]
    assert [round(p, 6) for _, p, _ in frames] == [2.0, 1.875, 1.75]
    assert [n for _, _, n in frames] == [2, 2, 2]
    assert PBFSolver.future_p0(2.0, 1.5, 10, 4) == 1.5


def test_solver_hands_its_state_to_the_fused_step_and_takes_the_result_back():
    """The seams between the solver tick and a frame's optimisation (train_physical_particle.py:300-301, 432-434): the solver exposes
    what step.FrameState reads, and confirm_guess_hidden_particles_from_nn takes the optimised tensor back in scaled units."""
    import torch
    from fluidnexus_b200.solver import PBFSolver
    sol = object.__new__(PBFSolver)
    sol.dev, sol.scale_factor = torch.device("cpu"), 100.0
    for k, cols in (("_xyz", 3), ("_estimate_xyz", 3), ("_buoyancy", 3), ("_force", 3), ("_imass", 1)):
        setattr(sol, k, torch.rand(7, cols))
    assert sol.N == 7 and sol.xyz is sol._xyz and sol.estimate_xyz is sol._estimate_xyz and sol.buoyancy is sol._buoyancy
    assert sol.force is sol._force and sol.imass is sol._imass
    e = (sol._estimate_xyz / 100.0 + 0.001).requires_grad_(True)          # FrameState.e after some Adam steps
    sol.confirm_guess_hidden_particles_from_nn(e)
    assert torch.allclose(sol._estimate_xyz, e.detach() * 100.0) and not sol._estimate_xyz.requires_grad
    assert PBFSolver.confirm_guess_hidden_particles_wo_velocity is PBFSolver.confirm_guess_hidden_particles
