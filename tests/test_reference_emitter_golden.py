"""CPU: particle creation and emission (fluidnexus_b200/emitter.py) against the reference's OWN methods.

tests/golden/pyref_emitter.npz holds what gm_dynamics.GaussianModel.create_particles_visual / create_particles_hidden /
prepare_emitter_points / prepare_emitter_future_first_points / emit_new_particles (FD/gaussian_splatting/gm_dynamics.py:510-609,
674-788, 844-976) produced under fixed numpy / torch seeds (tools/make_emitter_golden.py, run where /root/reference exists).  The mirror
replays the same calls under the same seeds: same emitter sites in the same order, same random streams consumed by the same calls
=> the arrays must be IDENTICAL, not close."""
import os
import types

import numpy as np
import pytest
import torch

from fluidnexus_b200.emitter import EmitterMixin, disc_sites

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "pyref_emitter.npz"))
STATE = ["_xyz", "_estimate_xyz", "_buoyancy", "_force", "_velocity", "_imass", "_counts", "_particle_id", "_visual_xyz"]


class _Particles(EmitterMixin):
    """The state attributes the mixin works on, on the CPU (the solver class holds the same on the GPU)."""

    def __init__(self, alpha):
        self.dev, self.scale_factor, self.alpha = torch.device("cpu"), 100.0, float(alpha)
        self._gravity = (0.0, -9.8, 0.0)
        self._visual_xyz = torch.zeros((0, 3))


def _args(case, group):
    prefix = f"{case}/{group}/"
    return types.SimpleNamespace(**{k[len(prefix):]: G[k].item() for k in G.files if k.startswith(prefix)})


def _same(p, case, tag):
    for k in STATE:
        want = G[f"{case}/{tag}{k}"]
        got = getattr(p, k).numpy()
        assert got.shape == want.shape and got.dtype == want.dtype, (tag, k, got.shape, want.shape, got.dtype, want.dtype)
        assert np.array_equal(got, want), (tag, k, float(np.abs(got - want).max()))


@pytest.mark.parametrize("case", ["default", "busy"])
def test_creation_and_emission_reproduce_the_reference_particle_for_particle(case):
    margs, oargs = _args(case, "model"), _args(case, "optim")
    seed = int(G[f"{case}/seed"])
    np.random.seed(seed)
    torch.manual_seed(seed)
    p = _Particles(oargs.alpha)
    p.setup_emitter(oargs)
    p.create_particles_visual(margs)
    assert np.array_equal(p._visual_xyz.numpy(), G[f"{case}/visual_created"])
    p._visual_xyz = p._visual_xyz * p.scale_factor                  # detach_visual_and_scale (gm_dynamics.py:506-507)
    p.create_particles_hidden(margs)
    p.prepare_emitter_points(margs, is_future=bool(G[f"{case}/is_future"]))
    p.prepare_emitter_future_first_points(margs)
    for k in ("visual_emitter_points", "hidden_emitter_points", "visual_emitter_first_points", "hidden_emitter_first_points"):
        assert np.array_equal(getattr(p, k).numpy(), G[f"{case}/{k}"]), k
    _same(p, case, "created")
    for it in range(3):
        p.emit_new_particles()
        _same(p, case, f"emit{it}")
    p.emit_new_particles(future_time_index=0)
    _same(p, case, "future0")
    assert p.emit_counter == int(G[f"{case}/emit_counter"]) == 4


def test_emission_bookkeeping():
    """Closed forms: a ratio of r emits int(r) whole copies of the site set plus int(frac * sites) random sites; every emission
    resets the solver counts of all particles; new particles start with buoyancy gravity * alpha, unit mass, the initial velocity."""
    margs = _args("default", "model")
    p = _Particles(alpha=-0.2)
    p.setup_emitter(emit_ratio_hidden=2.25, emit_ratio_visual=0.0, init_hidden_velocity=1.5)
    p.create_particles_hidden(margs)
    p.prepare_emitter_points(margs)
    n0, sites = p._xyz.shape[0], p.hidden_emitter_points.shape[0]
    p._counts += 3.0
    p.emit_new_particles()
    n_new = 2 * sites + int(0.25 * sites)
    assert p._xyz.shape[0] == n0 + n_new and p._visual_xyz.shape[0] == 0            # ratio 0: nothing visual
    assert not p._counts.any() and p._counts.shape == (n0 + n_new, 1)
    new = slice(n0, None)
    assert torch.equal(p._xyz[n0:n0 + sites], p.hidden_emitter_points * 100.0)       # whole copies sit exactly on the sites
    assert torch.allclose(p._buoyancy[new], torch.tensor([[0.0, 1.96, 0.0]]).expand(n_new, 3))
    assert bool((p._velocity[new, 1] == 1.5).all()) and not p._velocity[new][:, [0, 2]].any() and bool((p._imass[new] == 1).all())
    assert torch.equal(p._particle_id[:, 0], torch.arange(n0 + n_new)) and p._particle_id_max == n0 + n_new
    # sites: a disc of lattice points, x-major, inside the radius
    s = disc_sites(0.3, -0.2, 0.05, 0.01, [0.0, 0.5])
    assert s.shape[1] == 3 and bool(((s[:, 0] - 0.3) ** 2 + (s[:, 2] + 0.2) ** 2 <= 0.05 ** 2 + 1e-15).all())
    assert bool((np.diff(s[:, 0]) >= 0).all()) and set(np.unique(s[:, 1])) == {0.0, 0.5}


def test_solver_class_carries_the_emitter():
    from fluidnexus_b200.solver import PBFSolver
    for name in ("create_particles_visual", "create_particles_hidden", "prepare_emitter_points", "prepare_emitter_future_first_points",
                 "emit_new_particles"):
        assert getattr(PBFSolver, name) is getattr(EmitterMixin, name), name
