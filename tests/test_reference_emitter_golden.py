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
    # constant raw appearance the render pipes read for these sets (gm_dynamics.py:1636-1700)
    p.prepare_hidden_particles_for_rendering()
    p.prepare_visual_particles_for_rendering()
    for k in ("_color_dummy", "_scales_dummy", "_rotation_dummy", "_opacity_dummy", "_visual_color", "_visual_scales", "_visual_rotation",
              "_visual_opacity"):
        assert np.array_equal(getattr(p, k).numpy(), G[f"{case}/render{k}"]), k
    p._visual_color = p._visual_color * 0.5 + 0.1
    p.emit_new_particles()
    p.prepare_future_visual_particles_for_rendering(True)               # earlier particles keep theirs, new ones get the constants
    for k in ("_visual_color", "_visual_scales", "_visual_rotation", "_visual_opacity"):
        assert np.array_equal(getattr(p, k).numpy(), G[f"{case}/render_future{k}"]), k
    assert p._visual_color.shape[0] == p._visual_xyz.shape[0]


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


@pytest.mark.parametrize("kind", ["cuboid", "sphere", "cylinder"])
def test_rigid_body_samples_equal_the_reference(kind):
    """create_rigid_body (gm_dynamics.py:612-672): the same surface samples, in the same order, as the reference's own method."""
    np.random.seed(21)
    p = _Particles(alpha=-0.2)
    p.rigid_body, p.rigid_particle_diameter = kind, 2 * 0.25
    prefix = f"rigid/{kind}/"
    for k in G.files:
        if k.startswith(prefix) and k[len(prefix):] not in ("xyz", "imass") and not k[len(prefix):].startswith("render"):
            v = G[k]
            setattr(p, k[len(prefix):], v.tolist() if v.ndim else v.item())
    p.rigid_body_center = torch.tensor([0.34, 0.3, -0.225]) * 100.0
    p.create_rigid_body()
    assert np.array_equal(p._rigid_xyz.numpy(), G[prefix + "xyz"]) and np.array_equal(p._rigid_imass.numpy(), G[prefix + "imass"])
    p.prepare_rigid_body_particles_for_rendering()
    for k in ("_rigid_color", "_rigid_scales", "_rigid_rotation", "_rigid_opacity"):
        assert np.array_equal(getattr(p, k).numpy(), G[prefix + "render" + k]), k
    if kind == "cuboid":                                   # 5 x 4 x 6 lattice without its 3 x 2 x 4 interior
        assert p._rigid_xyz.shape[0] == 5 * 4 * 6 - 3 * 2 * 4
    with pytest.raises(ValueError):
        p.rigid_body = "torus"
        p.create_rigid_body()


def test_solver_class_carries_the_emitter():
    from fluidnexus_b200.solver import PBFSolver
    for name in ("create_particles_visual", "create_particles_hidden", "prepare_emitter_points", "prepare_emitter_future_first_points",
                 "emit_new_particles", "create_rigid_body"):
        assert getattr(PBFSolver, name) is getattr(EmitterMixin, name), name


@pytest.mark.parametrize("kind", ["cuboid", "sphere", "cylinder"])
def test_setup_rigid_body_registers_the_sampled_body(kind):
    """PBFSolver.setup_rigid_body = the rigid block of setup_constants + create_rigid_body + set_rigid_body: the half extents it hands
    to fnx_rigid_project are those of the reference's check_inside_rigid_body (gm_dynamics.py:1138-1170).  The solver object is made
    without its (CUDA-only) constructor; nothing here launches a kernel."""
    from fluidnexus_b200.solver import PBFSolver
    sol = object.__new__(PBFSolver)
    sol.dev, sol.scale_factor = torch.device("cpu"), 100.0
    oargs = types.SimpleNamespace(rigid_body=kind, rigid_body_center=[0.34, 0.5, -0.225], rigid_particle_radius=0.25, rigid_cuboid_num=[5, 10, 55],
                                  rigid_sphere_radius=5, rigid_sphere_num=1000, rigid_cylinder_radius=4, rigid_cylinder_num=[50, 50])
    np.random.seed(3)
    sol.setup_rigid_body(oargs)
    n = {"cuboid": 5 * 10 * 55 - 3 * 8 * 53, "sphere": 1000, "cylinder": 50 * 50}[kind]
    assert sol._rigid_xyz.shape == (n, 3) and sol._rigid_imass.shape == (n, 1) and sol.rigid_body == kind
    assert np.allclose(list(sol._rigid_center), [34.0, 50.0, -22.5])
    want = {"cuboid": [5 * 0.5 / 2, 10 * 0.5 / 2, 55 * 0.5 / 2], "sphere": [5.0, 0.0, 0.0], "cylinder": [4.0, 50 * 0.5 / 2, 0.0]}[kind]
    assert np.allclose(list(sol._rigid_prm), want)
    # every sample lies on the body registered for the inside test (cuboid: lattice offset by -n//2, so within one pitch of the box)
    d = (sol._rigid_xyz - torch.tensor([34.0, 50.0, -22.5])).abs()
    if kind == "sphere":
        assert torch.allclose(d.norm(dim=1), torch.full((n,), 5.0), atol=1e-4)
    elif kind == "cylinder":
        assert torch.allclose(d[:, :2].norm(dim=1), torch.full((n,), 4.0), atol=1e-4) and float(d[:, 2].max()) <= 50 * 0.5 / 2 + 1e-4
    else:
        assert bool((d <= torch.tensor(want) + 0.5 + 1e-4).all())


@pytest.mark.parametrize("case", ["inherit", "dist_scales"])
def test_level_two_initial_quantities_equal_the_reference(case):
    """init_quantities_current_level_two (gm_dynamics.py:363-378): inherited rows and distance-based log-scales exactly as the
    reference's own method produces them (distCUDA2 stood in for by the oracle's kNN on both sides)."""
    from fluidnexus_b200.level_two import init_quantities_current_level_two
    from oracle import pbf_ref as O
    pre = f"l2init/{case}/"
    t = lambda k: torch.from_numpy(G[pre + k])
    flags = {k[len(pre) + 5:]: bool(G[k]) for k in G.files if k.startswith(pre + "flag/")}
    prev = {k: t("prev_" + k) for k in ("color", "opacity", "scales", "rotation")}
    got = init_quantities_current_level_two(t("in_visual_xyz"), t("in_visual_color"), t("in_visual_opacity"), t("in_visual_scales"),
                                            t("in_visual_rotation"), prev=prev, dist2_fn=lambda p: torch.tensor(O.knn3_mean_dist2(p.numpy())), **flags)
    for name, a in zip(("_visual_color", "_visual_opacity", "_visual_scales", "_visual_rotation"), got):
        want = G[pre + "out" + name]
        assert a.numpy().shape == want.shape and np.array_equal(a.numpy(), want), name
    # inputs are not modified; without `prev` nothing is inherited
    assert np.array_equal(t("in_visual_color").numpy(), G[pre + "in_visual_color"])
    first = init_quantities_current_level_two(t("in_visual_xyz"), t("in_visual_color"), t("in_visual_opacity"), t("in_visual_scales"),
                                              t("in_visual_rotation"), prev=None, **{**flags, "init_scales_w_xyz_dist": False})
    for name, a in zip(("color", "opacity", "scales", "rotation"), first):
        assert torch.equal(a, t("in_visual_" + name)), name


def test_quantity_snapshots_equal_the_reference_writers(tmp_path):
    """save_particles_* (gm_dynamics.py:1938-2017): same file names, shapes, dtypes and contents as the reference's own methods."""
    from fluidnexus_b200 import io as IO
    from fluidnexus_b200.solver import PBFSolver
    t = lambda k: torch.from_numpy(G["quant/in" + k])
    sol = object.__new__(PBFSolver)
    sol.dev, sol.scale_factor = torch.device("cpu"), 100.0
    sol._xyz, sol._estimate_xyz, sol._visual_xyz, sol._rigid_xyz = t("_xyz"), t("_estimate_xyz"), t("_visual_xyz"), t("_rigid_xyz")
    d = str(tmp_path)
    sol.save_particles_rigid_body(d, 3)
    sol.save_particles_frame(d, 3)
    sol.save_particles_simulation(d, 41)
    sol.save_particles_simulation_guess(d, 41)
    IO.save_particles("optimization_first", d, dict(visual_xyz=t("_visual_xyz")), 0, 250)
    IO.save_particles("optimization", d, dict(estimate_xyz_nn=t("_estimate_xyz_nn"), visual_xyz=t("_advected")), 7, 120)
    IO.save_particles("optimization_level_two", d, {k: t("_" + k) for k in IO.VISUAL_ARRAYS[1:]}, 7, 30)
    assert sorted(os.listdir(d)) == [str(f) for f in G["quant/files"]]
    for f in G["quant/files"]:
        a, b = np.load(os.path.join(d, str(f))), G["quant/file/" + str(f)]
        assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b), f
    # an empty visual set writes no visual file (gm_dynamics.py:1953, 1970)
    sol._visual_xyz = torch.zeros((0, 3))
    assert [os.path.basename(p) for p in sol.save_particles_frame(str(tmp_path / "e"), 1)] == ["frame_001_xyz.npy"]


def test_particle_state_renders_through_the_render_glue_mirror():
    """The solver-side state (emitter.py accessors + load_ply) is what renderer.render_dynamics / render_fluid read: with the
    recording rasterizer of the render-glue golden, the particles arrive in render units with their constant appearance activated
    (sigmoid / exp / normalise), grey repeated to RGB, followed by the frozen background set of the reference-written point cloud."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from fluidnexus_b200 import renderer as RD
    from fluidnexus_b200 import synthetic as S
    from make_render_glue_golden import run_case
    margs = _args("default", "model")
    np.random.seed(0)
    p = _Particles(alpha=-0.2)
    p.create_particles_visual(margs)
    p._visual_xyz = p._visual_xyz * p.scale_factor
    p.create_particles_hidden(margs)
    p.prepare_visual_particles_for_rendering()
    p.prepare_hidden_particles_for_rendering()
    n_bg = p.load_ply(os.path.join(os.path.dirname(__file__), "golden", "pyref_background_ply.bin"))
    V = p._visual_xyz.shape[0]
    assert n_bg > 0 and p.get_gs_xyz.shape == (n_bg, 3) and p.get_gs_color.shape[1] == 3
    cam = S.make_cameras(5, 32, height=24)[1]
    rec = run_case(RD.render_dynamics, p, cam, dict(pos_type="visual", scale=True))
    assert rec["in_means3D"].shape == (V + n_bg, 3)
    assert np.allclose(rec["in_means3D"][:V], p._visual_xyz.numpy() / 100.0, rtol=1e-6) and np.array_equal(rec["in_means3D"][V:], p._gs_xyz.numpy())
    assert np.allclose(rec["in_opacities"][:V], 0.1, atol=1e-6) and np.allclose(rec["in_scales"][:V], np.exp(-5.9), rtol=1e-6)
    assert np.allclose(rec["in_colors_precomp"][:V], 0.7) and rec["in_colors_precomp"].shape == (V + n_bg, 3)
    assert np.array_equal(rec["in_rotations"][:V], np.tile([1.0, 0, 0, 0], (V, 1)).astype(np.float32))
    assert np.allclose(np.linalg.norm(rec["in_rotations"][V:], axis=1), 1.0, atol=1e-5)                 # normalised background quaternions
    assert np.allclose(rec["in_opacities"][V:, 0], 1 / (1 + np.exp(-p._gs_opacity.numpy()[:, 0])), rtol=1e-5)
    hid = run_case(RD.render_fluid, p, cam, dict(pos_type="hidden", scale=True))
    assert hid["in_means3D"].shape == (p._xyz.shape[0], 3) and np.allclose(hid["in_means3D"], p._xyz.numpy() / 100.0, rtol=1e-6)
    assert hid["in_colors_precomp"].shape[1] == 1 and np.allclose(hid["in_opacities"], 0.1, atol=1e-6)
