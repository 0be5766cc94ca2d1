"""Run in the build container (needs /root/reference): executes the reference's OWN particle creation / emission methods
(FluidDynamics/gaussian_splatting/gm_dynamics.py: create_particles_visual, create_particles_hidden, prepare_emitter_points,
prepare_emitter_future_first_points, emit_new_particles) on the CPU under fixed numpy / torch seeds and stores the settings, the
emitter sites and the particle state after every call in tests/golden/pyref_emitter.npz.  tests/test_reference_emitter_golden.py
replays the same calls through fluidnexus_b200/emitter.py under the same seeds and requires identical arrays.

The model object is created without its constructor (GPU-only); `model_args` / `optim_args` are the reference's own
arguments.ModelParams / OptimizationParams defaults (overridden per case below).  device="cuda" / .cuda() land on the CPU
(tools/make_physics_golden.py:cuda_as_cpu)."""
import contextlib
import io
import os
import sys
from argparse import ArgumentParser

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/FluidDynamics"
OUT = os.path.join(ROOT, "tests", "golden", "pyref_emitter.npz")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_physics_golden import cuda_as_cpu, install_stubs  # noqa: E402

MODEL_KEYS = ["init_visual_num_pts", "init_thick_visual_num_pts", "init_visual_radius_small_max", "init_visual_radius_max", "init_x_mid",
              "init_visual_y_min", "init_visual_y_max", "init_z_mid", "init_visual_y_thick_min", "init_hidden_radius_max", "init_hidden_delta",
              "init_hidden_y_min", "init_hidden_y_max", "emitter_hidden_delta", "emitter_visual_delta", "emitter_center_y_hidden",
              "emitter_center_y_visual", "emitter_center_y_hidden_max", "emitter_center_y_visual_max", "emitter_visual_radius_ratio",
              "emitter_hidden_radius_ratio"]
OPTIM_KEYS = ["emit_ratio_hidden", "emit_ratio_visual", "extra_visual_ratio", "extra_visual_num", "extra_visual_y_min", "extra_visual_min_num",
              "init_hidden_velocity", "alpha"]
CASES = {
    # the reference's defaults (arguments/__init__.py): emit ratio 1.32 (one whole copy + a random 32 %), no extra particles
    "default": dict(seed=7, is_future=False, optim={"alpha": -0.2}, model={}),
    # everything switched on: several whole copies, a fraction below one, both kinds of extra visual particles, an initial velocity
    "busy": dict(seed=11, is_future=True, model={"init_thick_visual_num_pts": 0, "init_visual_num_pts": 400},
                 optim={"alpha": -0.35, "emit_ratio_hidden": 2.5, "emit_ratio_visual": 0.4, "extra_visual_ratio": 0.1, "extra_visual_min_num": 5,
                        "extra_visual_num": 7, "extra_visual_y_min": 0.05, "init_hidden_velocity": 3.0}),
}
STATE = ["_xyz", "_estimate_xyz", "_buoyancy", "_force", "_velocity", "_imass", "_counts", "_particle_id", "_visual_xyz"]


def snapshot(out, tag, gm):
    for k in STATE:
        out[f"{tag}{k}"] = getattr(gm, k).detach().cpu().numpy().copy()


def main():
    install_stubs()
    import types
    ply = types.ModuleType("plyfile")          # imported at gm_dynamics.py:7 for load_ply only; not installed here, not used by these methods
    ply.PlyData, ply.PlyElement = object, object
    sys.modules.setdefault("plyfile", ply)
    sys.path.insert(0, REF)
    import arguments
    from gaussian_splatting.gm_dynamics import GaussianModel as GM
    out = {}
    for case, spec in CASES.items():
        margs, oargs = arguments.ModelParams(ArgumentParser()), arguments.OptimizationParams(ArgumentParser())
        for k, v in spec["model"].items():
            setattr(margs, k, v)
        for k, v in spec["optim"].items():
            setattr(oargs, k, v)
        for k in MODEL_KEYS:
            out[f"{case}/model/{k}"] = getattr(margs, k)
        for k in OPTIM_KEYS:
            out[f"{case}/optim/{k}"] = getattr(oargs, k)
        out[f"{case}/seed"], out[f"{case}/is_future"] = spec["seed"], spec["is_future"]
        np.random.seed(spec["seed"])
        torch.manual_seed(spec["seed"])
        gm = object.__new__(GM)
        with cuda_as_cpu(), contextlib.redirect_stdout(io.StringIO()):
            # what setup_constants (gm_dynamics.py:76-160) would set, for the attributes these methods read
            gm._gravity = torch.tensor([0.0, -9.8, 0.0], dtype=torch.float, device="cuda").reshape((1, 3))
            gm.scale_factor, gm.emit_counter = 100.0, 0
            for k in OPTIM_KEYS:
                setattr(gm, k, getattr(oargs, k))
            gm.create_particles_visual(margs)
            out[f"{case}/visual_created"] = gm._visual_xyz.numpy().copy()
            gm.detach_visual_and_scale()                   # the entries scale the visual particles before the first emission
            gm.create_particles_hidden(margs)
            gm.prepare_emitter_points(margs, is_future=spec["is_future"])
            gm.prepare_emitter_future_first_points(margs)
            for k in ("visual_emitter_points", "hidden_emitter_points", "visual_emitter_first_points", "hidden_emitter_first_points"):
                out[f"{case}/{k}"] = getattr(gm, k).numpy().copy()
            snapshot(out, f"{case}/created", gm)
            for it in range(3):
                gm.emit_new_particles()
                snapshot(out, f"{case}/emit{it}", gm)
            gm.emit_new_particles(future_time_index=0)
            snapshot(out, f"{case}/future0", gm)
            out[f"{case}/emit_counter"] = gm.emit_counter
            # ---- constant raw appearance for rendering (gm_dynamics.py:1636-1700) ----
            gm.constant_color, gm.constant_scale, gm.constant_opacity = 0.7, -5.9, 0.1        # setup_constants :158-160
            gm.prepare_hidden_particles_for_rendering()
            gm.prepare_visual_particles_for_rendering()
            for k in ("_color_dummy", "_scales_dummy", "_rotation_dummy", "_opacity_dummy", "_visual_color", "_visual_scales", "_visual_rotation",
                      "_visual_opacity"):
                out[f"{case}/render{k}"] = getattr(gm, k).numpy().copy()
            gm._visual_color = gm._visual_color * 0.5 + 0.1          # stand-in for a level-two result on the existing particles
            gm.emit_new_particles()
            gm.prepare_future_visual_particles_for_rendering(True)
            for k in ("_visual_color", "_visual_scales", "_visual_rotation", "_visual_opacity"):
                out[f"{case}/render_future{k}"] = getattr(gm, k).numpy().copy()
        print(case, "visual", out[f"{case}/visual_created"].shape, "hidden", out[f"{case}/created_xyz"].shape, "sites",
              out[f"{case}/visual_emitter_points"].shape, out[f"{case}/hidden_emitter_points"].shape, "after 3 ticks",
              out[f"{case}/emit2_xyz"].shape, out[f"{case}/emit2_visual_xyz"].shape, "future", out[f"{case}/future0_xyz"].shape)
    # ---- create_rigid_body (gm_dynamics.py:612-672) for the three body kinds ----
    for kind, attrs in {"cuboid": dict(rigid_cuboid_num=[5, 4, 6]), "sphere": dict(rigid_sphere_num=300, rigid_sphere_radius=6.5),
                        "cylinder": dict(rigid_cylinder_radius=4.0, rigid_cylinder_num=[24, 9])}.items():
        np.random.seed(21)
        gm = object.__new__(GM)
        gm.rigid_body, gm.rigid_particle_diameter = kind, 2 * 0.25
        for k, v in attrs.items():
            setattr(gm, k, v)
            out[f"rigid/{kind}/{k}"] = np.asarray(v)
        with cuda_as_cpu():
            gm.rigid_body_center = torch.tensor([0.34, 0.3, -0.225], dtype=torch.float, device="cuda") * 100.0
            gm.create_rigid_body()
        out[f"rigid/{kind}/xyz"], out[f"rigid/{kind}/imass"] = gm._rigid_xyz.numpy().copy(), gm._rigid_imass.numpy().copy()
        with cuda_as_cpu():
            gm.prepare_rigid_body_particles_for_rendering()
        for k in ("_rigid_color", "_rigid_scales", "_rigid_rotation", "_rigid_opacity"):
            out[f"rigid/{kind}/render{k}"] = getattr(gm, k).numpy().copy()
        print("rigid", kind, out[f"rigid/{kind}/xyz"].shape)
    # ---- init_quantities_current_level_two (gm_dynamics.py:363-378): what a frame's level-two fit starts from ----
    import types as _types
    for case, (flags, fit) in {
            "inherit": (dict(init_scales_w_xyz_dist=False, inherit_prev_color=True, inherit_prev_opacity=True, inherit_prev_scales=True,
                             inherit_prev_rotation=True), dict(fit_color=True, fit_opacity=True, fit_scales=True, fit_rotation=True)),
            "dist_scales": (dict(init_scales_w_xyz_dist=True, inherit_prev_color=False, inherit_prev_opacity=True, inherit_prev_scales=True,
                                 inherit_prev_rotation=True), dict(fit_color=True, fit_opacity=True, fit_scales=True, fit_rotation=False))}.items():
        g = torch.Generator().manual_seed(3)
        V, Vp = 90, 60
        gm = object.__new__(GM)
        gm._visual_xyz = torch.rand(V, 3, generator=g) * 20.0
        gm._visual_color, gm._visual_opacity = torch.rand(V, 1, generator=g), torch.randn(V, 1, generator=g)
        gm._visual_scales, gm._visual_rotation = torch.randn(V, 3, generator=g) - 5.0, torch.randn(V, 4, generator=g)
        prev = dict(color=torch.rand(Vp, 1, generator=g), opacity=torch.randn(Vp, 1, generator=g), scales=torch.randn(Vp, 3, generator=g) - 5.0,
                    rotation=torch.randn(Vp, 4, generator=g))
        for k, v in fit.items():
            setattr(gm, k, v)
        for k in ("_visual_xyz", "_visual_color", "_visual_opacity", "_visual_scales", "_visual_rotation"):
            out[f"l2init/{case}/in{k}"] = getattr(gm, k).numpy().copy()
        for k, v in prev.items():
            out[f"l2init/{case}/prev_{k}"] = v.numpy().copy()
        for k, v in {**flags, **fit}.items():
            out[f"l2init/{case}/flag/{k}"] = v
        with cuda_as_cpu():
            gm.init_quantities_current_level_two(_types.SimpleNamespace(**flags), prev["color"], prev["opacity"], prev["scales"], prev["rotation"])
        for k in ("_visual_color", "_visual_opacity", "_visual_scales", "_visual_rotation"):
            out[f"l2init/{case}/out{k}"] = getattr(gm, k).numpy().copy()
        print("l2init", case, out[f"l2init/{case}/out_visual_scales"].dtype, float(out[f"l2init/{case}/out_visual_scales"].mean()))
    # ---- save_particles_* (gm_dynamics.py:1938-2017): file names and contents of the quantity snapshots ----
    import tempfile
    g = torch.Generator().manual_seed(9)
    gm = object.__new__(GM)
    gm.scale_factor = 100.0
    gm._xyz, gm._estimate_xyz, gm._visual_xyz = torch.rand(12, 3, generator=g) * 30, torch.rand(12, 3, generator=g) * 30, torch.rand(5, 3, generator=g) * 30
    gm._rigid_xyz, gm._estimate_xyz_nn = torch.rand(4, 3, generator=g) * 30, torch.rand(12, 3, generator=g)
    gm._visual_color, gm._visual_scales = torch.rand(5, 1, generator=g), torch.rand(5, 3, generator=g)
    gm._visual_rotation, gm._visual_opacity = torch.rand(5, 4, generator=g), torch.rand(5, 1, generator=g)
    for k in ("_xyz", "_estimate_xyz", "_visual_xyz", "_rigid_xyz", "_estimate_xyz_nn", "_visual_color", "_visual_scales", "_visual_rotation", "_visual_opacity"):
        out[f"quant/in{k}"] = getattr(gm, k).numpy().copy()
    advected = torch.rand(5, 3, generator=g)
    out["quant/in_advected"] = advected.numpy().copy()
    with tempfile.TemporaryDirectory() as d:
        gm.save_particles_rigid_body(d, 3)
        gm.save_particles_frame(d, 3)
        gm.save_particles_simulation(d, 41)
        gm.save_particles_simulation_guess(d, 41)
        gm.save_particles_optimization_first(d, 0, 250)
        gm.save_particles_optimization(d, advected, 7, 120)
        gm.save_particles_optimization_level_two(d, 7, 30)
        files = sorted(os.listdir(d))
        out["quant/files"] = np.array(files)
        for f in files:
            out["quant/file/" + f] = np.load(os.path.join(d, f))
    print("quantities", len(files), "files")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
