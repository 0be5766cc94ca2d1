"""CPU: on-disk formats and rigid coupling pinned against the REFERENCE'S OWN code (tools/make_io_rigid_golden.py ran
gm_background.save_ply / load_ply, gm_dynamics.load_ply, gm_fluid.save_hidden / save_visual / load_hidden / load_visual,
check_inside_rigid_body and the two rigid projection methods from /root/reference and stored what they wrote / returned).

  * fluidnexus_b200/io.py writes the SAME BYTES as the reference's save_ply for the same tensors and reads the file the
    reference wrote into the same arrays as both of the reference's loaders;
  * the per-frame checkpoint writers produce the same file names, arrays and scalar JSON; the loaders the same state;
  * oracle/pbf_ref.py's rigid restatement (what the GPU kernel fnx_rigid_project is tested against) reproduces the reference's
    inside masks and projected positions for the three body kinds."""
import json
import os

import numpy as np
import pytest
import torch

from fluidnexus_b200 import io as IO
from oracle import pbf_ref as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLD, "pyref_io_rigid.npz"), allow_pickle=False)


def test_background_ply_bytes_equal_the_reference_writers(g, tmp_path):
    want = open(os.path.join(GOLD, "pyref_background_ply.bin"), "rb").read()
    p = tmp_path / "pc" / "point_cloud.ply"
    IO.save_background_ply(str(p), g["ply_in_xyz"], g["ply_in_color"], g["ply_in_opacity"], g["ply_in_scaling"], g["ply_in_rotation"])
    assert p.read_bytes() == want


def test_background_ply_reads_like_both_reference_loaders(g):
    d = IO.load_background_ply(os.path.join(GOLD, "pyref_background_ply.bin"))
    for mine, bg, dyn in (("xyz", "xyz", "xyz"), ("color", "color", "color"), ("opacity", "opacity", "opacity"), ("scaling", "scaling", "scales"),
                          ("rotation", "rotation", "rotation")):
        assert np.array_equal(d[mine], g["ply_bg_" + bg]), mine
        assert np.array_equal(d[mine], g["ply_dyn_" + dyn]), mine
    assert np.array_equal(d["xyz"], g["ply_in_xyz"]) and np.array_equal(d["rotation"], g["ply_in_rotation"])   # round trip of the raw values


def _model_state(g):
    names = ("xyz", "estimate_xyz", "velocity", "force", "buoyancy", "imass", "counts")
    st = {k: g["ckpt_model_" + k] for k in names}
    st["gravity"] = np.array([[0.0, -9.8, 0.0]], np.float32)
    st["particle_id"] = np.arange(st["xyz"].shape[0])[:, None]
    vis = {k: g["ckpt_model_" + k] for k in ("visual_xyz", "visual_color", "visual_scales", "visual_rotation", "visual_opacity")}
    return st, vis


def test_checkpoint_files_equal_the_reference_writers(g, tmp_path):
    st, vis = _model_state(g)
    scal = json.loads(str(g["ckpt_scalar_json"]))
    IO.save_hidden(str(tmp_path), 12, st, scal)
    IO.save_visual(str(tmp_path), 12, vis, scal["scale_factor"])
    assert sorted(os.listdir(tmp_path)) == [str(f) for f in g["ckpt_files"]]
    for f in g["ckpt_files"]:
        f = str(f)
        if f.endswith(".npy"):
            a, b = np.load(tmp_path / f), g["ckpt_" + f[:-4]]
            assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b), f
    assert (tmp_path / "frame_012_scalar_values.json").read_text() == str(g["ckpt_scalar_json"])


def test_checkpoint_loaders_equal_the_reference_loaders(g, tmp_path):
    for f in g["ckpt_files"]:
        f = str(f)
        if f.endswith(".npy"):
            np.save(tmp_path / f, g["ckpt_" + f[:-4]])
    (tmp_path / "frame_012_scalar_values.json").write_text(str(g["ckpt_scalar_json"]))
    st, scal = IO.load_hidden(str(tmp_path), 12)
    vis = IO.load_visual(str(tmp_path), 12, scal["scale_factor"])
    for k in ("xyz", "estimate_xyz", "velocity", "force", "buoyancy", "imass", "counts"):
        assert np.array_equal(st[k], g["ckpt_loaded_" + k]), k
    for k in ("visual_xyz", "visual_color", "visual_scales", "visual_rotation", "visual_opacity"):
        assert np.array_equal(vis[k], g["ckpt_loaded_" + k]), k
    ref = json.loads(str(g["ckpt_loaded_scalars"]))
    for k, v in ref.items():
        assert scal[k] == v, k


@pytest.mark.parametrize("kind", ["cuboid", "sphere", "cylinder"])
def test_rigid_oracle_equals_the_reference_methods(g, kind):
    prm = json.loads(str(g["rigid_params"]))
    geo = dict(cuboid_num=prm["cuboid_num"], particle_diameter=prm["diameter"], sphere_radius=prm["sphere_radius"],
               cylinder_radius=prm["cylinder_radius"], cylinder_num=prm["cylinder_num"])
    center, samples = g[f"rigid_{kind}_center"], torch.tensor(g[f"rigid_{kind}_samples"])
    for tag, mask_key, after, cap in (("xyz0", "mask", "xyz1", 0), ("vis0", "mask_vis", "vis1", 32)):
        x = torch.tensor(g[f"rigid_{kind}_{tag}"])
        mask = O.check_inside_rigid_body(kind, center, x, **geo)
        assert np.array_equal(mask.numpy(), g[f"rigid_{kind}_{mask_key}"]), (kind, tag)
        out = O.project_rigid(x, samples, mask, prm["H"], cap)
        assert np.array_equal(out.numpy(), g[f"rigid_{kind}_{after}"]), (kind, tag)
    assert abs(float(g[f"rigid_{kind}_ret_mask"]) - float(g[f"rigid_{kind}_mask"].mean())) < 1e-7   # the reference averages the mask in fp32


def test_solver_class_round_trips_the_reference_checkpoint(g, tmp_path):
    """PBFSolver.load_all / save_all (the reference's method names, gm_fluid.py:1653-1911, 1969-1972) on files WRITTEN BY THE
    REFERENCE: the restored state equals what the reference's own loaders restore, and saving it again reproduces the reference's
    files byte for byte (arrays and the scalar JSON text).  The solver object is made without its CUDA-only constructor; checkpoints
    are host-side I/O."""
    import ctypes as C
    from fluidnexus_b200.solver import PBFSolver
    src, dst = tmp_path / "in", tmp_path / "out"
    src.mkdir()
    for f in g["ckpt_files"]:
        f = str(f)
        if f.endswith(".npy"):
            np.save(src / f, g["ckpt_" + f[:-4]])
    (src / "frame_012_scalar_values.json").write_text(str(g["ckpt_scalar_json"]))
    sol = object.__new__(PBFSolver)
    sol.dev, sol.scale_factor = torch.device("cpu"), 100.0
    sol._gravity = (C.c_float * 3)(0.0, 0.0, 0.0)
    sol.load_all(str(src), 12)
    for k in ("xyz", "estimate_xyz", "velocity", "force", "buoyancy", "imass", "counts", "visual_xyz", "visual_color", "visual_scales",
              "visual_rotation", "visual_opacity"):
        assert np.array_equal(getattr(sol, "_" + k).numpy(), g["ckpt_loaded_" + k]), k
    ref = json.loads(str(g["ckpt_loaded_scalars"]))
    got = dict(secs=sol._secs, alpha=sol.alpha, k=sol.k, p0=sol.p0, buoyancy_max_y=sol.buoyancy_max_y, min_neighbors=sol.min_neighbors,
               emit_counter=sol.emit_counter, total_iterations=sol.total_iterations, particle_id_max=sol._particle_id_max)
    for k, v in ref.items():
        assert got[k] == v, k
    assert [round(float(c), 5) for c in sol._gravity] == [0.0, -9.8, 0.0]
    sol.save_all(str(dst), 12)
    assert sorted(os.listdir(dst)) == [str(f) for f in g["ckpt_files"]]
    for f in g["ckpt_files"]:
        f = str(f)
        if f.endswith(".npy"):
            a, b = np.load(dst / f), g["ckpt_" + f[:-4]]
            # positions go through render units -> scaled units -> render units in fp32: the reference's own round trip
            exact = not f.endswith(("_xyz.npy",))
            assert a.shape == b.shape and a.dtype == b.dtype and (np.array_equal(a, b) if exact else np.allclose(a, b, rtol=2e-7, atol=0)), f
    assert (dst / "frame_012_scalar_values.json").read_text() == str(g["ckpt_scalar_json"])
