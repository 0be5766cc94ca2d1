"""CPU: on-disk formats (fluidnexus_b200/io.py) -- the file names, unit conventions and byte layout the reference writes
(gm_fluid.py:1653-1911, gm_background.py:184-269), checked by reading the files back with independent numpy code."""
import json
import os

import numpy as np
import pytest

from fluidnexus_b200 import io as IO


def _scalars():
    return dict(scale_factor=100.0, secs=0.033, alpha=-0.2, k=3, p0=1.5, buoyancy_decay_rate=0.0, buoyancy_max_y=0.0, min_neighbors=-1,
                remove_out_boundary=False, emit_ratio_hidden=1.0, emit_ratio_visual=0.5, emit_counter=7, total_iterations=11,
                total_sim_iterations=5, total_tb_log_iterations=3, particle_id_max=42)


def test_hidden_and_visual_checkpoints_round_trip_in_reference_layout(tmp_path):
    rng = np.random.default_rng(0)
    N, V = 37, 19
    st = {n: rng.normal(size=(N, 3)).astype(np.float32) for n in ("xyz", "estimate_xyz", "buoyancy", "force", "velocity")}
    st.update(imass=np.ones((N, 1), np.float32), counts=np.zeros((N, 1), np.float32), gravity=np.array([[0.0, -9.8, 0.0]], np.float32),
              particle_id=np.arange(N, dtype=np.int32))
    IO.save_hidden(str(tmp_path), 7, st, _scalars())
    names = sorted(os.listdir(tmp_path))
    assert names == sorted([f"frame_007_{n}.npy" for n, _ in IO.HIDDEN_ARRAYS] + ["frame_007_scalar_values.json"])
    # positions are stored in render units, everything else verbatim
    assert np.allclose(np.load(tmp_path / "frame_007_xyz.npy"), st["xyz"] / 100.0)
    assert np.array_equal(np.load(tmp_path / "frame_007_velocity.npy"), st["velocity"])
    assert json.load(open(tmp_path / "frame_007_scalar_values.json"))["particle_id_max"] == 42
    back, sc = IO.load_hidden(str(tmp_path), 7)
    for k in st:
        assert np.allclose(back[k], st[k], rtol=1e-6, atol=1e-6), k
    assert sc == _scalars() and back["particle_id"].dtype == np.int32
    # older checkpoints: no particle ids, no iteration counters
    os.remove(tmp_path / "frame_007_particle_id.npy")
    d = json.load(open(tmp_path / "frame_007_scalar_values.json"))
    for k in ("total_iterations", "particle_id_max", "emit_counter"):
        d.pop(k)
    json.dump(d, open(tmp_path / "frame_007_scalar_values.json", "w"))
    back, sc = IO.load_hidden(str(tmp_path), 7, defaults=dict(emit_counter=0))
    assert np.array_equal(back["particle_id"], np.arange(N)) and sc["total_iterations"] == 0 and sc["emit_counter"] == 0
    vis = dict(visual_xyz=rng.normal(size=(V, 3)).astype(np.float32), visual_color=rng.uniform(size=(V, 1)).astype(np.float32),
               visual_scales=rng.normal(size=(V, 3)).astype(np.float32), visual_rotation=rng.normal(size=(V, 4)).astype(np.float32),
               visual_opacity=rng.normal(size=(V, 1)).astype(np.float32))
    IO.save_visual(str(tmp_path), 7, vis, 100.0)
    assert np.allclose(np.load(tmp_path / "frame_007_visual_xyz.npy"), vis["visual_xyz"] / 100.0)
    vb = IO.load_visual(str(tmp_path), 7, 100.0)
    for k in vis:
        assert np.allclose(vb[k], vis[k], rtol=1e-6, atol=1e-6), k
    IO.save_visual(str(tmp_path), 8, vis, 100.0, scale=False)
    assert np.array_equal(np.load(tmp_path / "frame_008_visual_xyz.npy"), vis["visual_xyz"])


def test_background_ply_layout(tmp_path):
    rng = np.random.default_rng(1)
    P = 23
    xyz, color = rng.normal(size=(P, 3)).astype(np.float32), rng.uniform(size=(P, 3)).astype(np.float32)
    opac, scal, rot = rng.normal(size=(P, 1)).astype(np.float32), rng.normal(size=(P, 3)).astype(np.float32), rng.normal(size=(P, 4)).astype(np.float32)
    path = str(tmp_path / "pc" / "point_cloud.ply")
    IO.save_background_ply(path, xyz, color, opac, scal, rot)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", f"element vertex {P}"]
    props = [l.split()[2] for l in lines[3:]]
    assert all(l.startswith("property float ") for l in lines[3:])
    assert props == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "f_rest_0", "f_rest_1", "f_rest_2", "opacity", "scale_0",
                     "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3", "color_0", "color_1", "color_2"]
    assert props == IO.background_ply_properties(3)
    table = np.frombuffer(body, dtype="<f4").reshape(P, len(props))                     # independent reader
    assert np.array_equal(table[:, 0], -xyz[:, 0]) and np.array_equal(table[:, 1], -xyz[:, 1]) and np.array_equal(table[:, 2], xyz[:, 2])
    assert np.all(table[:, 3:6] == 0) and np.all(table[:, 9:12] == 0)
    assert np.allclose(table[:, 6:9], (color - 0.5) / 0.28209479177387814, rtol=1e-6)
    assert np.array_equal(table[:, 12:13], opac) and np.array_equal(table[:, 13:16], scal) and np.array_equal(table[:, 16:20], rot)
    assert np.array_equal(table[:, 20:23], color)
    back = IO.load_background_ply(path)
    assert np.array_equal(back["xyz"], xyz) and np.array_equal(back["color"], color) and np.array_equal(back["opacity"], opac)
    assert np.array_equal(back["scaling"], scal) and np.array_equal(back["rotation"], rot)
    assert xyz[0, 0] == back["xyz"][0, 0]                                                # the caller's array was not flipped in place


def test_ply_reader_accepts_ascii_and_reordered_columns(tmp_path):
    path = tmp_path / "a.ply"
    path.write_text("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
                    "property float opacity\nproperty float scale_1\nproperty float scale_0\nproperty float scale_2\n"
                    "property float rot_0\nproperty float rot_1\nproperty float rot_2\nproperty float rot_3\nproperty float color_0\n"
                    "end_header\n1 2 3 0.5 11 10 12 1 0 0 0 0.7\n-1 -2 -3 0.25 21 20 22 0 1 0 0 0.1\n")
    b = IO.load_background_ply(str(path))
    assert np.array_equal(b["xyz"], np.array([[-1, -2, 3], [1, 2, -3]], np.float32))
    assert np.array_equal(b["scaling"], np.array([[10, 11, 12], [20, 21, 22]], np.float32))       # ordered by suffix, not by column
    assert b["color"].shape == (2, 1) and b["rotation"].shape == (2, 4) and b["opacity"].shape == (2, 1)
    with pytest.raises(AssertionError):
        (tmp_path / "b.ply").write_text("not a ply\n")
        IO.read_ply_vertices(str(tmp_path / "b.ply"))


def test_load_visual_repeats_a_grey_colour_for_three_channel_scenes(tmp_path):
    """gm_dynamics.load_visual(..., color_3ch=True), gm_dynamics.py:2076-2078."""
    rng = np.random.default_rng(0)
    vis = dict(visual_xyz=rng.random((9, 3)).astype(np.float32), visual_color=rng.random((9, 1)).astype(np.float32),
               visual_scales=rng.random((9, 3)).astype(np.float32), visual_rotation=rng.random((9, 4)).astype(np.float32),
               visual_opacity=rng.random((9, 1)).astype(np.float32))
    IO.save_visual(str(tmp_path), 3, vis, 100.0, scale=False)
    plain = IO.load_visual(str(tmp_path), 3, 100.0, scale=False)
    rgb = IO.load_visual(str(tmp_path), 3, 100.0, scale=False, color_3ch=True)
    assert plain["visual_color"].shape == (9, 1) and rgb["visual_color"].shape == (9, 3)
    assert all(np.array_equal(rgb["visual_color"][:, c], vis["visual_color"][:, 0]) for c in range(3))
    assert np.array_equal(rgb["visual_xyz"], vis["visual_xyz"])                      # scale=False: render units kept


def test_load_visual_smoothed_reads_the_smoothed_files_where_asked(tmp_path):
    """load_visual_smoothed (gm_dynamics.py:2093-2150) through the solver class, made without its CUDA-only constructor."""
    import torch
    from fluidnexus_b200.solver import PBFSolver
    rng = np.random.default_rng(1)
    vis = {k: rng.random((6, c)).astype(np.float32) for k, c in (("visual_xyz", 3), ("visual_color", 1), ("visual_scales", 3),
                                                                  ("visual_rotation", 4), ("visual_opacity", 1))}
    IO.save_visual(str(tmp_path), 4, vis, 100.0)
    smooth = {k: v + 1.0 for k, v in vis.items() if k != "visual_xyz"}
    for k, v in smooth.items():
        np.save(tmp_path / f"frame_004_{k}_smoothed_ws5.npy", v)
    sol = object.__new__(PBFSolver)
    sol.dev, sol.scale_factor = torch.device("cpu"), 100.0
    assert sol.load_visual_smoothed(str(tmp_path), 4, smoothed_rotation=False) == 6
    assert np.allclose(sol._visual_xyz.numpy(), vis["visual_xyz"], rtol=1e-6)                   # saved / 100, loaded * 100
    assert np.array_equal(sol._visual_color.numpy(), smooth["visual_color"]) and np.array_equal(sol._visual_scales.numpy(), smooth["visual_scales"])
    assert np.array_equal(sol._visual_rotation.numpy(), vis["visual_rotation"]) and np.array_equal(sol._visual_opacity.numpy(), smooth["visual_opacity"])
    with pytest.raises(AssertionError, match="File not found"):
        sol.load_visual_smoothed(str(tmp_path), 4, window_size=7)
