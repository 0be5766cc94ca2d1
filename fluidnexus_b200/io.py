"""On-disk formats either side of the hot path, bit-compatible with what the reference writes:

* per-frame particle checkpoints   FD/gaussian_splatting/gm_fluid.py:1653-1760 (save_hidden / save_visual),
                                   :1811-1911 (load_hidden / load_visual):  `frame_{idx:03d}_{name}.npy` + one
                                   `frame_{idx:03d}_scalar_values.json`; positions are stored in RENDER units
                                   (divided by scale_factor on save, multiplied back on load)
* static-background point cloud    FD/gaussian_splatting/gm_background.py:184-269 (save_ply / load_ply): binary
                                   little-endian PLY, float32 vertex properties x y z nx ny nz f_dc_* f_rest_* opacity
                                   scale_* rot_* color_*, x and y NEGATED on disk (for the supersplat viewer), f_dc =
                                   (color - 0.5) / C0, raw (pre-activation) opacity / scale / rotation

Host-side numpy code: nothing here touches the GPU, and the reference's `plyfile` dependency is not needed (the PLY subset
it emits -- one `vertex` element of float properties -- is read and written directly).
"""
import json
import os

import numpy as np

C0 = 0.28209479177387814  # FD/utils/sh_utils.py:27

# name -> stored in render units (x / scale_factor)?   (gm_fluid.py:1653-1693)
HIDDEN_ARRAYS = (("xyz", True), ("estimate_xyz", True), ("buoyancy", False), ("force", False), ("velocity", False), ("imass", False),
                 ("counts", False), ("gravity", False), ("particle_id", False))
VISUAL_ARRAYS = ("visual_xyz", "visual_color", "visual_scales", "visual_rotation", "visual_opacity")   # gm_fluid.py:1720-1745
SCALAR_KEYS = ("scale_factor", "secs", "alpha", "k", "p0", "buoyancy_decay_rate", "buoyancy_max_y", "min_neighbors", "remove_out_boundary",
               "emit_ratio_hidden", "emit_ratio_visual", "emit_counter", "total_iterations", "total_sim_iterations",
               "total_tb_log_iterations", "particle_id_max")
_OPTIONAL_SCALARS = {"total_iterations": 0, "total_sim_iterations": 0, "total_tb_log_iterations": 0, "particle_id_max": 0}


def _frame_file(checkpoint_path, frame_idx, name, ext="npy"):
    return os.path.join(checkpoint_path, f"frame_{frame_idx:03d}_{name}.{ext}")


def _np(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def save_hidden(checkpoint_path, frame_idx, state, scalars):
    """state: dict with the HIDDEN_ARRAYS names (scaled units, like the model holds them); scalars: dict with SCALAR_KEYS."""
    os.makedirs(checkpoint_path, exist_ok=True)
    sf = float(scalars["scale_factor"])
    for name, in_render_units in HIDDEN_ARRAYS:
        a = _np(state[name])
        np.save(_frame_file(checkpoint_path, frame_idx, name), a / sf if in_render_units else a)
    with open(_frame_file(checkpoint_path, frame_idx, "scalar_values", "json"), "w") as f:
        json.dump({k: scalars[k] for k in SCALAR_KEYS}, f)


def load_hidden(checkpoint_path, frame_idx, defaults=None):
    """-> (state dict of float32 arrays in scaled units (particle_id int32), scalars dict).  A missing particle_id file
    gives arange(N) like the reference; optional scalar keys fall back to `defaults` / 0 (gm_fluid.py:1847-1884)."""
    with open(_frame_file(checkpoint_path, frame_idx, "scalar_values", "json")) as f:
        stored = json.load(f)
    defaults = dict(defaults or {})
    scalars = {}
    for k in SCALAR_KEYS:
        if k in stored:
            scalars[k] = stored[k]
        elif k in _OPTIONAL_SCALARS or k in defaults or k.startswith("emit_"):
            scalars[k] = defaults.get(k, _OPTIONAL_SCALARS.get(k))
        else:
            raise KeyError(f"{k} missing from {_frame_file(checkpoint_path, frame_idx, 'scalar_values', 'json')}")
    sf = float(scalars["scale_factor"])
    state = {}
    for name, in_render_units in HIDDEN_ARRAYS:
        path = _frame_file(checkpoint_path, frame_idx, name)
        if name == "particle_id":
            state[name] = np.load(path).astype(np.int32) if os.path.exists(path) else np.arange(state["xyz"].shape[0], dtype=np.int32)
            continue
        assert os.path.exists(path), f"File not found: {path}"
        a = np.load(path).astype(np.float32)
        state[name] = a * np.float32(sf) if in_render_units else a
    return state, scalars


def save_visual(checkpoint_path, frame_idx, state, scale_factor, scale=True):
    os.makedirs(checkpoint_path, exist_ok=True)
    for name in VISUAL_ARRAYS:
        a = _np(state[name])
        np.save(_frame_file(checkpoint_path, frame_idx, name), a / float(scale_factor) if (name == "visual_xyz" and scale) else a)


def load_visual(checkpoint_path, frame_idx, scale_factor, scale=True, color_3ch=False, smoothed_window=None,
                smoothed=("visual_color", "visual_scales", "visual_rotation", "visual_opacity")):
    """color_3ch (gm_dynamics.py:2067-2078, the level-two stage of 3-channel scenes): a one-channel colour is repeated to three.
    smoothed_window = w (load_visual_smoothed, gm_dynamics.py:2093-2150): the attributes named in `smoothed` are read from the
    temporally smoothed files `frame_XXX_<name>_smoothed_ws<w>.npy` that the reference's post-processing writes next to the plain ones."""
    state = {}
    for name in VISUAL_ARRAYS:
        stem = f"{name}_smoothed_ws{int(smoothed_window)}" if (smoothed_window is not None and name in smoothed) else name
        path = _frame_file(checkpoint_path, frame_idx, stem)
        assert os.path.exists(path), f"File not found: {path}"
        a = np.load(path).astype(np.float32)
        state[name] = a * np.float32(scale_factor) if (name == "visual_xyz" and scale) else a
    if color_3ch and state["visual_color"].shape[1] == 1:
        state["visual_color"] = np.repeat(state["visual_color"], 3, axis=1)
    return state


# ---------------------------------------------------------------------------------------------------------------------
# "quantities": the .npy snapshots the entries drop for visualisation / debugging (gm_dynamics.py:1938-2017).  One table instead of
# seven methods: kind -> (file-name prefix, ((file stem, key into `tensors`, stored in render units?, skipped when empty?), ...))
# ---------------------------------------------------------------------------------------------------------------------
QUANTITIES = {
    "rigid_body": ("frame_{a:03d}_", (("rigid_xyz", "rigid_xyz", True, False),)),
    "frame": ("frame_{a:03d}_", (("xyz", "xyz", True, False), ("visual_xyz", "visual_xyz", True, True))),
    "simulation": ("{a:03d}_", (("xyz", "xyz", True, False), ("estimated_xyz", "estimate_xyz", True, False), ("visual_xyz", "visual_xyz", True, True))),
    "simulation_guess": ("{a:03d}_", (("guess_estimated_xyz", "estimate_xyz", True, False),)),
    "optimization_first": ("{a:03d}_{b:05d}_", (("visual_xyz", "visual_xyz", False, False),)),
    # estimate_xyz_nn is the trainable tensor (already in render units); visual_xyz here is the advected set handed in by the loop
    "optimization": ("{a:03d}_{b:05d}_", (("estimate_xyz_nn", "estimate_xyz_nn", False, False), ("visual_xyz", "visual_xyz", False, True))),
    "optimization_level_two": ("{a:03d}_{b:05d}_", tuple((n, n, False, False) for n in VISUAL_ARRAYS[1:])),
}


def save_particles(kind, quantities_path, tensors, a, b=0, scale_factor=100.0):
    """save_particles_<kind>(quantities_path, a[, b]) of the reference: a = frame / simulation index, b = iteration.  `tensors`: dict
    of arrays / tensors in the units the model holds them (scaled for positions).  Returns the files written."""
    prefix, entries = QUANTITIES[kind]
    os.makedirs(quantities_path, exist_ok=True)
    written = []
    for stem, key, render_units, skip_empty in entries:
        arr = _np(tensors[key])
        if skip_empty and arr.shape[0] == 0:
            continue
        path = os.path.join(quantities_path, prefix.format(a=a, b=b) + stem + ".npy")
        np.save(path, arr / float(scale_factor) if render_units else arr)
        written.append(path)
    return written


# ---------------------------------------------------------------------------------------------------------------------
# background point cloud (PLY)
# ---------------------------------------------------------------------------------------------------------------------
def background_ply_properties(n_color, n_scale=3, n_rot=4):
    """The property list of construct_list_of_attributes (gm_background.py:184-201), in file order."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_color)] + [f"f_rest_{i}" for i in range(n_color)] + ["opacity"]
    names += [f"scale_{i}" for i in range(n_scale)] + [f"rot_{i}" for i in range(n_rot)] + [f"color_{i}" for i in range(n_color)]
    return names


def save_background_ply(path, xyz, color, raw_opacity, raw_scaling, raw_rotation):
    """xyz [P,3], color [P,C], raw_opacity [P,1] (logit), raw_scaling [P,3] (log), raw_rotation [P,4] (un-normalised)."""
    xyz, color = np.array(_np(xyz), dtype=np.float32), np.asarray(_np(color), dtype=np.float32)
    opac = np.asarray(_np(raw_opacity), dtype=np.float32).reshape(-1, 1)
    scal, rot = np.asarray(_np(raw_scaling), dtype=np.float32), np.asarray(_np(raw_rotation), dtype=np.float32)
    xyz[:, 0] *= -1.0
    xyz[:, 1] *= -1.0
    cols = np.concatenate((xyz, np.zeros_like(xyz), (color - 0.5) / C0, np.zeros_like(xyz)[:, :color.shape[1]], opac, scal, rot, color), axis=1)
    names = background_ply_properties(color.shape[1], scal.shape[1], rot.shape[1])
    assert cols.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % cols.shape[0]
    header += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(np.ascontiguousarray(cols, dtype="<f4").tobytes())


_PLY_TYPES = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1", "char": "i1", "int8": "i1",
              "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4"}


def read_ply_vertices(path):
    """The first element of a PLY file as a numpy structured array (binary little/big endian or ascii, scalar properties)."""
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply", "not a PLY file"
        fmt, count, props, in_first = None, None, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError("PLY header without end_header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                if count is None:
                    count, in_first = int(tok[2]), True
                else:
                    in_first = False
            elif tok[0] == "property" and in_first:
                if tok[1] == "list":
                    raise ValueError("list properties are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=count, ndmin=2)
            out = np.empty(count, dtype=[(n, "<" + t) for n, t in props])
            for k, (n, _) in enumerate(props):
                out[n] = rows[:, k]
            return out
        order = "<" if fmt == "binary_little_endian" else ">"
        return np.frombuffer(f.read(count * sum(np.dtype(t).itemsize for _, t in props)), dtype=[(n, order + t) for n, t in props], count=count)


def load_background_ply(path):
    """-> dict(xyz, color, opacity [P,1], scaling, rotation) of float32 arrays, raw values as the model stores them; x and y
    are flipped back, colours come from color_* and the scale_* / rot_* / color_* columns are ordered by their numeric suffix
    (gm_background.py:232-262)."""
    v = read_ply_vertices(path)
    by_suffix = lambda prefix: sorted((n for n in v.dtype.names if n.startswith(prefix)), key=lambda n: int(n.split("_")[-1]))
    stack = lambda names: np.stack([np.asarray(v[n], dtype=np.float32) for n in names], axis=1) if names else np.zeros((v.shape[0], 0), np.float32)
    xyz = np.stack((-np.asarray(v["x"], np.float32), -np.asarray(v["y"], np.float32), np.asarray(v["z"], np.float32)), axis=1)
    return dict(xyz=xyz, color=stack(by_suffix("color_")), opacity=np.asarray(v["opacity"], np.float32)[:, None],
                scaling=stack(by_suffix("scale_")), rotation=stack(by_suffix("rot")))
