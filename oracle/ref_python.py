"""Loader for the reference's own, unmodified Python staged as sourceless bytecode in oracle/_ref (TEST INFRASTRUCTURE ONLY;
produced by oracle/build_ref.py:stage_python where /root/reference exists, shipped to the GPU box with the snapshot).

    use_reference_python(native="fnx")      the five plugin imports resolve to libfnx's drop-ins (fluidnexus_b200/compat):
                                            the reference's gm_*.py / pipe_*.py / loss_utils.py then run on CUDA through libfnx
    use_reference_python(native="reference") diff_gaussian_rasterization_ch1/_ch3 resolve to the reference's own wrapper packages
                                            on top of the compiled reference extensions (oracle/_ref/ch*/...so); torch_cluster /
                                            torch_scatter (third party, not installable here) resolve to the host-side
                                            restatements in oracle/pbf_ref.py -- bench.py's reference arm

Third-party modules the reference imports at module top but that are absent from this image and unused on the paths we run
(kornia.create_meshgrid, plyfile, lovely_tensors) are replaced by empty stand-ins.
"""
import importlib
import marshal
import os
import sys
import types

from . import build_ref

_STATE = {"native": None}


def staged():
    return build_ref.python_staged()


def _stub_missing():
    def have(name):
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False
    if not have("kornia"):
        k = types.ModuleType("kornia")

        def create_meshgrid(*a, **k_):
            raise RuntimeError("kornia is not installed (stand-in from oracle/ref_python.py)")
        k.create_meshgrid = create_meshgrid
        sys.modules["kornia"] = k
    if not have("plyfile"):
        p = types.ModuleType("plyfile")
        p.PlyData, p.PlyElement = object, object
        sys.modules["plyfile"] = p
    if not have("diff_gaussian_rasterization"):
        # upstream 3DGS rasterizer (PyPI), imported by renderer/pipe.py:5-8 for the `render_gs` pipe that no config selects
        d = types.ModuleType("diff_gaussian_rasterization")
        d.GaussianRasterizationSettings, d.GaussianRasterizer = None, None
        sys.modules["diff_gaussian_rasterization"] = d
    if not have("lovely_tensors"):
        lt = types.ModuleType("lovely_tensors")
        lt.monkey_patch = lambda *a, **k_: None
        sys.modules["lovely_tensors"] = lt


def use_reference_python(native="fnx"):
    """Put the staged reference Python on sys.path with the plugin boundary bound to `native`.  One binding per process."""
    if not staged():
        raise FileNotFoundError("oracle/_ref/FluidDynamics is missing: run `python oracle/build_ref.py` where /root/reference exists")
    if _STATE["native"] is not None:
        if _STATE["native"] != native:
            raise RuntimeError(f"reference Python already bound to {_STATE['native']!r} in this process")
        return
    _stub_missing()
    if native == "fnx":
        import fluidnexus_b200
        fluidnexus_b200.install_compat()
    elif native == "reference":
        from . import pbf_ref as O
        from . import ref_ext
        import torch
        for key, pkg in (("ch3", "diff_gaussian_rasterization_ch3"), ("ch1", "diff_gaussian_rasterization_ch1")):
            if ref_ext.available(key):
                sys.modules[pkg + "._C"] = ref_ext.load(key)       # `from . import _C` of the wrapper finds it here
        if build_ref.PKG_ZIP not in sys.path:
            sys.path.insert(0, build_ref.PKG_ZIP)
        # host-side stand-ins for the third-party neighbour search (documented as such in every report)
        tc = types.ModuleType("torch_cluster")
        tc.radius = lambda x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, **kw: O.radius(x, y, r, max_num_neighbors=max_num_neighbors)
        tc.radius_graph = lambda x, r, batch=None, loop=False, max_num_neighbors=32, **kw: O.radius_graph(x, r, loop=loop, max_num_neighbors=max_num_neighbors)
        ts = types.ModuleType("torch_scatter")
        ts.scatter_min = lambda src, index, dim=0, dim_size=None: O.scatter_min(src, index, dim=dim, dim_size=dim_size)
        sk, skc = types.ModuleType("simple_knn"), types.ModuleType("simple_knn._C")
        if ref_ext.available("knn"):
            skc.distCUDA2 = ref_ext.load("knn").distCUDA2
        else:
            skc.distCUDA2 = lambda pts: torch.tensor(O.knn3_mean_dist2(pts.detach().cpu().numpy()), device=pts.device)
        sk._C = skc
        sys.modules.update({"torch_cluster": tc, "torch_scatter": ts, "simple_knn": sk, "simple_knn._C": skc})
    else:
        raise ValueError(native)
    if build_ref.PY_ZIP not in sys.path:
        sys.path.insert(0, build_ref.PY_ZIP)
    _STATE["native"] = native


def reference_wrapper_on(pkg, c_module, alias):
    """The reference's own wrapper package `pkg` (`diff_gaussian_rasterization_ch3` / `_ch1`: its unmodified __init__, staged
    bytecode) executed under the module name `alias` with its `from . import _C` bound to `c_module` -- e.g. libfnx's `_C`-level
    drop-in (fluidnexus_b200/compat/<pkg>/_C.py).  Independent of use_reference_python()'s per-process binding."""
    import zipimport
    if not staged():
        raise FileNotFoundError("oracle/_ref/FluidDynamics is missing: run `python oracle/build_ref.py` where /root/reference exists")
    code = zipimport.zipimporter(build_ref.PKG_ZIP).get_code(pkg)
    mod = types.ModuleType(alias)
    mod.__path__, mod.__package__ = [], alias
    mod._C = c_module
    sys.modules[alias], sys.modules[alias + "._C"] = mod, c_module
    exec(code, mod.__dict__)
    return mod


def loop_body(name):
    """Code object of one optimisation-loop body of the reference's entry scripts (see build_ref.LOOPS), and where it
    was cut from: (relative path, (first line, last line))."""
    with open(os.path.join(build_ref.PY_OUT, "_loop_bodies.marshal"), "rb") as fh:
        blob = marshal.load(fh)
    if tuple(blob["python"]) != tuple(sys.version_info[:2]):
        raise RuntimeError(f"loop bodies were compiled by Python {blob['python']}, this is {sys.version_info[:2]}")
    return marshal.loads(blob["bodies"][name]), blob["where"][name]


def config_path(name):
    return os.path.join(build_ref.PY_OUT, "configs", name + ".json")


def parse_args(config, model_path, extra=()):
    """The reference's own get_parser() (helpers/helper_parser.py:15-66) on `--config_path <config>.json`: returns
    (args, model_args, optim_args, pipe_args) exactly as the entry scripts receive them.  safe_state() seeds the RNGs and
    selects cuda:0 like the reference does."""
    from helpers.helper_parser import get_parser
    argv = sys.argv
    try:
        sys.argv = ["reference_entry", "--config_path", config_path(config), "--model_path", model_path, "--quiet", *extra]
        return get_parser()
    finally:
        sys.argv = argv


class NullWriter:
    """tb_writer stand-in: the loop bodies log ~10 scalars per view through .item() (train_physical_particle.py:359-373)."""

    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, value, step=None):
        self.scalars[tag] = value
