"""CPU: this package's mirrors of the reference's Python interfaces keep the reference's names, argument order, defaults and
return-dictionary keys.  The expected values were extracted from the reference sources by tools/make_api_golden.py
(tests/golden/api_signatures.json); nothing under /root/reference is read here."""
import inspect
import json
import os

import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "api_signatures.json")))


def _sig(fn):
    ps = list(inspect.signature(fn).parameters.values())
    pos = [p.name for p in ps if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    defaults = [repr(p.default) if not isinstance(p.default, str) else repr(p.default) for p in ps
                if p.kind == p.POSITIONAL_OR_KEYWORD and p.default is not p.empty]
    kw = [p.name for p in ps if p.kind == p.VAR_KEYWORD]
    return pos, defaults, (kw[0] if kw else None)


@pytest.mark.parametrize("name", ["render_fluid", "render_dynamics", "render_background"])
def test_render_glue_signatures_and_return_keys(name):
    from fluidnexus_b200 import renderer
    g = GOLD[name]
    pos, defaults, kw = _sig(getattr(renderer, name))
    assert pos == g["args"]
    assert [d.replace("'", '"') for d in defaults] == [d.replace("'", '"') for d in g["defaults"]]
    assert kw == g["kwargs"]
    # the dictionary literal the mirror returns (parsed, not executed: needs no GPU)
    import ast
    src = inspect.getsource(renderer._rasterize)
    keys = []
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.Return) and isinstance(node.value, ast.Dict):
            keys = sorted({k.value for k in node.value.keys})
    assert keys == g["return_keys"]


def test_rasterizer_module_surface():
    from fluidnexus_b200 import rasterizer as R
    for channels in (1, 3):
        Settings, Rasterizer, _, _ = R.make_module(channels)
        assert list(Settings._fields) == GOLD["GaussianRasterizationSettings"]
        pos, defaults, _ = _sig(Rasterizer.forward)
        assert pos == GOLD["GaussianRasterizer.forward"]["args"]
        assert defaults == GOLD["GaussianRasterizer.forward"]["defaults"]
        assert _sig(Rasterizer.mark_visible)[0] == GOLD["GaussianRasterizer.mark_visible"]["args"]


def test_solver_method_surface():
    from fluidnexus_b200.solver import PBFSolver
    for name, g in GOLD["solver"].items():
        pos, defaults, _ = _sig(getattr(PBFSolver, name))
        # the mirror may add keyword arguments with defaults behind the reference's (project_gas_constraints(stats=False))
        assert pos[:len(g["args"])] == g["args"], name
        assert defaults[:len(g["defaults"])] == g["defaults"], name


def test_compat_packages_export_the_reference_names():
    import fluidnexus_b200
    compat = fluidnexus_b200.COMPAT_DIR
    for pkg, names in (("diff_gaussian_rasterization_ch1", ("GaussianRasterizationSettings", "GaussianRasterizer")),
                       ("diff_gaussian_rasterization_ch3", ("GaussianRasterizationSettings", "GaussianRasterizer")),
                       ("torch_cluster", ("radius", "radius_graph")), ("torch_scatter", ("scatter_min",))):
        src = open(os.path.join(compat, pkg, "__init__.py")).read()
        for n in names:
            assert n in src, (pkg, n)
    assert "distCUDA2" in open(os.path.join(compat, "simple_knn", "_C.py")).read()


def test_emitter_method_surface():
    """Particle creation / emission methods of the solver class against gm_dynamics.GaussianModel's (names, arguments, defaults)."""
    import inspect
    from fluidnexus_b200.solver import PBFSolver
    for name, g in GOLD["emitter"].items():
        sig = inspect.signature(getattr(PBFSolver, name))
        pos = [p.name for p in sig.parameters.values()]
        defaults = [repr(p.default) for p in sig.parameters.values() if p.default is not inspect.Parameter.empty]
        assert pos == g["args"] and defaults == g["defaults"], (name, pos, defaults, g)


def test_C_level_dropin_has_the_pybind_signatures():
    """`diff_gaussian_rasterization_chN._C` of the drop-in packages against the reference's pybind module: the functions ext.cpp
    binds, with the parameter names, order and count of their declarations in rasterize_points.h (parsed by tools/make_api_golden.py)."""
    import inspect
    from fluidnexus_b200 import rasterizer as R
    gold = GOLD["_C"]
    assert sorted(gold) == ["mark_visible", "rasterize_gaussians", "rasterize_gaussians_backward"]
    for C_ in (1, 3):
        fns = dict(zip(("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"), R.make_C(C_)))
        for name, g in gold.items():
            params = list(inspect.signature(fns[name]).parameters.values())
            got = [p.name for p in params]
            want = ["positions" if (name == "mark_visible" and a == "means3D") else a for a in g["args"]]
            assert got == want, (name, got, want)
            assert all(p.default is inspect.Parameter.empty for p in params), name      # pybind functions have no defaults
    assert (gold["rasterize_gaussians"]["returns"], gold["rasterize_gaussians_backward"]["returns"], gold["mark_visible"]["returns"]) == (7, 8, 1)
    import fluidnexus_b200
    for pkg in ("diff_gaussian_rasterization_ch1", "diff_gaussian_rasterization_ch3"):
        src = open(os.path.join(fluidnexus_b200.COMPAT_DIR, pkg, "_C.py")).read()
        assert all(n in src for n in gold), pkg


def test_hard_coded_constants_equal_the_reference_configs():
    """StepParams defaults, bench.py's per-workload constants, the solver defaults and the background learning rates against the
    reference's effective configs (arguments/__init__.py overridden by configs/*.json; tools/make_config_golden.py)."""
    import sys
    cfg = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pyref_configs.json")))
    from fluidnexus_b200.step import StepParams
    smoke, scalar = cfg["fluid_nexus_smoke_dynamics"], cfg["scalar_real"]
    d = StepParams()
    for mine, theirs in (("H", "H"), ("KNN_K", "KNN_K"), ("p0", "p0"), ("secs", "secs"), ("buoyancy_max_y", "buoyancy_max_y"),
                         ("lambda_dssim", "lambda_dssim"), ("lambda_image", "lambda_image"), ("lambda_current_distance", "lambda_current_distance"),
                         ("lambda_exyz", "lambda_exyz"), ("lambda_gas_constraints", "lambda_gas_constraints"),
                         ("lambda_next_gas_constraints", "lambda_next_gas_constraints"), ("distance_threshold_visual", "distance_threshold_visual")):
        assert getattr(d, mine) == pytest.approx(smoke[theirs]), mine
    assert d.lr == pytest.approx(smoke["position_lr_init"])          # spatial_lr_scale 1; never rescheduled (gm_fluid.py:401-407)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    for name, ref in (("smoke", smoke), ("scalar", scalar)):
        nf, nb, C, grey, size, N, p0, bmax, thr = bench.WORKLOADS[name]
        assert p0 == pytest.approx(ref["p0"]) and bmax == pytest.approx(ref["buoyancy_max_y"]) and thr == pytest.approx(ref["distance_threshold_visual"])
    # the two scene families differ exactly where the workloads differ
    assert smoke["p0"] != scalar["p0"] and smoke["buoyancy_max_y"] != scalar["buoyancy_max_y"]
    from oracle import pbf_ref as O
    sp = O.SolverParams()
    assert sp.H == smoke["H"] and sp.KNN_K == smoke["KNN_K"] and sp.p0 == smoke["p0"] and sp.secs == smoke["secs"]
    bgd = cfg["fluid_nexus_smoke_background"]
    # the learning rates the background tests / tools train with are the reference's
    for k, v in (("position_lr_init", 1.6e-4), ("position_lr_final", 1.6e-6), ("position_lr_delay_mult", 0.01), ("position_lr_max_steps", 30_000),
                 ("color_lr", 2.5e-3), ("opacity_lr", 0.05), ("scaling_lr", 5e-3), ("rotation_lr", 1e-3), ("percent_dense", 0.01)):
        assert bgd[k] == pytest.approx(v), k
