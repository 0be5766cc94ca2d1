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
