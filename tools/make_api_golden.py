"""Run in the build container (needs /root/reference): extracts the call signatures and return-dictionary keys of the
reference's render glue and the plugin classes by parsing their sources (nothing is imported or executed), and writes
tests/golden/api_signatures.json -- the fixture tests/test_api_surface.py pins this package's mirrors against."""
import ast
import json
import os

REF = "/root/reference/FluidDynamics"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "api_signatures.json")


def func_sig(path, name):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            a = node.args
            pos = [x.arg for x in a.args]
            defaults = [ast.unparse(d) for d in a.defaults]
            keys = []
            for sub in ast.walk(node):
                if isinstance(sub, ast.Return) and isinstance(sub.value, ast.Dict):
                    keys = [k.value for k in sub.value.keys]
            return {"args": pos, "defaults": defaults, "kwargs": a.kwarg.arg if a.kwarg else None, "return_keys": sorted(set(keys))}
    raise KeyError(name)


def class_fields(path, cls):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            return [s.target.id for s in node.body if isinstance(s, ast.AnnAssign)]
    raise KeyError(cls)


def method_sig(path, cls, name):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for s in node.body:
                if isinstance(s, ast.FunctionDef) and s.name == name:
                    return {"args": [x.arg for x in s.args.args], "defaults": [ast.unparse(d) for d in s.args.defaults]}
    raise KeyError((cls, name))


def pybind_sigs(header, ext):
    """The three functions the reference's `_C` module exports (ext.cpp: m.def(python name, &C++ function)) with the parameter names
    and C++ types of their declarations in rasterize_points.h and the arity of their std::tuple return types."""
    import re
    h = open(header).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    binds = dict(re.findall(r'm\.def\("(\w+)",\s*&(\w+)\)', open(ext).read()))
    out = {}
    for py, cpp in binds.items():
        m = re.search(r"([\w:<>,\s]+?)\s+" + cpp + r"\s*\((.*?)\)\s*;", h, flags=re.S)
        ret, params = m.group(1).strip(), m.group(2)
        names, types = [], []
        for prm in params.split(","):
            toks = prm.replace("&", " ").split()
            names.append(toks[-1])
            types.append("Tensor" if "Tensor" in prm else [t for t in toks[:-1] if t != "const"][-1])
        out[py] = {"cpp": cpp, "args": names, "types": types, "returns": ret.count("Tensor") + ret.count("int")}
    return out


def main():
    r3 = os.path.join(REF, "submodules/gaussian_rasterization_ch3/diff_gaussian_rasterization_ch3/__init__.py")
    gm = os.path.join(REF, "gaussian_splatting/gm_fluid.py")
    out = {
        "render_fluid": func_sig(os.path.join(REF, "renderer/pipe_fluid.py"), "render_fluid"),
        "render_dynamics": func_sig(os.path.join(REF, "renderer/pipe_dynamics.py"), "render_dynamics"),
        "render_background": func_sig(os.path.join(REF, "renderer/pipe_background.py"), "render_background"),
        "GaussianRasterizationSettings": class_fields(r3, "GaussianRasterizationSettings"),
        "GaussianRasterizer.forward": method_sig(r3, "GaussianRasterizer", "forward"),
        "GaussianRasterizer.mark_visible": method_sig(r3, "GaussianRasterizer", "mark_visible"),
        "_C": pybind_sigs(os.path.join(REF, "submodules/gaussian_rasterization_ch3/rasterize_points.h"),
                          os.path.join(REF, "submodules/gaussian_rasterization_ch3/ext.cpp")),
        "emitter": {m: method_sig(os.path.join(REF, "gaussian_splatting/gm_dynamics.py"), "GaussianModel", m) for m in
                    ("create_particles_visual", "create_particles_hidden", "prepare_emitter_points", "prepare_emitter_future_first_points",
                     "emit_new_particles", "create_rigid_body", "prepare_hidden_particles_for_rendering",
                     "prepare_visual_particles_for_rendering", "prepare_future_visual_particles_for_rendering",
                     "prepare_rigid_body_particles_for_rendering")},
        "solver": {m: method_sig(gm, "GaussianModel", m) for m in
                   ("guess_hidden_particles", "project_gas_constraints", "confirm_guess_hidden_particles", "update_visual_particles",
                    "remove_invalid_particles", "update_solver_counts")},
    }
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1)[:600])


if __name__ == "__main__":
    main()
