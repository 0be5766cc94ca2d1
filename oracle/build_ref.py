"""Build recipe for the *unmodified reference* CUDA extensions (TEST INFRASTRUCTURE ONLY).

Compiles the reference's own sources **where they lie** under /root/reference
(nothing is copied into this repo) and writes objects + shared libraries only
into ``oracle/_ref/`` (git-ignored, but shipped to the GPU box by gpurun):

    oracle/_ref/ch3/fnx_ref_raster_ch3.so   <- FluidDynamics/submodules/gaussian_rasterization_ch3
    oracle/_ref/ch1/fnx_ref_raster_ch1.so   <- FluidDynamics/submodules/gaussian_rasterization_ch1
    oracle/_ref/knn/fnx_ref_simple_knn.so   <- FluidDynamics/submodules/simple-knn

Reference build description followed: R3/setup.py:9-29 (five sources + the
vendored glm include path, no extra flags), KNN/setup.py.  The reference's own
build system is NOT run; we hand the same source list to torch's ninja JIT
builder with an explicit build directory.

`stage_python()` additionally compiles the reference's *Python* (FluidDynamics/{arguments,gaussian_splatting,helpers,renderer,scene,
utils,entries_*}/*.py and the two rasterizer wrapper packages) to SOURCELESS bytecode under ``oracle/_ref/FluidDynamics`` and
``oracle/_ref/pkgs`` (py_compile of the files where they lie; no source text is written anywhere), copies the JSON configs
(data), and stores the hot-loop bodies of the entry scripts -- `train()` is one long function, so the statements of its
`for itr in ...` loops are cut out of the parsed module and compiled as they are -- as marshalled code objects in
``oracle/_ref/FluidDynamics/_loop_bodies.marshal``.  That is what lets the `-m gpu` tests and bench.py's reference arm execute
the reference's own, unmodified Python on the GPU box, where /root/reference does not exist.

Only tests/, bench.py's reference/cpu_baseline legs and
__graft_entry__.build()/smoke() may use what this produces.  The product
(fluidnexus_b200) never imports it.
"""
import os
import sys

REF = "/root/reference/FluidDynamics/submodules"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

TARGETS = {
    "ch3": dict(
        name="fnx_ref_raster_ch3",
        root=os.path.join(REF, "gaussian_rasterization_ch3"),
        sources=["cuda_rasterizer/rasterizer_impl.cu", "cuda_rasterizer/forward.cu",
                 "cuda_rasterizer/backward.cu", "rasterize_points.cu", "ext.cpp"],
        glm=True,
    ),
    "ch1": dict(
        name="fnx_ref_raster_ch1",
        root=os.path.join(REF, "gaussian_rasterization_ch1"),
        sources=["cuda_rasterizer/rasterizer_impl.cu", "cuda_rasterizer/forward.cu",
                 "cuda_rasterizer/backward.cu", "rasterize_points.cu", "ext.cpp"],
        glm=True,
    ),
    "knn": dict(
        name="fnx_ref_simple_knn",
        root=os.path.join(REF, "simple-knn"),
        sources=["spatial.cu", "simple_knn.cu", "ext.cpp"],
        glm=False,
    ),
}


def so_path(key):
    t = TARGETS[key]
    return os.path.join(OUT, key, t["name"] + ".so")


def build(keys=None, verbose=False):
    """Build the requested reference extensions if /root/reference exists. Returns list of built keys."""
    if not os.path.isdir(REF):
        return []
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load

    built = []
    for key in keys or list(TARGETS):
        t = TARGETS[key]
        bdir = os.path.join(OUT, key)
        os.makedirs(bdir, exist_ok=True)
        if os.path.exists(so_path(key)):
            built.append(key)
            continue
        inc = [t["root"]]
        if t["glm"]:
            inc.append(os.path.join(t["root"], "third_party", "glm"))
        load(
            name=t["name"],
            sources=[os.path.join(t["root"], s) for s in t["sources"]],
            extra_include_paths=inc,
            extra_cuda_cflags=["-lineinfo"],
            build_directory=bdir,
            is_python_module=False,  # just build; loading needs libcuda at import on some setups
            verbose=verbose,
        )
        built.append(key)
    return built


FD = "/root/reference/FluidDynamics"
PY_OUT = os.path.join(OUT, "FluidDynamics")
PY_ZIP = os.path.join(PY_OUT, "reference_python.zip")        # sys.path entry: FluidDynamics/ as sourceless bytecode
PKG_ZIP = os.path.join(PY_OUT, "reference_wrappers.zip")      # sys.path entry: diff_gaussian_rasterization_ch{1,3}/__init__
PY_DIRS = ["arguments", "gaussian_splatting", "helpers", "renderer", "scene", "utils", "entries_fluid_nexus", "entries_scalar_real"]
WRAPPERS = {"diff_gaussian_rasterization_ch3": "submodules/gaussian_rasterization_ch3/diff_gaussian_rasterization_ch3/__init__.py",
            "diff_gaussian_rasterization_ch1": "submodules/gaussian_rasterization_ch1/diff_gaussian_rasterization_ch1/__init__.py"}
# entry script -> {name: how to find the loop inside train()}: "nested" = the `for itr` loop inside `for cur_time_index`,
# "top" = the first `for itr` loop directly in train()
LOOPS = {
    "entries_fluid_nexus/train_physical_particle.py": {"fluid_nexus_physical_current": "nested", "fluid_nexus_physical_first": "top"},
    "entries_scalar_real/train_physical_particle.py": {"scalar_real_physical_current": "nested", "scalar_real_physical_first": "top"},
    "entries_fluid_nexus/train_visual_particle.py": {"fluid_nexus_visual_current": "nested"},
    "entries_scalar_real/train_visual_particle.py": {"scalar_real_visual_current": "nested"},
}


def _loop_body_code(path, which):
    """The statements of one optimisation loop of train(), up to (not including) its first top-level `if` (the periodic
    np.save / report blocks that follow the optimiser step), compiled unchanged."""
    import ast
    tree = ast.parse(open(path).read(), filename=path)
    train = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "train")

    def is_itr_loop(n):
        if not isinstance(n, ast.For):
            return False
        t = n.target
        return isinstance(t, ast.Name) and t.id == "itr"

    def find(nodes, nested):
        for n in nodes:
            if is_itr_loop(n) and not nested:
                return n
            if isinstance(n, ast.For) and isinstance(n.target, ast.Name) and n.target.id == "cur_time_index":
                if nested:
                    return find(n.body, False)
        return None
    loop = find(train.body, which == "nested")
    if loop is None:
        raise KeyError(f"{which} itr loop not found in {path}")
    body = []
    for st in loop.body:
        if isinstance(st, ast.If):
            break
        body.append(st)
    mod = ast.Module(body=body, type_ignores=[])
    return compile(mod, path, "exec"), (loop.lineno, body[-1].end_lineno)


def stage_python(verbose=False):
    """Sourceless bytecode of the reference's Python (two zip archives: zipimport loads .pyc members, and loose .pyc files do not
    travel with gpurun snapshots) + configs + loop bodies -> oracle/_ref (see the module docstring)."""
    if not os.path.isdir(FD):
        return False
    import marshal
    import py_compile
    import shutil
    import tempfile
    import zipfile
    os.makedirs(PY_OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="fnx_stage_")

    def bytecode(src, dfile):
        dst = os.path.join(tmp, "x.pyc")
        py_compile.compile(src, cfile=dst, dfile=dfile, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        return open(dst, "rb").read()
    empty = os.path.join(tmp, "empty.py")
    open(empty, "w").close()
    n = 0
    with zipfile.ZipFile(PY_ZIP, "w", zipfile.ZIP_DEFLATED) as z:
        for d in PY_DIRS:
            has_init = False
            for root, _, files in os.walk(os.path.join(FD, d)):
                for f in sorted(files):
                    if not f.endswith(".py"):
                        continue
                    src = os.path.join(root, f)
                    rel = os.path.relpath(src, FD)
                    has_init |= rel == os.path.join(d, "__init__.py")
                    z.writestr(rel + "c", bytecode(src, os.path.join("FluidDynamics", rel)))
                    n += 1
            if not has_init:   # the reference uses implicit namespace packages here; inside a zip an (empty) regular package is the safe form
                z.writestr(os.path.join(d, "__init__.pyc"), bytecode(empty, os.path.join("FluidDynamics", d, "__init__.py")))
    with zipfile.ZipFile(PKG_ZIP, "w", zipfile.ZIP_DEFLATED) as z:
        for pkg, rel in WRAPPERS.items():
            z.writestr(os.path.join(pkg, "__init__.pyc"), bytecode(os.path.join(FD, rel), os.path.join("FluidDynamics", rel)))
            n += 1
    os.makedirs(os.path.join(PY_OUT, "configs"), exist_ok=True)
    for f in os.listdir(os.path.join(FD, "configs")):
        if f.endswith(".json"):
            shutil.copyfile(os.path.join(FD, "configs", f), os.path.join(PY_OUT, "configs", f))
    bodies, where = {}, {}
    for rel, loops in LOOPS.items():
        for name, which in loops.items():
            code, lines = _loop_body_code(os.path.join(FD, rel), which)
            bodies[name] = marshal.dumps(code)
            where[name] = (rel, lines)
    with open(os.path.join(PY_OUT, "_loop_bodies.marshal"), "wb") as fh:
        marshal.dump({"bodies": bodies, "where": where, "python": sys.version_info[:2]}, fh)
    shutil.rmtree(tmp, ignore_errors=True)
    if verbose:
        print(f"staged {n} bytecode modules, {len(bodies)} loop bodies:", where)
    return True


def python_staged():
    return all(os.path.exists(p) for p in (PY_ZIP, PKG_ZIP, os.path.join(PY_OUT, "_loop_bodies.marshal")))


if __name__ == "__main__":
    keys = [k for k in sys.argv[1:] if k in TARGETS] or None
    print("built:", build(keys, verbose=True))
    print("python staged:", stage_python(verbose=True))
