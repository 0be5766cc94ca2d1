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


if __name__ == "__main__":
    keys = sys.argv[1:] or None
    print("built:", build(keys, verbose=True))
