"""CPU: the reference's own Python, staged as sourceless bytecode by oracle/build_ref.py:stage_python (git-ignored oracle/_ref),
imports through the drop-in packages, and the loop bodies cut out of its entry scripts are the ranges DESIGN.md cites; the opt-in
accelerators patch / restore the reference's modules; CachedImage behaves like the tensor it wraps."""
import os
import subprocess
import sys

import pytest
import torch

from oracle import build_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not build_ref.python_staged(), reason="oracle/_ref/FluidDynamics not staged (needs /root/reference at build time)")


def _run(code):
    p = subprocess.run([sys.executable, "-c", f"import sys; sys.path.insert(0, {ROOT!r})\n" + code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-3000:]
    return p.stdout


def test_staged_reference_python_imports_through_the_dropins():
    out = _run("""
from oracle import ref_python as RP
RP.use_reference_python("fnx")
import diff_gaussian_rasterization_ch3, diff_gaussian_rasterization_ch1, torch_cluster, torch_scatter, simple_knn._C
from gaussian_splatting.gm_dynamics import GaussianModel as GD
from gaussian_splatting.gm_fluid import GaussianModel as GF
from renderer.pipe_dynamics import render_dynamics
from renderer.pipe_fluid import render_fluid
from helpers.helper_pipe import get_render_pipe
from utils.loss_utils import l1_loss, ssim, distance_loss, l2_loss, l2_loss_consistency
from scene.camera import Camera
f, S, Rz = get_render_pipe("render_dynamics")
assert f is render_dynamics and S.__module__.startswith("fluidnexus_b200") and "compat" in torch_cluster.__file__
assert GD.__module__ == "gaussian_splatting.gm_dynamics" and GD.get_visual_xyz_from_nn.__code__.co_filename.startswith("FluidDynamics/")
for name, rng in (("fluid_nexus_physical_current", (329, 432)), ("scalar_real_physical_current", (301, 381)),
                  ("fluid_nexus_visual_current", (133, 222)), ("scalar_real_visual_current", (129, 218))):
    code, where = RP.loop_body(name)
    assert tuple(where[1]) == rng, (name, where)
    assert "gaussians" in code.co_names and "optimizer" in code.co_names and "save_particles_optimization" not in code.co_names
print("ok")
""")
    assert out.strip().endswith("ok")


def test_accelerators_patch_and_restore_the_reference_modules():
    out = _run("""
from oracle import ref_python as RP
RP.use_reference_python("fnx")
from fluidnexus_b200 import accelerate
import utils.loss_utils as LU
from gaussian_splatting.gm_fluid import GaussianModel as GF
orig = (LU.ssim, LU.distance_loss, GF.get_visual_xyz_from_nn)
accelerate.install_accelerators()
from gaussian_splatting.gm_dynamics import GaussianModel as GD      # imported AFTER the hook was installed
from scene.camera import Camera
assert LU.ssim.__module__ == LU.distance_loss.__module__ == "fluidnexus_b200.accelerate"
assert GF.get_visual_xyz_from_nn.__module__ == GD.get_gas_constraints_from_exyz_nn.__module__ == "fluidnexus_b200.accelerate"
assert Camera._fnx_patched
import torch
a, b = torch.rand(3, 16, 16), torch.rand(3, 16, 16)
assert abs(float(LU.ssim(a, b)) - float(orig[0](a, b))) < 1e-7      # CPU tensors fall through to the reference's own code
accelerate.uninstall_accelerators()
assert (LU.ssim, LU.distance_loss, GF.get_visual_xyz_from_nn) == orig and GD.get_gas_constraints_from_exyz_nn.__module__ == "gaussian_splatting.gm_dynamics"
print("ok")
""")
    assert out.strip().endswith("ok")


def test_cached_image_is_a_plain_tensor_everywhere_else():
    from fluidnexus_b200.accelerate import CachedImage
    img = torch.rand(3, 8, 8)
    c = CachedImage(img)
    assert c.float() is c and c.shape == img.shape and torch.equal(c * 2, img * 2)
    assert type(c.clamp(0, 1)) is torch.Tensor and type(c[None]) is torch.Tensor and type(c.double()) is torch.Tensor
    v0 = c._version
    c.mul_(0.5)
    assert c._version == v0 + 1      # an in-place edit is visible to the cache (the device copy would be refreshed)


def test_reference_wrapper_package_binds_to_the_C_level_dropin():
    """`diff_gaussian_rasterization_chN._C` of the drop-in packages has the reference's three pybind functions; the reference's OWN
    wrapper package (unmodified __init__, staged bytecode) imports on top of it and reaches libfnx -- which, on this GPU-less box,
    refuses CPU tensors loudly (there is no CPU path) after the wrapper's own argument checks have run."""
    out = _run("""
import importlib, inspect, torch
import fluidnexus_b200
fluidnexus_b200.install_compat()
from oracle import ref_python as RP
for C_ in (3, 1):
    pkg = f"diff_gaussian_rasterization_ch{C_}"
    c = importlib.import_module(pkg + "._C")
    # positional signatures of R3/rasterize_points.h:18-64 (18 / 20 / 3 arguments)
    assert [len(inspect.signature(getattr(c, n)).parameters) for n in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible")] == [18, 20, 3]
    w = RP.reference_wrapper_on(pkg, c, f"refwrap_ch{C_}")
    assert w.GaussianRasterizer.forward.__code__.co_filename.startswith("FluidDynamics/submodules/")     # the reference's code, not ours
    assert w._C is c and w.GaussianRasterizationSettings._fields == importlib.import_module(pkg).GaussianRasterizationSettings._fields
    rs = w.GaussianRasterizationSettings(image_height=32, image_width=32, tan_fov_x=0.5, tan_fov_y=0.5, bg=torch.zeros(C_), scale_modifier=1.0,
                                         view_matrix=torch.eye(4), proj_matrix=torch.eye(4), sh_degree=0, campos=torch.zeros(3), prefiltered=False)
    rz = w.GaussianRasterizer(raster_settings=rs)
    P = 7
    kw = dict(means3D=torch.rand(P, 3), means2D=torch.zeros(P, 3), opacities=torch.rand(P, 1), scales=torch.rand(P, 3), rotations=torch.rand(P, 4))
    try:
        rz(**kw)                                        # the wrapper's own check (__init__.py:184-185)
        raise SystemExit("no exception")
    except Exception as e:
        assert "exactly one of either SHs or precomputed colors" in str(e), e
    try:
        rz(colors_precomp=torch.rand(P, C_), **kw)
        raise SystemExit("no exception")
    except RuntimeError as e:
        assert "no CPU fallback" in str(e), e
print("ok")
""")
    assert out.strip().endswith("ok")
