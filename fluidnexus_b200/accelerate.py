"""Opt-in accelerators for the UNCHANGED reference scripts, beyond the five plugin imports.

`install_compat()` makes the reference's native boundary (diff_gaussian_rasterization_ch1/_ch3, simple_knn, torch_cluster,
torch_scatter) resolve to libfnx.  After the kernels are fast, an unchanged training loop spends its time in the Python around
them: the ground-truth image is uploaded from the CPU for every view of every iteration
(FD/entries_fluid_nexus/train_physical_particle.py:353), `ssim` rebuilds its window on the CPU and runs five grouped convolutions
plus ~20 element-wise kernels and their autograd twins (FD/utils/loss_utils.py:21-64), `distance_loss` materialises a dense V x V
cdist (loss_utils.py:98-121, infeasible beyond ~10^5 particles).  `install_accelerators()` removes those WITHOUT editing a file of
the reference: an import hook patches two of its modules as they are imported,

  utils.loss_utils.l1_loss / ssim      -> fluidnexus_b200.losses (fused fnx_image_loss kernels, same signatures and values)
  utils.loss_utils.distance_loss       -> fluidnexus_b200.physics.pair_distance_loss (grid hash, O(V) memory, same value)
  scene.camera.Camera.original_image   -> a tensor that answers `.float().cuda()` / `.cuda()` with a device copy made once
                                          (the image of a camera never changes; the reference re-uploads it ~10^5 times per run)
  gaussian_splatting.gm_fluid / gm_dynamics GaussianModel.get_visual_xyz_from_nn, get_gas_constraints_from_exyz_nn,
  get_gas_constraints_from_vel_nn_guess -> fluidnexus_b200.physics.visual_advect / density_ratio (one fused forward and one fused
                                          backward gather each instead of an int64 edge list + ~12 gather / index_add_ kernels and
                                          their autograd twins, gm_fluid.py:1107-1158, 1291-1336; same values and gradients)

Everything else -- the loop, the model classes, the render pipes, logging with .item() -- stays the reference's.  This is a
convenience layer on top of the drop-in boundary, not part of it: nothing in fluidnexus_b200 depends on it, and the parity tests
(tests/test_reference_dropin_gpu.py) run the reference loop both ways.
"""
import importlib
import importlib.abc
import importlib.util
import sys

import torch

_PATCHERS = {}
_INSTALLED = {"finder": None}


class CachedImage(torch.Tensor):
    """A CPU image tensor whose transfer to the GPU happens once.  Behaves like the tensor it wraps everywhere else."""

    @staticmethod
    def __new__(cls, data):
        t = torch.Tensor._make_subclass(cls, data.detach(), False)
        t._fnx_dev = {}
        return t

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        # results of operations on a cached image are ordinary tensors: only the image itself carries the device copies
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **(kwargs or {}))

    def float(self):
        return self if self.dtype == torch.float32 else torch.Tensor.float(self.as_subclass(torch.Tensor))

    def cuda(self, device=None, non_blocking=False, **kw):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        cache = self.__dict__.setdefault("_fnx_dev", {})
        hit = cache.get(dev.index)
        if hit is None or hit[1] != self._version:      # (an in-place edit of the image bumps _version: upload again)
            hit = (self.as_subclass(torch.Tensor).to(dev), self._version)
            cache[dev.index] = hit
        return hit[0]

    def to(self, *a, **k):
        return self.as_subclass(torch.Tensor).to(*a, **k)


def _patch_loss_utils(mod):
    from . import losses, physics
    mod._fnx_original = {k: getattr(mod, k) for k in ("l1_loss", "ssim", "distance_loss") if hasattr(mod, k)}
    orig_l1, orig_ssim, orig_dist = mod._fnx_original.get("l1_loss"), mod._fnx_original.get("ssim"), mod._fnx_original.get("distance_loss")

    def l1_loss(network_output, gt):
        if network_output.is_cuda and network_output.dim() in (3, 4) and network_output.shape == gt.shape:
            return losses.l1_loss(network_output, gt)
        return orig_l1(network_output, gt)

    def ssim(img1, img2, window_size=11, size_average=True):
        if img1.is_cuda and window_size == 11 and size_average and img1.dim() in (3, 4) and img1.shape == img2.shape:
            return losses.ssim(img1, img2)
        return orig_ssim(img1, img2, window_size, size_average)

    def distance_loss(positions, threshold):
        if positions.is_cuda and positions.dim() == 2 and positions.size(1) == 3:
            return physics.pair_distance_loss(positions, threshold)
        return orig_dist(positions, threshold)

    mod.l1_loss, mod.ssim, mod.distance_loss = l1_loss, ssim, distance_loss


def _patch_model(mod):
    from . import physics
    gm = mod.GaussianModel
    if getattr(gm, "_fnx_patched", False):
        return
    gm._fnx_original = {k: getattr(gm, k) for k in ("get_visual_xyz_from_nn", "get_gas_constraints_from_exyz_nn", "get_gas_constraints_from_vel_nn_guess")}

    def get_visual_xyz_from_nn(self):                       # gm_fluid.py:1291-1336
        if not self._estimate_xyz_nn.is_cuda:
            return gm._fnx_original["get_visual_xyz_from_nn"](self)
        return physics.visual_advect(self._estimate_xyz_nn * self.scale_factor, self._xyz, self._visual_xyz.detach(), self.H, self._secs, self.KNN_K)

    def get_gas_constraints_from_exyz_nn(self):             # gm_fluid.py:1107-1132
        if not self._estimate_xyz_nn.is_cuda:
            return gm._fnx_original["get_gas_constraints_from_exyz_nn"](self)
        return physics.density_ratio(self._estimate_xyz_nn * self.scale_factor, self._imass, self.H, self.p0, self.KNN_K)

    def get_gas_constraints_from_vel_nn_guess(self):        # gm_fluid.py:1134-1158 (the next-tick map itself stays the reference's)
        if not self._estimate_xyz_nn.is_cuda:
            return gm._fnx_original["get_gas_constraints_from_vel_nn_guess"](self)
        return physics.density_ratio(self.get_guess_hidden_particles_from_nn(), self._imass, self.H, self.p0, self.KNN_K)

    gm.get_visual_xyz_from_nn = get_visual_xyz_from_nn
    gm.get_gas_constraints_from_exyz_nn = get_gas_constraints_from_exyz_nn
    gm.get_gas_constraints_from_vel_nn_guess = get_gas_constraints_from_vel_nn_guess
    gm._fnx_patched = True


def _unpatch_model(mod):
    gm = mod.GaussianModel
    if getattr(gm, "_fnx_patched", False):
        for k, v in gm._fnx_original.items():
            setattr(gm, k, v)
        gm._fnx_patched = False


def _patch_camera(mod):
    cam = mod.Camera
    if getattr(cam, "_fnx_patched", False):
        return
    init = cam.__init__

    def __init__(self, *a, **k):
        init(self, *a, **k)
        for name in ("original_image", "original_image_real"):
            img = getattr(self, name, None)
            if isinstance(img, torch.Tensor) and not img.is_cuda and not isinstance(img, CachedImage):
                setattr(self, name, CachedImage(img))

    cam.__init__ = __init__
    cam._fnx_patched = True


class _PatchLoader(importlib.abc.Loader):
    def __init__(self, loader, name):
        self.loader, self.name = loader, name

    def create_module(self, spec):
        return self.loader.create_module(spec) if hasattr(self.loader, "create_module") else None

    def exec_module(self, module):
        self.loader.exec_module(module)
        _PATCHERS[self.name](module)


class _PatchFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path=None, target=None):
        if name not in _PATCHERS:
            return None
        for finder in sys.meta_path:
            if finder is self or not hasattr(finder, "find_spec"):
                continue
            spec = finder.find_spec(name, path, target)
            if spec is not None and spec.loader is not None:
                spec.loader = _PatchLoader(spec.loader, name)
                return spec
        return None


def install_accelerators(loss_utils=True, ground_truth_cache=True, physics_terms=True):
    """Patch the reference's `utils.loss_utils`, `scene.camera` and model classes as (or if already) imported; see the module
    docstring.  Call after fluidnexus_b200.install_compat() and before the reference's entry script imports its modules."""
    if loss_utils:
        _PATCHERS["utils.loss_utils"] = _patch_loss_utils
    if ground_truth_cache:
        _PATCHERS["scene.camera"] = _patch_camera
    if physics_terms:
        _PATCHERS["gaussian_splatting.gm_fluid"] = _patch_model
        _PATCHERS["gaussian_splatting.gm_dynamics"] = _patch_model
    for name, fn in list(_PATCHERS.items()):
        if name in sys.modules:
            fn(sys.modules[name])
    if _INSTALLED["finder"] is None:
        _INSTALLED["finder"] = _PatchFinder()
        sys.meta_path.insert(0, _INSTALLED["finder"])


def uninstall_accelerators():
    """Undo install_accelerators() for modules patched so far (tests)."""
    if _INSTALLED["finder"] is not None and _INSTALLED["finder"] in sys.meta_path:
        sys.meta_path.remove(_INSTALLED["finder"])
    _INSTALLED["finder"] = None
    mod = sys.modules.get("utils.loss_utils")
    if mod is not None and hasattr(mod, "_fnx_original"):
        for k, v in mod._fnx_original.items():
            setattr(mod, k, v)
        del mod._fnx_original
    for name in ("gaussian_splatting.gm_fluid", "gaussian_splatting.gm_dynamics"):
        if name in sys.modules:
            _unpatch_model(sys.modules[name])
    _PATCHERS.clear()
