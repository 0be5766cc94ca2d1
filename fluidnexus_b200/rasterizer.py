"""Host-side mirror of the reference rasterizer interface, on top of libfnx's C ABI.

Mirrors (same names, argument meaning and error behaviour):
  R3/diff_gaussian_rasterization_ch3/__init__.py:143-154  GaussianRasterizationSettings (11 fields)
  R3/diff_gaussian_rasterization_ch3/__init__.py:157-215  GaussianRasterizer.forward / .mark_visible
  R3/diff_gaussian_rasterization_ch3/__init__.py:33-140   _RasterizeGaussians (autograd boundary, 9 grads)
  R3/rasterize_points.cu:35-215                           allocation of outputs / scratch, P == 0 short-circuit
(R1 = same with one channel.)  `make_module(C)` builds the class set for a channel count; the drop-in
packages in fluidnexus_b200/compat/ re-export them under the reference's module names.

Everything runs on torch's *current* stream and the tensors' device (the reference launches on the legacy
default stream, rasterizer_impl.cu:137,272,296 -- that breaks multi-GPU ranks and stream capture).
"""
import ctypes as C
import os
from typing import NamedTuple

import torch

from . import _lib as L

__all__ = ["make_module", "make_C", "raster_forward", "raster_backward", "mark_visible", "RasterContext", "RasterWorkspace", "MergedRasterWorkspace",
           "StaticStream"]


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _f32c(t):
    """contiguous float32 view/copy (the reference calls .contiguous().data<float>() on everything)."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _Buf:
    """A growable uint8 CUDA buffer handed to libfnx through an allocation callback
    (the reference's resizeFunctional, R3/rasterize_points.cu:27-33)."""

    def __init__(self, device):
        self.device = device
        self.t = None
        self.cb = L.ALLOC_FN(self._alloc)

    def _alloc(self, _ctx, nbytes):
        try:
            if self.t is None or self.t.numel() < nbytes:
                self.t = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            return self.t.data_ptr()
        except Exception:  # never let an exception cross the C boundary
            return None


class RasterContext:
    """Everything the backward needs from a forward (the reference keeps the same in `ctx`)."""
    __slots__ = ("args", "scratch", "bufs", "num_rendered", "radii", "keep", "C", "V", "P")


# General (single-stream) path: bin by per-tile buckets sorted in shared memory (FNX_BUCKET_BINNING) instead of two global
# radix sorts.  Results are identical.  Measured on B200: a win for small sets spread over many tiles (it is always used for
# the dynamic set of MergedRasterWorkspace), a loss for dense all-dynamic plumes (1600+ instances per tile contend on the tile's
# histogram / cursor atomics and make long sorts: scalar workload 673 -> 610 it/s, c2 1473 -> 1446 it/s -- though one c2 frame alone
# runs 1.013 -> 0.975 ms -- measured again in round 2 with the register / shuffle bucket sort), so it is off by default here.
BUCKET_BINNING = os.environ.get("FNX_BUCKET_BINNING", "0") == "1"   # (the environment variable is a measurement switch)

# running estimate of the instance count per (device, C, V, W, H): lets the forward size its binning buffers
# without blocking on the device-side count (see FNX_NO_HOST_SYNC / instance_capacity_hint in include/fnx.h)
_capacity_hint = {}


def raster_forward(C_, bg, means3D, colors, opacities, scales, rotations, scale_modifier, cov3D_precomp, view_matrix,
                   proj_matrix, tan_fov_x, tan_fov_y, H, W, sh=None, prefiltered=False, exact_rect=False,
                   speculative=True, grad_range=None, sh_degree=0, campos=None):
    """One forward through libfnx.  view_matrix/proj_matrix may be [4,4] (one camera, reference API) or
    [V,4,4] (V cameras batched into one launch sequence).  Returns (ctx, color, radii, depth) with shapes
    [C,H,W]/[P]/[1,H,W] for a single camera and [V,C,H,W]/[V,P]/[V,1,H,W] for a batch."""
    lib = L.lib()
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:56-58
    if not means3D.is_cuda:
        raise RuntimeError("libfnx needs CUDA tensors (there is no CPU fallback)")
    dev = means3D.device
    P = means3D.size(0)
    batched = view_matrix.dim() == 3
    V = view_matrix.size(0) if batched else 1
    use_sh = sh is not None and sh.numel() != 0
    if P and (colors is None or colors.numel() == 0):
        if C_ != 3:
            raise RuntimeError("For non-RGB, provide precomputed Gaussian colors!")  # rasterizer_impl.cu:226-228
        if not use_sh:
            raise RuntimeError("Please provide exactly one of either SHs or precomputed colors!")
    if use_sh:
        if sh.dim() != 3 or sh.size(0) != P or sh.size(2) != 3:
            raise RuntimeError("shs must have dimensions (num_points, num_coefficients, 3)")
        if campos is None:
            raise RuntimeError("campos is needed to evaluate SH colours")
        colors = None

    keep = dict(
        means3D=_f32c(means3D), colors=_f32c(colors), opacities=_f32c(opacities), sh=_f32c(sh) if use_sh else None,
        campos=_f32c(campos).reshape(-1, 3).expand(V, 3).contiguous() if use_sh else None,
        scales=_f32c(scales) if scales is not None and scales.numel() else None,
        rotations=_f32c(rotations) if rotations is not None and rotations.numel() else None,
        cov=_f32c(cov3D_precomp) if cov3D_precomp is not None and cov3D_precomp.numel() else None,
        view=_f32c(view_matrix), proj=_f32c(proj_matrix), bg=_f32c(bg),
    )
    if P and not use_sh and keep["colors"].numel() != P * C_:
        raise RuntimeError(f"colors_precomp must have {C_} channels per Gaussian")
    out_shape = (V, C_, H, W) if batched else (C_, H, W)
    with torch.cuda.device(dev):
        color = torch.empty(out_shape, dtype=torch.float32, device=dev)
        depth = torch.empty((V, 1, H, W) if batched else (1, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((V, P) if batched else (P,), dtype=torch.int32, device=dev)
        a = L.RasterArgs()
        a.P, a.V, a.C, a.W, a.H = P, V, C_, W, H
        a.means3D, a.colors, a.opacities = _ptr(keep["means3D"]), _ptr(keep["colors"]), _ptr(keep["opacities"])
        a.scales, a.rotations, a.cov3D_precomp, a.sh = _ptr(keep["scales"]), _ptr(keep["rotations"]), _ptr(keep["cov"]), _ptr(keep["sh"])
        if use_sh:
            a.sh_degree, a.sh_coeffs, a.campos = int(sh_degree), int(sh.size(1)), keep["campos"].data_ptr()
        a.view_matrix, a.proj_matrix, a.bg = _ptr(keep["view"]), _ptr(keep["proj"]), _ptr(keep["bg"])
        a.tan_fov_x, a.tan_fov_y, a.scale_modifier = float(tan_fov_x), float(tan_fov_y), float(scale_modifier)
        a.prefiltered = int(bool(prefiltered))
        a.flags = (L.FNX_EXACT_RECT if exact_rect else 0) | (L.FNX_BUCKET_BINNING if BUCKET_BINNING else 0)
        a.grad_begin, a.grad_end = (0, 0) if grad_range is None else (int(grad_range[0]), int(grad_range[1]))
        a.tile_order, a.static_view_map, a.static_views = None, None, 0
        key = (dev.index, C_, V, W, H, P)
        a.instance_capacity_hint = _capacity_hint.get(key, 0) if speculative else 0
        bufs = (_Buf(dev), _Buf(dev), _Buf(dev))
        scratch = L.RasterScratch()
        nr = C.c_int64(0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        fwd = lib.fnx_raster_forward_ch3 if C_ == 3 else lib.fnx_raster_forward_ch1
        L.check(fwd(C.byref(a), bufs[0].cb, None, bufs[1].cb, None, bufs[2].cb, None, color.data_ptr(), depth.data_ptr(),
                    radii.data_ptr() if P else None, C.byref(nr), C.byref(scratch), stream))
    if speculative and P:
        _capacity_hint[key] = int(nr.value * 1.25) + 4096
    ctx = RasterContext()
    ctx.args, ctx.scratch, ctx.bufs, ctx.num_rendered, ctx.radii, ctx.keep = a, scratch, bufs, int(nr.value), radii, keep
    ctx.C, ctx.V, ctx.P = C_, V, P
    return ctx, color, radii, depth


def raster_backward(ctx, dL_dout_color, want_means2D=True):
    """Backward through libfnx.  Returns a dict of gradient tensors summed over the ctx's views
    (means2D is per view: [P,3] or [V,P,3])."""
    lib = L.lib()
    P, V, C_ = ctx.P, ctx.V, ctx.C
    dev = ctx.keep["means3D"].device
    g = {}
    with torch.cuda.device(dev):
        g["means3D"] = torch.empty((P, 3), dtype=torch.float32, device=dev)
        g["means2D"] = torch.empty((V, P, 3) if V > 1 or ctx.radii.dim() == 2 else (P, 3), dtype=torch.float32, device=dev)
        use_sh = ctx.keep.get("sh") is not None
        g["colors"] = torch.zeros((P, C_), dtype=torch.float32, device=dev) if use_sh else torch.empty((P, C_), dtype=torch.float32, device=dev)
        if use_sh:
            g["sh"] = torch.empty(ctx.keep["sh"].shape, dtype=torch.float32, device=dev)
        g["opacity"] = torch.empty((P, 1), dtype=torch.float32, device=dev)
        g["cov3D"] = torch.empty((P, 6), dtype=torch.float32, device=dev)
        has_sr = ctx.keep["scales"] is not None
        # rasterize_points.cu:150-158 returns zero tensors for the paths that are not taken
        g["scales"] = torch.empty((P, 3), dtype=torch.float32, device=dev) if has_sr else torch.zeros((P, 3), device=dev)
        g["rotations"] = torch.empty((P, 4), dtype=torch.float32, device=dev) if has_sr else torch.zeros((P, 4), device=dev)
        if P == 0:
            return g
        dpix = _f32c(dL_dout_color)
        gr = L.RasterGrads()
        gr.dL_dmeans3D, gr.dL_dmeans2D = g["means3D"].data_ptr(), g["means2D"].data_ptr() if want_means2D else None
        gr.dL_dcolors, gr.dL_dopacity, gr.dL_dcov3D = None if use_sh else g["colors"].data_ptr(), g["opacity"].data_ptr(), g["cov3D"].data_ptr()
        gr.dL_dsh = g["sh"].data_ptr() if use_sh else None
        gr.dL_dscales = g["scales"].data_ptr() if has_sr else None
        gr.dL_drotations = g["rotations"].data_ptr() if has_sr else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        bwd = lib.fnx_raster_backward_ch3 if C_ == 3 else lib.fnx_raster_backward_ch1
        L.check(bwd(C.byref(ctx.args), C.byref(ctx.scratch), ctx.num_rendered, ctx.radii.data_ptr(), dpix.data_ptr(),
                    C.byref(gr), stream))
    return g


def _overflow_flag(geom):
    sc = L.RasterScratch()
    sc.geom, sc.geom_bytes = geom.data_ptr(), geom.numel()
    p = C.c_void_p()
    L.check(L.lib().fnx_raster_overflow_flag(C.byref(sc), C.byref(p)))
    return p.value


class RasterWorkspace:
    """Persistent buffers for repeated forward/backward of one fixed-size problem with NO allocation, NO host
    synchronisation and no events inside the calls (FNX_NO_HOST_SYNC), so a whole training iteration can be captured
    into a CUDA graph.  `capacity` = instances the binning buffers hold; `overflowed()` reports (after a sync) whether a
    forward needed more -- in that case that forward rendered only the background."""

    def __init__(self, dev, C_, P, V, H, W, capacity, want=("means3D",)):
        lib = L.lib()
        self.dev, self.C, self.P, self.V, self.H, self.W, self.capacity = torch.device(dev), C_, P, V, H, W, int(capacity)
        u8 = lambda n: torch.empty(int(n), dtype=torch.uint8, device=self.dev)
        with torch.cuda.device(self.dev):
            self.geom, self.image = u8(lib.fnx_raster_geom_bytes(P, V)), u8(lib.fnx_raster_image_bytes(W, H, V))
            self.binning = u8(lib.fnx_raster_binning_bytes(self.capacity, C_))
            self.color = torch.empty((V, C_, H, W), device=self.dev)
            self.depth = torch.empty((V, 1, H, W), device=self.dev)
            self.radii = torch.empty((V, P), dtype=torch.int32, device=self.dev)
            shapes = dict(means3D=(P, 3), means2D=(V, P, 3), colors=(P, C_), opacity=(P, 1), scales=(P, 3), rotations=(P, 4),
                          cov3D=(P, 6))
            self.grads = {k: torch.empty(shapes[k], device=self.dev) for k in want}
            # start order of the (view, tile) units for the blend kernels; identity until update_tile_order() ranks them by work
            self.tile_order = torch.arange(V * ((W + 15) // 16) * ((H + 15) // 16), dtype=torch.int32, device=self.dev)
        self.count = torch.full((1,), -1, dtype=torch.int64).pin_memory()
        self._cbs = tuple(L.ALLOC_FN(self._fixed(t)) for t in (self.geom, self.binning, self.image))
        self.args, self.scratch, self._keep = L.RasterArgs(), L.RasterScratch(), None
        self.forwards = 0

    @staticmethod
    def _fixed(t):
        def alloc(_ctx, nbytes):
            return t.data_ptr() if nbytes <= t.numel() else None
        return alloc

    def forward(self, bg, means3D, colors, opacities, scales, rotations, scale_modifier, view_matrix, proj_matrix, tan_fov_x,
                tan_fov_y, exact_rect=False, grad_range=None):
        """All tensors must be contiguous float32 CUDA tensors that stay alive and in place (graph replays read them)."""
        a = self.args
        a.P, a.V, a.C, a.W, a.H = self.P, self.V, self.C, self.W, self.H
        a.means3D, a.colors, a.opacities = means3D.data_ptr(), colors.data_ptr(), opacities.data_ptr()
        a.scales, a.rotations, a.cov3D_precomp, a.sh = scales.data_ptr(), rotations.data_ptr(), None, None
        a.view_matrix, a.proj_matrix, a.bg = view_matrix.data_ptr(), proj_matrix.data_ptr(), bg.data_ptr()
        a.tan_fov_x, a.tan_fov_y, a.scale_modifier = float(tan_fov_x), float(tan_fov_y), float(scale_modifier)
        a.prefiltered, a.flags = 0, L.FNX_NO_HOST_SYNC | (L.FNX_EXACT_RECT if exact_rect else 0) | (L.FNX_BUCKET_BINNING if BUCKET_BINNING else 0)
        a.instance_capacity_hint, a.num_rendered_pinned = self.capacity, self.count.data_ptr()
        a.grad_begin, a.grad_end = (0, 0) if grad_range is None else (int(grad_range[0]), int(grad_range[1]))
        a.tile_order = self.tile_order.data_ptr()
        self._keep = (bg, means3D, colors, opacities, scales, rotations, view_matrix, proj_matrix)
        nr = C.c_int64(0)
        fwd = L.lib().fnx_raster_forward_ch3 if self.C == 3 else L.lib().fnx_raster_forward_ch1
        L.check(fwd(C.byref(a), self._cbs[0], None, self._cbs[1], None, self._cbs[2], None, self.color.data_ptr(),
                    self.depth.data_ptr(), self.radii.data_ptr(), C.byref(nr), C.byref(self.scratch),
                    torch.cuda.current_stream(self.dev).cuda_stream))
        self.forwards += 1
        return self.color

    def update_tile_order(self):
        """Rank the (view, tile) units by the work of the last forward, longest first (fnx_raster_tile_order): later forward /
        backward calls start their CTAs in that order.  Enqueued on the current stream; scheduling only, results are unchanged."""
        L.check(L.lib().fnx_raster_tile_order(C.byref(self.scratch), None, self.W, self.H, self.V, self.tile_order.data_ptr(),
                                              torch.cuda.current_stream(self.dev).cuda_stream))

    def backward(self, dL_dout_color):
        gr = L.RasterGrads()
        for name, field in (("means3D", "dL_dmeans3D"), ("means2D", "dL_dmeans2D"), ("colors", "dL_dcolors"), ("opacity", "dL_dopacity"),
                            ("scales", "dL_dscales"), ("rotations", "dL_drotations"), ("cov3D", "dL_dcov3D")):
            setattr(gr, field, self.grads[name].data_ptr() if name in self.grads else None)
        bwd = L.lib().fnx_raster_backward_ch3 if self.C == 3 else L.lib().fnx_raster_backward_ch1
        L.check(bwd(C.byref(self.args), C.byref(self.scratch), -1, self.radii.data_ptr(), dL_dout_color.data_ptr(), C.byref(gr),
                    torch.cuda.current_stream(self.dev).cuda_stream))
        return self.grads

    def overflow_flag(self):
        """Device address (int) of the forward's "instance capacity overflowed" flag (fnx_raster_overflow_flag)."""
        return _overflow_flag(self.geom)

    def num_rendered(self):
        """Instance count of the last finished forward (call after a synchronisation point)."""
        return int(self.count[0])

    def overflowed(self):
        return self.num_rendered() > self.capacity


def _fill_args(a, C_, P, V, H, W, bg, means3D, colors, opacities, scales, rotations, scale_modifier, view_matrix, proj_matrix,
               tan_fov_x, tan_fov_y, flags, capacity=0, pinned=None):
    a.P, a.V, a.C, a.W, a.H = P, V, C_, W, H
    a.means3D, a.colors, a.opacities = means3D.data_ptr(), colors.data_ptr(), opacities.data_ptr()
    a.scales, a.rotations, a.cov3D_precomp, a.sh = scales.data_ptr(), rotations.data_ptr(), None, None
    a.view_matrix, a.proj_matrix, a.bg = view_matrix.data_ptr(), proj_matrix.data_ptr(), bg.data_ptr()
    a.tan_fov_x, a.tan_fov_y, a.scale_modifier = float(tan_fov_x), float(tan_fov_y), float(scale_modifier)
    a.prefiltered, a.flags = 0, flags
    a.instance_capacity_hint, a.num_rendered_pinned = int(capacity), (pinned.data_ptr() if pinned is not None else None)
    a.grad_begin, a.grad_end = 0, 0
    a.tile_order, a.static_view_map, a.static_views = None, None, 0
    return a


class StaticStream:
    """The frozen Gaussian set of a frame (FD/renderer/pipe_dynamics.py:51-57: the background set concatenated behind the fluid
    particles), binned ONCE for all `V` cameras of the frame: its depth-sorted, tile-partitioned record stream stays in HBM and,
    with prepare=True, so does its static-only render (color / depth [V,...]) and how deep that render reaches per tile.
    Any number of MergedRasterWorkspaces -- one per subset of the cameras that an iteration renders; the reference draws
    random.sample(cur_viewpoint_set, batch) -- share it through `view_ids`."""

    def __init__(self, dev, V, H, W, bg, static, view_matrix, proj_matrix, tan_fov_x, tan_fov_y, prepare=True):
        lib = L.lib()
        self.dev, self.V, self.H, self.W = torch.device(dev), V, H, W
        self.P = static["means3D"].size(0)
        self.cam = (bg, view_matrix, proj_matrix, float(tan_fov_x), float(tan_fov_y))
        self._static = static
        self.prepared = bool(prepare)
        st = torch.cuda.current_stream(self.dev).cuda_stream
        with torch.cuda.device(self.dev):
            self._bufs = (_Buf(self.dev), _Buf(self.dev), _Buf(self.dev))
            self.args, self.scratch = L.RasterArgs(), L.RasterScratch()
            self.radii = torch.empty((V, self.P), dtype=torch.int32, device=self.dev)
            _fill_args(self.args, 3, self.P, V, H, W, bg, static["means3D"], static["colors"], static["opacities"], static["scales"],
                       static["rotations"], 1.0, view_matrix, proj_matrix, tan_fov_x, tan_fov_y, L.FNX_BIN_ONLY | L.FNX_ALL_FROZEN)
            nr = C.c_int64(0)
            L.check(lib.fnx_raster_forward_ch3(C.byref(self.args), self._bufs[0].cb, None, self._bufs[1].cb, None, self._bufs[2].cb, None,
                                               None, None, self.radii.data_ptr(), C.byref(nr), C.byref(self.scratch), st))
            self.R = int(nr.value)
            nt = ((W + 15) // 16) * ((H + 15) // 16)
            self.color = self.depth = None
            self.reach = None            # [V] static records per view that a blend can reach (sum over tiles)
            if prepare:
                self.color = torch.empty((V, 3, H, W), device=self.dev)
                self.depth = torch.empty((V, 1, H, W), device=self.dev)
                L.check(lib.fnx_raster_static_prepare(C.byref(self.args), C.byref(self.scratch), self.color.data_ptr(), self.depth.data_ptr(), st))
                last = torch.zeros((V * nt, 4), dtype=torch.int32, device=self.dev)  # per 8x8 patch
                L.check(lib.fnx_raster_read_tiles(C.byref(self.scratch), W, H, V, 0, None, last.data_ptr(), None, None, st))
                self.reach = last.max(dim=1).values.view(V, nt).sum(dim=1).cpu()


class MergedRasterWorkspace:
    """Rasterizer workspace for [dynamic ; static] Gaussian sets (FD/renderer/pipe_dynamics.py:51-57 concatenates the
    moving fluid particles with a frozen background set).  The static set's record stream lives in a StaticStream (built
    here for exactly these cameras, or shared: `static_stream` + `view_ids` = which of its cameras this workspace renders);
    every forward bins only the dynamic set and merges the two streams per tile (fnx_raster_blend_merged).  Results equal
    one forward over the concatenated set; gradients are produced for the dynamic Gaussians only.  3 channels.  No
    allocation / host sync / events inside forward() and backward() (CUDA-graph capturable)."""

    def __init__(self, dev, P_dyn, V, H, W, bg, dyn, static, view_matrix, proj_matrix, tan_fov_x, tan_fov_y, margin=1.2,
                 static_prepare=True, static_tile_cache=True, bucket_binning=True, static_stream=None, view_ids=None, slack=65536,
                 want=("means3D",)):
        """dyn / static: dicts of contiguous float32 CUDA tensors means3D, colors, opacities, scales, rotations (static may be
        None when a static_stream is given).  view_matrix / proj_matrix: the V cameras this workspace renders.
        static_prepare (private stream only): blend the static stream alone once, which bounds the static records every later
        merge has to copy; static_tile_cache (needs a prepared stream): tiles without a dynamic instance keep their static-only
        pixels in self.color / self.depth instead of being re-blended every forward."""
        lib = L.lib()
        self.dev, self.C, self.P, self.V, self.H, self.W = torch.device(dev), 3, P_dyn, V, H, W
        if static_stream is None:
            static_stream = StaticStream(dev, V, H, W, bg, static, view_matrix, proj_matrix, tan_fov_x, tan_fov_y, prepare=static_prepare)
            view_ids = list(range(V))
        assert view_ids is not None and len(view_ids) == V and static_stream.H == H and static_stream.W == W
        self.static = static_stream
        self.view_ids = [int(v) for v in view_ids]
        self.P_static, self.R_static = static_stream.P, static_stream.R
        self.sscratch = static_stream.scratch
        identity = self.view_ids == list(range(static_stream.V))
        self.cam = (bg, view_matrix, proj_matrix, float(tan_fov_x), float(tan_fov_y))
        self.static_tile_cache = bool(static_tile_cache and static_stream.prepared)
        self.bucket_binning = bool(bucket_binning)  # per-tile buckets sorted inside the merge instead of two global radix sorts
        st = torch.cuda.current_stream(self.dev).cuda_stream
        with torch.cuda.device(self.dev):
            self.view_map = None if identity else torch.tensor(self.view_ids, dtype=torch.int32, device=self.dev)
            if static_stream.prepared:   # start from the static-only render of these cameras: the tile cache's pixels
                idx = torch.tensor(self.view_ids, dtype=torch.long, device=self.dev)
                self.color, self.depth = static_stream.color[idx].contiguous(), static_stream.depth[idx].contiguous()
            else:
                self.color = torch.empty((V, 3, H, W), device=self.dev)
                self.depth = torch.empty((V, 1, H, W), device=self.dev)
            # ---- dynamic stream: one exact binning to size the capacity ----
            tmp = (_Buf(self.dev), _Buf(self.dev), _Buf(self.dev))
            targs, tscratch = L.RasterArgs(), L.RasterScratch()
            nr = C.c_int64(0)
            self.radii = torch.empty((V, P_dyn), dtype=torch.int32, device=self.dev)
            _fill_args(targs, 3, P_dyn, V, H, W, bg, dyn["means3D"], dyn["colors"], dyn["opacities"], dyn["scales"], dyn["rotations"],
                       1.0, view_matrix, proj_matrix, tan_fov_x, tan_fov_y, L.FNX_BIN_ONLY)
            L.check(lib.fnx_raster_forward_ch3(C.byref(targs), tmp[0].cb, None, tmp[1].cb, None, tmp[2].cb, None, None, None,
                                               self.radii.data_ptr(), C.byref(nr), C.byref(tscratch), st))
            torch.cuda.synchronize(self.dev)
            self.capacity = int(nr.value * margin) + int(slack)   # dynamic instances the binning buffers hold
            del tmp
            u8 = lambda n: torch.empty(int(n), dtype=torch.uint8, device=self.dev)
            self.geom, self.image = u8(lib.fnx_raster_geom_bytes(P_dyn, V)), u8(lib.fnx_raster_image_bytes(W, H, V))
            self.binning = u8(lib.fnx_raster_binning_bytes(self.capacity, 3))
            # merged spans hold the tile's dynamic records + the static records a blend can reach: after the static-only
            # blend that is sum(tile_last) static records of these cameras instead of all of them
            n_static_reach = int(static_stream.reach[self.view_ids].sum()) if static_stream.prepared else self.R_static
            self.merged = u8(48 * (self.capacity + n_static_reach) + 256)
            # gradients of the DYNAMIC Gaussians the backward writes (the physical stage needs the means only, level two the attributes)
            shapes = dict(means3D=(P_dyn, 3), means2D=(V, P_dyn, 3), colors=(P_dyn, 3), opacity=(P_dyn, 1), scales=(P_dyn, 3), rotations=(P_dyn, 4),
                          cov3D=(P_dyn, 6))
            self.grads = {k: torch.empty(shapes[k], device=self.dev) for k in want}
            self.tile_order = torch.arange(V * ((W + 15) // 16) * ((H + 15) // 16), dtype=torch.int32, device=self.dev)
            # which tiles of self.color / self.depth hold their static-only pixels: all of them (copied above) or none
            sc = L.RasterScratch()
            sc.image, sc.image_bytes = self.image.data_ptr(), self.image.numel()
            L.check(lib.fnx_raster_tile_cache_set(C.byref(sc), W, H, V, int(static_stream.prepared), st))
        self.forwards = 0
        self.count = torch.full((1,), -1, dtype=torch.int64).pin_memory()
        self._cbs = tuple(L.ALLOC_FN(RasterWorkspace._fixed(t)) for t in (self.geom, self.binning, self.image))
        self.args, self.scratch, self._keep = L.RasterArgs(), L.RasterScratch(), None

    def forward(self, means3D, colors, opacities, scales, rotations):
        lib = L.lib()
        bg, vm, pm, tfx, tfy = self.cam
        _fill_args(self.args, 3, self.P, self.V, self.H, self.W, bg, means3D, colors, opacities, scales, rotations, 1.0, vm, pm, tfx, tfy,
                   L.FNX_BIN_ONLY | L.FNX_NO_HOST_SYNC | (L.FNX_STATIC_TILE_CACHE if self.static_tile_cache else 0)
                   | (L.FNX_BUCKET_BINNING if self.bucket_binning else 0), self.capacity, self.count)
        self.args.tile_order = self.tile_order.data_ptr()
        self.args.static_view_map = None if self.view_map is None else self.view_map.data_ptr()
        self.args.static_views = self.static.V
        self._keep = (means3D, colors, opacities, scales, rotations)
        st = torch.cuda.current_stream(self.dev).cuda_stream
        nr = C.c_int64(0)
        L.check(lib.fnx_raster_forward_ch3(C.byref(self.args), self._cbs[0], None, self._cbs[1], None, self._cbs[2], None, None, None,
                                           self.radii.data_ptr(), C.byref(nr), C.byref(self.scratch), st))
        L.check(lib.fnx_raster_blend_merged(C.byref(self.args), C.byref(self.scratch), C.byref(self.sscratch), self.P_static,
                                            self.merged.data_ptr(), self.color.data_ptr(), self.depth.data_ptr(), st))
        self.forwards += 1
        return self.color

    def update_tile_order(self):
        """See RasterWorkspace.update_tile_order; tiles served from the static tile cache rank last."""
        L.check(L.lib().fnx_raster_tile_order(C.byref(self.scratch), C.byref(self.sscratch), self.W, self.H, self.V, self.tile_order.data_ptr(),
                                              torch.cuda.current_stream(self.dev).cuda_stream))

    def backward(self, dL_dout_color):
        gr = L.RasterGrads()
        for name, field in (("means3D", "dL_dmeans3D"), ("means2D", "dL_dmeans2D"), ("colors", "dL_dcolors"), ("opacity", "dL_dopacity"),
                            ("scales", "dL_dscales"), ("rotations", "dL_drotations"), ("cov3D", "dL_dcov3D")):
            setattr(gr, field, self.grads[name].data_ptr() if name in self.grads else None)
        L.check(L.lib().fnx_raster_backward_merged(C.byref(self.args), C.byref(self.scratch), C.byref(self.sscratch), self.merged.data_ptr(),
                                                   self.radii.data_ptr(), dL_dout_color.data_ptr(), C.byref(gr),
                                                   torch.cuda.current_stream(self.dev).cuda_stream))
        return self.grads

    def tile_state(self):
        """Per-tile state of the last forward as numpy arrays [V, tiles] (synchronises): merged span begin/end, records the
        forward blended (tile_last), tile_src (1: static-only tile) and tile_dyn_last (where the backward starts)."""
        nt = self.V * ((self.W + 15) // 16) * ((self.H + 15) // 16)
        i32 = lambda *s: torch.zeros(s, dtype=torch.int32, device=self.dev)
        ranges, last, src, dyn = i32(nt, 2), i32(nt, 4), i32(nt), i32(nt)
        L.check(L.lib().fnx_raster_read_tiles(C.byref(self.scratch), self.W, self.H, self.V, 1, ranges.data_ptr(), last.data_ptr(),
                                              src.data_ptr(), dyn.data_ptr(), torch.cuda.current_stream(self.dev).cuda_stream))
        torch.cuda.synchronize(self.dev)
        r = ranges.cpu().numpy().astype("int64")
        pl = last.cpu().numpy().astype("int64")   # per 8x8 patch of the tile
        return dict(begin=r[:, 0], end=r[:, 1], tile_last=pl.max(axis=1), patch_last=pl, tile_src=src.cpu().numpy(),
                    tile_dyn_last=dyn.cpu().numpy().astype("int64"))

    def overflow_flag(self):
        return _overflow_flag(self.geom)

    def num_rendered(self):
        """Dynamic instances of the last finished forward + the static stream's instances."""
        n = int(self.count[0])
        return n + self.R_static if n >= 0 else n

    def overflowed(self):
        return int(self.count[0]) > self.capacity


def mark_visible(positions, view_matrix, proj_matrix):
    lib = L.lib()
    pos = _f32c(positions)
    P = pos.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=pos.device)
    if P:
        with torch.cuda.device(pos.device):
            L.check(lib.fnx_mark_visible(P, pos.data_ptr(), _f32c(view_matrix).data_ptr(), _f32c(proj_matrix).data_ptr(),
                                         present.data_ptr(), torch.cuda.current_stream(pos.device).cuda_stream))
    return present


def make_module(C_):
    """Build (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians, _C-like namespace)
    for a channel count, with the reference's names and call signatures."""

    class GaussianRasterizationSettings(NamedTuple):
        image_height: int
        image_width: int
        tan_fov_x: float
        tan_fov_y: float
        bg: torch.Tensor
        scale_modifier: float
        view_matrix: torch.Tensor
        proj_matrix: torch.Tensor
        sh_degree: int
        campos: torch.Tensor
        prefiltered: bool

    class _RasterizeGaussians(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
            fctx, color, radii, depth = raster_forward(
                C_, rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                rs.view_matrix, rs.proj_matrix, rs.tan_fov_x, rs.tan_fov_y, rs.image_height, rs.image_width, sh=sh,
                prefiltered=rs.prefiltered, sh_degree=rs.sh_degree, campos=rs.campos)
            ctx.fctx = fctx
            ctx.mark_non_differentiable(radii, depth)
            return color, radii, depth

        @staticmethod
        def backward(ctx, grad_out_color, _grad_radii, _grad_depth):
            g = raster_backward(ctx.fctx, grad_out_color)
            has_cov = ctx.fctx.keep["cov"] is not None
            # order of R3/diff_gaussian_rasterization_ch3/__init__.py:128-138
            return (g["means3D"], g["means2D"], g.get("sh"), g["colors"], g["opacity"],
                    None if has_cov else g["scales"], None if has_cov else g["rotations"],
                    g["cov3D"] if has_cov else None, None)

    def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
        return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                         cov3Ds_precomp, rs)

    class GaussianRasterizer(torch.nn.Module):
        def __init__(self, raster_settings):
            super().__init__()
            self.raster_settings = raster_settings

        def mark_visible(self, positions):
            with torch.no_grad():
                rs = self.raster_settings
                return mark_visible(positions, rs.view_matrix, rs.proj_matrix)

        def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3D_precomp=None):
            rs = self.raster_settings
            if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
                raise Exception("Please provide exactly one of either SHs or precomputed colors!")
            if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                    (scales is not None or rotations is not None) and cov3D_precomp is not None):
                raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
            empty = torch.Tensor([])
            return rasterize_gaussians(
                means3D, means2D, empty if shs is None else shs, empty if colors_precomp is None else colors_precomp,
                opacities, empty if scales is None else scales, empty if rotations is None else rotations,
                empty if cov3D_precomp is None else cov3D_precomp, rs)

    return GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians, _RasterizeGaussians


def make_C(C_):
    """The reference's pybind module `_C` (R3/ext.cpp:15-19; signatures R3/rasterize_points.h:18-64) for a channel count, on
    libfnx: (rasterize_gaussians, rasterize_gaussians_backward, mark_visible) with the reference's positional arguments, return
    tuples and buffer hand-over, so that the reference's OWN wrapper package (`diff_gaussian_rasterization_ch3/__init__.py`:
    `from . import _C`) runs unchanged on top of it.  geomBuffer / binningBuffer / imgBuffer are the three uint8 scratch tensors of
    the forward (private layout, like the reference's); the backward gets them back together with `R` and rebuilds the scratch
    handle from them.  The forward sizes its binning buffer exactly (one host sync, like rasterizer_impl.cu:263-264), so its
    capacity IS the returned instance count."""

    def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, view_matrix,
                            proj_matrix, tan_fov_x, tan_fov_y, image_height, image_width, sh, degree, campos, prefiltered):
        ctx, color, radii, depth = raster_forward(
            C_, background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, view_matrix, proj_matrix,
            tan_fov_x, tan_fov_y, int(image_height), int(image_width), sh=sh, prefiltered=prefiltered, speculative=False,
            sh_degree=int(degree), campos=campos)
        empty = lambda: torch.empty(0, dtype=torch.uint8, device=means3D.device)
        geom, binning, img = (b.t if b.t is not None else empty() for b in ctx.bufs)
        return ctx.num_rendered, color, radii, geom, binning, img, depth

    def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, view_matrix,
                                     proj_matrix, tan_fov_x, tan_fov_y, dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer,
                                     imageBuffer):
        if means3D.dim() != 2 or means3D.size(1) != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        P = means3D.size(0)
        H, W = int(dL_dout_color.size(-2)), int(dL_dout_color.size(-1))
        use_sh = sh is not None and sh.numel() != 0
        some = lambda t: t is not None and t.numel() != 0
        keep = dict(means3D=_f32c(means3D), colors=None if use_sh else _f32c(colors), opacities=None, sh=_f32c(sh) if use_sh else None,
                    campos=_f32c(campos).reshape(-1, 3)[:1].contiguous() if use_sh else None,
                    scales=_f32c(scales) if some(scales) else None, rotations=_f32c(rotations) if some(rotations) else None,
                    cov=_f32c(cov3D_precomp) if some(cov3D_precomp) else None,
                    view=_f32c(view_matrix), proj=_f32c(proj_matrix), bg=_f32c(background),
                    buffers=(geomBuffer, binningBuffer, imageBuffer))
        a = L.RasterArgs()
        a.P, a.V, a.C, a.W, a.H = P, 1, C_, W, H
        a.means3D, a.colors = _ptr(keep["means3D"]), _ptr(keep["colors"])
        a.opacities = None   # the reference's backward is not handed them; libfnx's reads them from the forward's record stream
        a.scales, a.rotations, a.cov3D_precomp, a.sh = _ptr(keep["scales"]), _ptr(keep["rotations"]), _ptr(keep["cov"]), _ptr(keep["sh"])
        if use_sh:
            a.sh_degree, a.sh_coeffs, a.campos = int(degree), int(sh.size(1)), keep["campos"].data_ptr()
        a.view_matrix, a.proj_matrix, a.bg = _ptr(keep["view"]), _ptr(keep["proj"]), _ptr(keep["bg"])
        a.tan_fov_x, a.tan_fov_y, a.scale_modifier = float(tan_fov_x), float(tan_fov_y), float(scale_modifier)
        a.prefiltered, a.flags, a.instance_capacity_hint = 0, 0, 0
        a.grad_begin, a.grad_end = 0, 0
        a.tile_order, a.static_view_map, a.static_views = None, None, 0
        sc = L.RasterScratch()
        sc.geom, sc.geom_bytes = _ptr(geomBuffer), geomBuffer.numel()
        sc.binning, sc.binning_bytes = _ptr(binningBuffer), binningBuffer.numel()
        sc.image, sc.image_bytes = _ptr(imageBuffer), imageBuffer.numel()
        sc.binning_capacity, sc.check_slot = int(R), -1
        ctx = RasterContext()
        ctx.args, ctx.scratch, ctx.bufs, ctx.num_rendered, ctx.radii, ctx.keep = a, sc, None, int(R), radii, keep
        ctx.C, ctx.V, ctx.P = C_, 1, P
        g = raster_backward(ctx, dL_dout_color)
        dsh = g["sh"] if use_sh else torch.zeros((P, 0, 3), dtype=torch.float32, device=means3D.device)
        # order of R3/rasterize_points.cu:193
        return g["means2D"], g["colors"], g["opacity"], g["means3D"], g["cov3D"], dsh, g["scales"], g["rotations"]

    return rasterize_gaussians, rasterize_gaussians_backward, mark_visible


def read_geom(ctx):
    """Intermediate per-(view, Gaussian) state of a forward (parity tests): xy, depth, conic_opacity, tiles_touched."""
    dev = ctx.keep["means3D"].device
    P, V = ctx.P, ctx.V
    out = dict(xy=torch.zeros((V, P, 2), device=dev), depth=torch.zeros((V, P), device=dev),
               conic_opacity=torch.zeros((V, P, 4), device=dev),
               tiles_touched=torch.zeros((V, P), dtype=torch.int32, device=dev))
    if P:
        L.check(L.lib().fnx_raster_read_geom(C.byref(ctx.scratch), P, V, out["xy"].data_ptr(), out["depth"].data_ptr(),
                                             out["conic_opacity"].data_ptr(), out["tiles_touched"].data_ptr(),
                                             torch.cuda.current_stream(dev).cuda_stream))
    return out


def read_image_state(ctx):
    dev = ctx.keep["means3D"].device
    W, H, V = ctx.args.W, ctx.args.H, ctx.V
    out = dict(final_T=torch.zeros((V, H, W), device=dev), n_contrib=torch.zeros((V, H, W), dtype=torch.int32, device=dev))
    if ctx.P:
        L.check(L.lib().fnx_raster_read_image(C.byref(ctx.scratch), W, H, V, out["final_T"].data_ptr(),
                                              out["n_contrib"].data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return out
