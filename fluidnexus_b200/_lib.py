"""ctypes binding of libfnx.so (include/fnx.h).  The library is the product; there is no fallback.

`lib()` raises if the shared library is missing -- build it with `python __graft_entry__.py build`
(or `fluidnexus_b200.build.build()`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FNX_LIBFNX") or os.path.join(_HERE, "libfnx.so")   # (the override is a development switch: A/B builds)

FNX_OK = 0
FNX_ERR_INVALID, FNX_ERR_CUDA, FNX_ERR_UNSUPPORTED, FNX_ERR_CAPACITY, FNX_ERR_ALLOC = 1, 2, 3, 4, 5
FNX_NO_HOST_SYNC = 1
FNX_EXACT_RECT = 2
FNX_BIN_ONLY = 4
FNX_ALL_FROZEN = 8
FNX_STATIC_TILE_CACHE = 16
FNX_BUCKET_BINNING = 32

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class RasterArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("V", C.c_int32), ("C", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
        ("means3D", C.c_void_p), ("colors", C.c_void_p), ("opacities", C.c_void_p), ("scales", C.c_void_p),
        ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p), ("sh", C.c_void_p),
        ("view_matrix", C.c_void_p), ("proj_matrix", C.c_void_p), ("bg", C.c_void_p),
        ("tan_fov_x", C.c_float), ("tan_fov_y", C.c_float), ("scale_modifier", C.c_float),
        ("prefiltered", C.c_int32), ("flags", C.c_uint32), ("instance_capacity_hint", C.c_int64),
        ("num_rendered_pinned", C.c_void_p), ("grad_begin", C.c_int32), ("grad_end", C.c_int32),
        ("tile_order", C.c_void_p), ("static_view_map", C.c_void_p), ("static_views", C.c_int32), ("sh_degree", C.c_int32),
        ("sh_coeffs", C.c_int32), ("reserved0", C.c_int32), ("campos", C.c_void_p),
    ]


class RasterScratch(C.Structure):
    _fields_ = [
        ("geom", C.c_void_p), ("geom_bytes", C.c_size_t), ("binning", C.c_void_p), ("binning_bytes", C.c_size_t),
        ("image", C.c_void_p), ("image_bytes", C.c_size_t), ("binning_capacity", C.c_int64),
        ("check_slot", C.c_int32), ("reserved", C.c_int32),
    ]


class RasterGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("dL_dmeans3D", "dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dcov3D", "dL_dsh")]


class FnxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfnx error {code}: {msg}")
        self.code = code


# every symbol include/fnx.h declares: name -> (restype, argtypes)
_V, _I, _I64, _SZ, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_float


class GsState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("xyz", "color", "opacity", "scaling", "rotation", "m_xyz", "v_xyz", "m_color", "v_color",
                                         "m_opacity", "v_opacity", "m_scaling", "v_scaling", "m_rotation", "v_rotation",
                                         "max_radii2D", "xyz_gradient_accum", "denom")]


class GsGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dL_dmeans3D", "dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dscales", "dL_drotations")]


class GsHparams(C.Structure):
    _fields_ = [("lr_xyz", C.c_float), ("lr_color", C.c_float), ("lr_opacity", C.c_float), ("lr_scaling", C.c_float),
                ("lr_rotation", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("step", C.c_int32),
                ("update_stats", C.c_int32), ("lambda_reg_scaling", C.c_float), ("reg_ratio_threshold", C.c_float)]
class GsLevelTwo(C.Structure):
    _fields_ = [("prev_color", C.c_void_p), ("prev_opacity", C.c_void_p), ("prev_scales", C.c_void_p), ("prev_rotation", C.c_void_p),
                ("prev_num", C.c_int32), ("color_channels", C.c_int32), ("fit_color", C.c_int32), ("fit_opacity", C.c_int32),
                ("fit_scales", C.c_int32), ("fit_rotation", C.c_int32), ("lambda_consistency_color", C.c_float),
                ("lambda_consistency_opacity", C.c_float), ("lambda_consistency_scales", C.c_float), ("lambda_consistency_rotation", C.c_float),
                ("lambda_reg_scaling", C.c_float), ("reg_ratio_threshold", C.c_float), ("lr_color", C.c_float), ("lr_opacity", C.c_float),
                ("lr_scaling", C.c_float), ("lr_rotation", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("step", C.c_int32)]


_FWD = (_I, [C.POINTER(RasterArgs), ALLOC_FN, _V, ALLOC_FN, _V, ALLOC_FN, _V, _V, _V, _V, C.POINTER(_I64),
             C.POINTER(RasterScratch), _V])
_BWD = (_I, [C.POINTER(RasterArgs), C.POINTER(RasterScratch), _I64, _V, _V, C.POINTER(RasterGrads), _V])
SYMBOLS = {
    "fnx_abi_version": (_I, []),
    "fnx_last_error": (C.c_char_p, []),
    "fnx_build_arch": (C.c_char_p, []),
    "fnx_launch_count": (C.c_uint64, []),
    "fnx_raster_geom_bytes": (_SZ, [_I, _I]),
    "fnx_raster_image_bytes": (_SZ, [_I, _I, _I]),
    "fnx_raster_binning_bytes": (_SZ, [_I64, _I]),
    "fnx_raster_forward": _FWD, "fnx_raster_forward_ch1": _FWD, "fnx_raster_forward_ch3": _FWD,
    "fnx_raster_backward": _BWD, "fnx_raster_backward_ch1": _BWD, "fnx_raster_backward_ch3": _BWD,
    "fnx_raster_blend_merged": (_I, [C.POINTER(RasterArgs), C.POINTER(RasterScratch), C.POINTER(RasterScratch), _I, _V, _V, _V, _V]),
    "fnx_raster_static_prepare": (_I, [C.POINTER(RasterArgs), C.POINTER(RasterScratch), _V, _V, _V]),
    "fnx_raster_read_tiles": (_I, [C.POINTER(RasterScratch), _I, _I, _I, _I, _V, _V, _V, _V, _V]),
    "fnx_raster_backward_merged": (_I, [C.POINTER(RasterArgs), C.POINTER(RasterScratch), C.POINTER(RasterScratch), _V, _V, _V,
                                        C.POINTER(RasterGrads), _V]),
    "fnx_raster_overflow_flag": (_I, [C.POINTER(RasterScratch), C.POINTER(_V)]),
    "fnx_raster_tile_cache_set": (_I, [C.POINTER(RasterScratch), _I, _I, _I, _I, _V]),
    "fnx_raster_tile_order": (_I, [C.POINTER(RasterScratch), C.POINTER(RasterScratch), _I, _I, _I, _V, _V]),
    "fnx_raster_check": (_I, [C.POINTER(RasterScratch), C.POINTER(_I64), _V]),
    "fnx_mark_visible": (_I, [_I, _V, _V, _V, _V, _V]),
    "fnx_grid_bytes": (_SZ, [_I]),
    "fnx_grid_build": (_I, [_V, _I, _F, _V, _V]),
    "fnx_radius_count": (_I, [_V, _I, _F, _V, _I, _F, _I, _V, _V, _V]),
    "fnx_radius_fill": (_I, [_V, _I, _F, _V, _I, _F, _V, _V, _V, _V, _V]),
    "fnx_pbf_density_fwd": (_I, [_V, _V, _I, _V, _V, _F, _F, _V, _V]),
    "fnx_gs_activate": (_I, [_I, _V, _V, _V, _V, _V, _V, _V]),
    "fnx_gs_update_level_two": (_I, [_I, _I, C.POINTER(GsState), C.POINTER(GsGrads), C.POINTER(GsLevelTwo), _V, _V]),
    "fnx_gs_update": (_I, [_I, _I, C.POINTER(GsState), C.POINTER(GsGrads), C.POINTER(GsHparams), _V, _V, _V]),
    "fnx_pbf_guess_hidden": (_I, [_I, _V, _V, _V, _V, _V, _V, C.POINTER(_F), _F, _F, _F, _F, _F, _I, C.POINTER(_F), _F, _F, _V]),
    "fnx_pbf_project_gas_constraints": (_I, [_V, _V, _I, _V, _V, _V, _V, _F, _F, _F, _I, _F, _F, _I, _F, _V, _V, _V, _V, _V]),
    "fnx_radius_graph_degree": (_I, [_V, _V, _I, _F, _I, _I, _V, _V, _V]),
    "fnx_rigid_project": (_I, [_V, _V, _I, _V, _I, _I, C.POINTER(_F), C.POINTER(_F), _F, _I, _V, _V]),
    "fnx_pbf_confirm_guess": (_I, [_I, _V, _V, _V, _F, _V]),
    "fnx_pbf_update_visual": (_I, [_V, _V, _V, _I, _V, _I, _I, _F, _F, _V, _V]),
    "fnx_pbf_density_fwd_counted": (_I, [_V, _V, _I, _V, _I, _F, _F, _V, _V, _V, _V]),
    "fnx_pbf_density_bwd": (_I, [_V, _V, _I, _V, _V, _F, _F, _V, _V, _I, _V]),
    "fnx_pbf_density_bwd_ratio": (_I, [_V, _V, _I, _V, _V, _F, _F, _V, _F, _V, _V, _V, _I, _V]),
    "fnx_visual_advect_fwd": (_I, [_V, _V, _V, _I, _V, _I, _V, _F, _F, _F, _V, _V, _V, _V]),
    "fnx_visual_advect_fwd_counted": (_I, [_V, _V, _V, _I, _V, _I, _I, _F, _F, _F, _V, _V, _V, _V, _V]),
    "fnx_visual_advect_bwd": (_I, [_V, _V, _V, _I, _I, _V, _V, _V, _V, _V, _F, _F, _F, _V, _I, _V]),
    "fnx_pair_distance_loss": (_I, [_V, _V, _I, _F, _F, _F, _V, _V, _V]),
    "fnx_knn3_mean_dist2": (_I, [_V, _V, _I, _F, _V, _V]),
    "fnx_pbf_next_tick_fwd": (_I, [_I, _V, _V, _V, _V, _F, _F, _F, _V, _V, _V]),
    "fnx_pbf_combine_grad": (_I, [_I, _V, _V, _F, _F, _F, _V, _V, _V, _F, _V, _V, _V]),
    "fnx_pbf_combine_grad_adam": (_I, [_I, _V, _V, _F, _F, _F, _V, _V, _V, _F, _V, _V, _V, _V, _F, _F, _F, _F, _V, _V, _V, _V]),
    "fnx_pbf_ratio_loss": (_I, [_I, _V, _F, _V, _V, _V]),
    "fnx_adam_step": (_I, [_I64, _V, _V, _V, _V, _F, _F, _F, _F, _F, _I, _V]),
    "fnx_adam_step_dev": (_I, [_I64, _V, _V, _V, _V, _F, _F, _F, _F, _F, _V, _V, _V]),
    "fnx_adam_step_dev_gated": (_I, [_I64, _V, _V, _V, _V, _F, _F, _F, _F, _F, _V, _V, _V, _V]),
    "fnx_scatter_min": (_I, [_I64, _V, _V, _I, _V, _V, _V]),
    "fnx_image_loss_bytes": (_SZ, [_I, _I, _I, _I]),
    "fnx_image_loss": (_I, [_I, _I, _I, _I, _V, _V, _I, _F, _F, _V, _V, _V, _V, _V]),
    "fnx_profile_sections": (_I, []),
    "fnx_profile_section_name": (C.c_char_p, [_I]),
    "fnx_profile_enable": (_I, [C.c_uint32]),
    "fnx_profile_collect": (_I, [_V, _V]),
    "fnx_raster_read_geom": (_I, [C.POINTER(RasterScratch), _I, _I, _V, _V, _V, _V, _V]),
    "fnx_raster_read_image": (_I, [C.POINTER(RasterScratch), _I, _I, _I, _V, _V, _V]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA library is the only implementation of this path "
                "(no CPU fallback).  Build it with `python __graft_entry__.py build`.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if l.fnx_abi_version() != 1:
            raise ImportError("libfnx ABI version mismatch")
        _lib = l
    return _lib


def check(code):
    if code != FNX_OK:
        raise FnxError(code, lib().fnx_last_error().decode())
