"""ctypes binding of libsvgir_b200.so (the C ABI declared in include/svgir_b200.h).

The library is the only compute path: if it is missing or cannot be loaded this module raises --
there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# SVGIR_B200_LIB: load another build of the same library (kernel A/B tuning runs); default = the in-tree build
_SO = os.environ.get("SVGIR_B200_LIB") or os.path.join(_HERE, "libsvgir_b200.so")
_CSRC = os.path.join(os.path.dirname(_HERE), "csrc")

REC_FLOATS = 24
GEO_GRAD_FLOATS = 16
MAX_S = 64
MAX_NV = 32
VARIANT_SVGSS = 0
VARIANT_RGSS = 1

c_fp = C.c_void_p  # device pointers travel as integers


class RasterCfg(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("S", C.c_int32), ("VS", C.c_int32), ("sh_degree", C.c_int32), ("M", C.c_int32),
        ("W", C.c_int32), ("H", C.c_int32), ("variant", C.c_int32),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float),
        ("prefiltered", C.c_int32), ("debug", C.c_int32), ("n_config", C.c_int32),
        ("backward_geometry", C.c_int32), ("computer_pseudo_normal", C.c_int32),
        ("cx", C.c_float), ("cy", C.c_float),
        ("bg", c_fp), ("viewmatrix", c_fp), ("projmatrix", c_fp), ("campos", c_fp),
        ("patch_bbox", c_fp), ("config", c_fp),
    ]


class RasterIn(C.Structure):
    _fields_ = [(n, c_fp) for n in ("means3D", "opacities", "scales", "rotations", "cov3D_precomp",
                                    "shs", "colors_precomp", "features", "vfeatures")]


class RasterState(C.Structure):
    _fields_ = [
        ("rec", c_fp), ("cov3D", c_fp), ("clamped", c_fp), ("rect", c_fp), ("tiles_touched", c_fp),
        ("tile_count", c_fp), ("tile_cursor", c_fp), ("ranges", c_fp), ("big_tiles", c_fp),
        ("num_rendered", c_fp), ("keys", c_fp), ("point_list", c_fp), ("sorted_keys", c_fp),
        ("cap_R", C.c_int64), ("final_T", c_fp), ("final_D", c_fp), ("n_contrib", c_fp),
        ("vis_list", c_fp), ("vis_count", c_fp),
    ]


class RasterOut(C.Structure):
    _fields_ = [(n, c_fp) for n in ("color", "normal", "depth", "opacity", "feature", "vfeature",
                                    "weights", "radii", "pseudo_normal", "surface_xyz")]


class RasterGrads(C.Structure):
    _fields_ = [(n, c_fp) for n in (
        "dL_dcolor", "dL_dnormal", "dL_ddepth", "dL_dopacity", "dL_dfeature", "dL_dvfeature",
        "geo_grad", "dL_dfeatures", "dL_dvfeatures",
        "dL_dmeans2D", "dL_dcolors", "dL_dopacities", "dL_dmeans3D", "dL_dcov3D", "dL_dsh",
        "dL_dscales", "dL_drotations", "dL_dconic", "dL_dnormal3", "dL_ddepths")]


class ParamGrads(C.Structure):
    """svgir_param_grads: parameter-gradient buffers svgir_raster_backward_params adds into."""
    _fields_ = [(n, c_fp) for n in ("d_means3D", "d_opacities", "d_scales", "d_rotations", "d_sh", "d_means2D")]


_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", _CSRC, "-j8"], stdout=None if verbose else subprocess.DEVNULL)
    return _SO


MAX_PEERS = 8
PEER_BLOCKS = 128
PEER_BANKS = 4
PEER_FLAG_WORDS = PEER_BANKS * 2 * PEER_BLOCKS * MAX_PEERS


class PeerComm(C.Structure):
    """svgir_peer_comm (include/svgir_b200.h)."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("bufs", c_fp * MAX_PEERS), ("flags", c_fp * MAX_PEERS),
                ("multicast", c_fp)]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise RuntimeError(
            f"svgir_b200: native library {_SO} is missing. Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` or `make -C {_CSRC}`. "
            "There is no CPU / PyTorch fallback for this path.")
    L = C.CDLL(_SO)
    L.svgir_last_error.restype = C.c_char_p
    for name in ("svgir_raster_preprocess", "svgir_raster_render", "svgir_raster_bin", "svgir_raster_composite"):
        getattr(L, name).argtypes = [C.POINTER(RasterCfg), C.POINTER(RasterIn), C.POINTER(RasterState),
                                     C.POINTER(RasterOut), C.c_void_p]
        getattr(L, name).restype = C.c_int
    L.svgir_raster_backward.argtypes = [C.POINTER(RasterCfg), C.POINTER(RasterIn), C.POINTER(RasterState),
                                        C.c_void_p, C.POINTER(RasterGrads), C.c_void_p]
    L.svgir_raster_backward.restype = C.c_int
    L.svgir_raster_backward_composite.argtypes = [C.POINTER(RasterCfg), C.POINTER(RasterIn), C.POINTER(RasterState),
                                                  C.POINTER(RasterGrads), C.c_void_p]
    L.svgir_raster_backward_composite.restype = C.c_int
    L.svgir_raster_backward_params.argtypes = [C.POINTER(RasterCfg), C.POINTER(RasterIn), C.POINTER(RasterState),
                                               C.c_void_p, C.POINTER(ParamGrads), C.c_void_p]
    L.svgir_raster_backward_params.restype = C.c_int
    L.svgir_mark_visible.argtypes = [C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, C.c_void_p]
    L.svgir_mark_visible.restype = C.c_int
    L.svgir_timing_enable.argtypes = [C.c_int]
    L.svgir_timing_enable.restype = None
    L.svgir_timing_collect.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    L.svgir_timing_collect.restype = C.c_int
    L.svgir_launch_count.argtypes = [C.c_int]
    L.svgir_launch_count.restype = C.c_longlong
    L.svgir_peer_allreduce.argtypes = [C.POINTER(PeerComm), C.c_longlong, C.c_void_p]
    L.svgir_peer_allreduce.restype = C.c_int
    L.svgir_peer_allreduce_range.argtypes = [C.POINTER(PeerComm), C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_void_p]
    L.svgir_peer_allreduce_range.restype = C.c_int
    L.svgir_shade_reserve_sms.argtypes = [C.c_int]
    L.svgir_shade_reserve_sms.restype = None
    _lib = L
    return L


def timing_enable(on: bool) -> None:
    lib().svgir_timing_enable(1 if on else 0)


def timing_collect(name=None, reset=False):
    """(total_ms, launches) of the recorded launches of kernel `name` (None = all kernels)."""
    ms, n = C.c_double(0), C.c_int(0)
    lib().svgir_timing_collect(name.encode() if name else None, C.byref(ms), C.byref(n), 1 if reset else 0)
    return ms.value, n.value


def launch_count(reset=False) -> int:
    return int(lib().svgir_launch_count(1 if reset else 0))


def check(rc: int, what: str) -> None:
    """Nonzero status -> RuntimeError, like the reference's AT_ERROR / std::runtime_error path
    (svgss_rasterization/rasterize_points.cu:65-67, cuda_rasterizer/auxiliary.h:425-432)."""
    if rc != 0:
        msg = lib().svgir_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"svgir_b200.{what} failed (status {rc}): {msg}")


EXPORTED_SYMBOLS = [
    "svgir_last_error", "svgir_version", "svgir_raster_preprocess", "svgir_raster_render",
    "svgir_raster_backward", "svgir_mark_visible", "svgir_shade_forward", "svgir_shade_backward",
    "svgir_direct_light_forward", "svgir_direct_light_backward", "svgir_timing_enable",
    "svgir_timing_collect", "svgir_launch_count", "svgir_bvh_workspace_bytes", "svgir_bvh_leaf_aabbs",
    "svgir_bvh_build", "svgir_bvh_pack_leaves", "svgir_bvh_trace_opacity",
    "svgir_sample_incident_rays", "svgir_render_equation_sh_forward", "svgir_render_equation_sh_backward",
    "svgir_train_loss_blocks", "svgir_train_loss_forward", "svgir_train_loss_backward", "svgir_peer_allreduce", "svgir_resolve_eval",
    "svgir_peer_allreduce_range", "svgir_shade_reserve_sms",
    "svgir_ssim_blocks", "svgir_ssim_forward", "svgir_ssim_backward",
    "svgir_raster_bin", "svgir_raster_composite", "svgir_raster_backward_composite", "svgir_raster_backward_params",
    "svgir_edge_aware_blocks", "svgir_edge_aware_forward", "svgir_edge_aware_backward", "svgir_tv_loss",
    "svgir_adam_step", "svgir_densify_stats", "svgir_densify_decide", "svgir_densify_index", "svgir_gather_rows",
    "svgir_densify_split",
    "svgir_radiance_pack_surfels", "svgir_radiance_cache_build", "svgir_radiance_loss_forward", "svgir_radiance_loss_backward",
    "svgir_radiance_loss_forward_backward",
]
