"""Host side of the surfel rasteriser: allocates outputs / state with torch and drives the C ABI.

This is the new equivalent of the reference's torch glue `RasterizeGaussiansCUDA` /
`RasterizeGaussiansBackwardCUDA` (svgss_rasterization/rasterize_points.cu:35-145, 147-265 and
rgss-rasterization/rasterize_points.cu:36-143, 145-243).  PyTorch is used for device memory and
the current stream only; all arithmetic happens in libsvgir_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import VARIANT_RGSS, VARIANT_SVGSS

# capacity hints for the binning buffers, keyed by (device, P, W, H): lets forward() enqueue every
# kernel before the single device->host read of num_rendered (the reference blocks mid-pipeline,
# rasterizer_impl.cu:311).
_CAP_HINT: dict = {}
SPECULATIVE = True


def _ptr(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _prep(t: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
    """Reference calls .contiguous() on every input; an empty tensor means 'absent'."""
    if t is None or t.numel() == 0:
        return None
    if t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


@dataclass
class RasterSettings:
    """Per-view constants (fields of both reference GaussianRasterizationSettings tuples)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool = False
    debug: bool = False
    variant: int = VARIANT_SVGSS
    patch_bbox: Optional[torch.Tensor] = None
    config: Optional[torch.Tensor] = None
    backward_geometry: bool = True
    computer_pseudo_normal: bool = False
    cx: float = 0.0
    cy: float = 0.0


class RasterState:
    """Typed replacement of geomBuffer / binningBuffer / imgBuffer (rasterizer_impl.h:32-75)."""

    def __init__(self):
        self.t = {}
        self.keep = []
        self.cfg = None
        self.cin = None
        self.cstate = None
        self.num_rendered = 0


def _make_cfg(s: RasterSettings, P, S, VS, M, dev, keep) -> _lib.RasterCfg:
    def dv(t, n):
        t = _prep(torch.as_tensor(t), dev)
        assert t is not None and t.numel() >= n, "camera constant has too few elements"
        keep.append(t)
        return t.data_ptr()

    cfg = _lib.RasterCfg()
    cfg.P, cfg.S, cfg.VS, cfg.sh_degree, cfg.M = P, S, VS, int(s.sh_degree), M
    cfg.W, cfg.H, cfg.variant = int(s.image_width), int(s.image_height), int(s.variant)
    cfg.tan_fovx, cfg.tan_fovy, cfg.scale_modifier = float(s.tanfovx), float(s.tanfovy), float(s.scale_modifier)
    cfg.prefiltered, cfg.debug = int(bool(s.prefiltered)), int(bool(s.debug))
    cfg.backward_geometry = int(bool(s.backward_geometry))
    cfg.computer_pseudo_normal = int(bool(s.computer_pseudo_normal))
    cfg.cx, cfg.cy = float(s.cx), float(s.cy)
    cfg.bg = dv(s.bg, 3)
    cfg.viewmatrix = dv(s.viewmatrix, 16)
    cfg.projmatrix = dv(s.projmatrix, 16)
    cfg.campos = dv(s.campos, 3)
    if s.variant == VARIANT_SVGSS:
        cfg.patch_bbox = dv(s.patch_bbox, 4)
        if s.config is not None and torch.as_tensor(s.config).numel() > 0:
            c = torch.as_tensor(s.config)
            cfg.n_config = int(c.numel())
            cfg.config = dv(c, 1)
        else:
            cfg.n_config = 0
            cfg.config = None
    else:
        cfg.n_config = 0
        cfg.patch_bbox = None
        cfg.config = None
    return cfg


def forward(s: RasterSettings, means3D, opacities, scales=None, rotations=None, cov3D_precomp=None,
            shs=None, colors_precomp=None, features=None, vfeatures=None, want_sorted_keys=False):
    """Returns (outputs dict, RasterState). Mirrors Rasterizer::forward (rasterizer_impl.cu:209-382)."""
    L = _lib.lib()
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:65-67
    if not means3D.is_cuda:
        raise RuntimeError("svgir_b200 rasteriser needs CUDA tensors (no CPU fallback)")
    dev = means3D.device
    P = means3D.shape[0]
    H, W = int(s.image_height), int(s.image_width)
    S = int(features.shape[1]) if features is not None and features.dim() == 2 else 0
    VS = int(vfeatures.shape[1]) if vfeatures is not None and vfeatures.dim() == 2 else 0
    shs_p = _prep(shs, dev)
    M = int(shs_p.shape[1]) if shs_p is not None else 0
    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    out = dict(
        color=torch.empty((3, H, W), **f32), normal=torch.empty((3, H, W), **f32),
        depth=torch.empty((1, H, W), **f32), opacity=torch.empty((1, H, W), **f32),
        feature=torch.empty((S, H, W), **f32), vfeature=torch.empty((VS // 4, H, W), **f32),
        weights=torch.zeros((P, 1), **f32), radii=torch.zeros((P,), **i32))
    if s.variant == VARIANT_RGSS:
        out["pseudo_normal"] = torch.zeros((3, H, W), **f32)
        out["surface_xyz"] = torch.zeros((3, H, W), **f32)
    st = RasterState()
    if P == 0:  # rasterize_points.cu:100: outputs stay zero
        for k in ("color", "normal", "depth", "opacity", "feature", "vfeature"):
            out[k].zero_()
        st.num_rendered = 0
        out["n_contrib"] = torch.zeros((H, W), **i32)
        return out, st

    keep = st.keep
    cin = _lib.RasterIn()
    tensors = dict(means3D=_prep(means3D, dev), opacities=_prep(opacities, dev), scales=_prep(scales, dev),
                   rotations=_prep(rotations, dev), cov3D_precomp=_prep(cov3D_precomp, dev), shs=shs_p,
                   colors_precomp=_prep(colors_precomp, dev), features=_prep(features, dev),
                   vfeatures=_prep(vfeatures, dev))
    for k, v in tensors.items():
        setattr(cin, k, _ptr(v))
    keep.extend(v for v in tensors.values() if v is not None)
    cfg = _make_cfg(s, P, S, VS, M, dev, keep)

    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    t = st.t
    t["rec"] = torch.empty((P, _lib.REC_FLOATS), **f32)
    t["cov3D"] = torch.empty((P, 6), **f32)
    t["clamped"] = torch.empty((P,), dtype=torch.uint8, device=dev)
    t["rect"] = torch.empty((P, 4), dtype=torch.int16, device=dev)
    t["tiles_touched"] = torch.empty((P,), **i32)
    t["tile_count"] = torch.empty((T,), **i32)
    t["tile_cursor"] = torch.empty((T,), **i32)
    t["ranges"] = torch.empty((T, 2), **i32)
    t["big_tiles"] = torch.empty((2 * T + 4,), **i32)
    t["num_rendered"] = torch.zeros((2,), **i32)
    t["final_T"] = torch.empty((H * W,), **f32)
    t["final_D"] = torch.empty((H * W,), **f32)
    t["n_contrib"] = torch.empty((H * W,), **i32)

    cst = _lib.RasterState()
    for k in ("rec", "cov3D", "clamped", "rect", "tiles_touched", "tile_count", "tile_cursor", "ranges",
              "big_tiles", "num_rendered", "final_T", "final_D", "n_contrib"):
        setattr(cst, k, t[k].data_ptr())
    cout = _lib.RasterOut()
    for k in ("color", "normal", "depth", "opacity", "feature", "vfeature", "weights", "radii"):
        setattr(cout, k, _ptr(out[k]))
    if s.variant == VARIANT_RGSS:
        cout.pseudo_normal = out["pseudo_normal"].data_ptr()
        cout.surface_xyz = out["surface_xyz"].data_ptr()

    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def alloc_bins(cap):
        cap = max(int(cap), 1)
        t["keys"] = torch.empty((cap,), dtype=torch.int64, device=dev)
        t["point_list"] = torch.empty((cap,), **i32)
        t["sorted_keys"] = torch.empty((cap,), dtype=torch.int64, device=dev) if want_sorted_keys else None
        cst.keys = t["keys"].data_ptr()
        cst.point_list = t["point_list"].data_ptr()
        cst.sorted_keys = _ptr(t["sorted_keys"])
        cst.cap_R = cap

    with torch.cuda.device(dev):
        _lib.check(L.svgir_raster_preprocess(C.byref(cfg), C.byref(cin), C.byref(cst), C.byref(cout), stream),
                   "raster_preprocess")
        hint_key = (dev.index, P, W, H)
        hint = _CAP_HINT.get(hint_key) if SPECULATIVE else None
        if hint is None:
            R = int(t["num_rendered"][0].item())  # one 4-byte D2H, like rasterizer_impl.cu:311
            alloc_bins(R)
            _lib.check(L.svgir_raster_render(C.byref(cfg), C.byref(cin), C.byref(cst), C.byref(cout), stream),
                       "raster_render")
        else:
            alloc_bins(hint)
            _lib.check(L.svgir_raster_render(C.byref(cfg), C.byref(cin), C.byref(cst), C.byref(cout), stream),
                       "raster_render")
            R, overflow = t["num_rendered"].tolist()  # the only sync, after everything is enqueued
            if overflow:
                alloc_bins(R)
                _lib.check(L.svgir_raster_render(C.byref(cfg), C.byref(cin), C.byref(cst), C.byref(cout), stream),
                           "raster_render")
        _CAP_HINT[hint_key] = int(R * 1.25) + 4096
    st.cfg, st.cin, st.cstate, st.num_rendered = cfg, cin, cst, R
    out["n_contrib"] = t["n_contrib"].view(H, W)
    if R == 0:
        # nothing binned: the compositor still wrote background-only images
        pass
    return out, st


def backward(st: RasterState, radii, grads: dict, want_debug=False):
    """grads: dL_dcolor, dL_dnormal, dL_ddepth, dL_dopacity, dL_dfeature, dL_dvfeature (pixel space).
    Returns dict named like the reference's 13-tuple (rasterize_points.cu:264)."""
    L = _lib.lib()
    cfg = st.cfg
    P, S, VS, M = cfg.P, cfg.S, cfg.VS, cfg.M
    dev = radii.device
    f32 = dict(dtype=torch.float32, device=dev)
    res = dict(
        dL_dmeans2D=torch.empty((P, 3), **f32), dL_dcolors=torch.empty((P, 3), **f32),
        dL_dopacity=torch.empty((P, 1), **f32), dL_dmeans3D=torch.empty((P, 3), **f32),
        dL_dfeatures=torch.zeros((P, S), **f32), dL_dvfeatures=torch.zeros((P, VS), **f32),
        dL_dcov3D=torch.empty((P, 6), **f32), dL_dsh=torch.empty((P, M, 3), **f32),
        dL_dscales=torch.empty((P, 3), **f32), dL_drotations=torch.empty((P, 4), **f32),
        dL_dviewmat=torch.zeros((4, 4), **f32), dL_dprojmat=torch.zeros((4, 4), **f32),
        dL_dcampos=torch.zeros((3,), **f32))
    if P == 0:
        return res
    geo = torch.zeros((P, _lib.GEO_GRAD_FLOATS), **f32)
    g = _lib.RasterGrads()
    keep = []
    for name, key in (("dL_dcolor", "dL_dcolor"), ("dL_dnormal", "dL_dnormal"), ("dL_ddepth", "dL_ddepth"),
                      ("dL_dopacity", "dL_dopacity"), ("dL_dfeature", "dL_dfeature"),
                      ("dL_dvfeature", "dL_dvfeature")):
        tns = _prep(grads.get(key), dev)
        keep.append(tns)
        setattr(g, name, _ptr(tns))
    H, W = cfg.H, cfg.W
    for name, shape in (("dL_dcolor", 3), ("dL_dnormal", 3), ("dL_ddepth", 1), ("dL_dopacity", 1)):
        if getattr(g, name) is None:  # autograd hands None for unused outputs
            z = torch.zeros((shape, H, W), **f32)
            keep.append(z)
            setattr(g, name, z.data_ptr())
    if S > 0 and g.dL_dfeature is None:
        z = torch.zeros((S, H, W), **f32); keep.append(z); g.dL_dfeature = z.data_ptr()
    if VS > 0 and g.dL_dvfeature is None:
        z = torch.zeros((VS // 4, H, W), **f32); keep.append(z); g.dL_dvfeature = z.data_ptr()
    g.geo_grad = geo.data_ptr()
    g.dL_dfeatures = _ptr(res["dL_dfeatures"])
    g.dL_dvfeatures = _ptr(res["dL_dvfeatures"])
    g.dL_dmeans2D = res["dL_dmeans2D"].data_ptr()
    g.dL_dcolors = res["dL_dcolors"].data_ptr()
    g.dL_dopacities = res["dL_dopacity"].data_ptr()
    g.dL_dmeans3D = res["dL_dmeans3D"].data_ptr()
    g.dL_dcov3D = res["dL_dcov3D"].data_ptr()
    g.dL_dsh = _ptr(res["dL_dsh"])
    g.dL_dscales = res["dL_dscales"].data_ptr()
    g.dL_drotations = res["dL_drotations"].data_ptr()
    if want_debug:
        res["dL_dconic"] = torch.empty((P, 4), **f32)
        res["dL_dnormal"] = torch.empty((P, 3), **f32)
        res["dL_ddepth"] = torch.empty((P, 1), **f32)
        g.dL_dconic = res["dL_dconic"].data_ptr()
        g.dL_dnormal3 = res["dL_dnormal"].data_ptr()
        g.dL_ddepths = res["dL_ddepth"].data_ptr()
    radii = radii.contiguous()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _lib.check(L.svgir_raster_backward(C.byref(cfg), C.byref(st.cin), C.byref(st.cstate),
                                           C.c_void_p(radii.data_ptr()), C.byref(g), stream), "raster_backward")
    res["_geo_grad"] = geo
    return res


def mark_visible(variant: int, positions, viewmatrix, projmatrix):
    """rasterize_points.cu:267-286."""
    L = _lib.lib()
    P = positions.shape[0]
    present = torch.zeros((P,), dtype=torch.bool, device=positions.device)
    if P:
        dev = positions.device
        pos, vm, pm = _prep(positions, dev), _prep(viewmatrix, dev), _prep(projmatrix, dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(L.svgir_mark_visible(variant, P, pos.data_ptr(), vm.data_ptr(), pm.data_ptr(),
                                            present.data_ptr(), stream), "mark_visible")
    return present
