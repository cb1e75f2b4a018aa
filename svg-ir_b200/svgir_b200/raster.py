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
# How forward() learns num_rendered (R), the size of the binning buffers:
#   "exact"        read R after the preprocess, then bin (the reference's own order, one mid-pipeline sync)
#   "speculative"  bin into a buffer sized from the previous call, read (R, overflow) once after every
#                  kernel is enqueued, re-render on overflow -- always correct, one sync per forward (default)
#   "async"        like speculative, but (R, overflow) is copied to pinned host memory asynchronously and
#                  NOT awaited: forward() never blocks, so the whole step can be enqueued ahead of the GPU or
#                  captured into a CUDA graph. The owner of the step must call RasterState.resolve() before
#                  trusting the results (pipeline.training_step / GraphedTrainingStep do and redo the step
#                  on overflow); LazyCount / backward() raise CapacityOverflow otherwise.
COUNT_MODE = "speculative"
OWNER_RESOLVES = False  # async: the caller of forward() promises to resolve() the state itself (no check in backward)
ASYNC_MARGIN = 65536
ASYNC_SLACK = 2.0      # async capacity = ASYNC_SLACK x the largest R seen for this (device, P, W, H)


class CapacityOverflow(RuntimeError):
    """Raised when an async forward binned more instances than its buffers hold."""

    def __init__(self, needed: int, cap: int):
        super().__init__("svgir_b200: binning capacity overflow (num_rendered=%d > capacity=%d); the images of "
                         "this forward are incomplete -- re-run it (the capacity hint has been raised)" % (needed, cap))
        self.needed, self.cap = needed, cap


class count_mode:
    """Context manager: `with raster.count_mode("async"): ...`"""

    def __init__(self, mode: str, owner_resolves: bool = False):
        assert mode in ("exact", "speculative", "async")
        self.mode, self.owner = mode, owner_resolves

    def __enter__(self):
        global COUNT_MODE, OWNER_RESOLVES
        self.prev, COUNT_MODE = (COUNT_MODE, OWNER_RESOLVES), self.mode
        OWNER_RESOLVES = self.owner
        return self

    def __exit__(self, *a):
        global COUNT_MODE, OWNER_RESOLVES
        COUNT_MODE, OWNER_RESOLVES = self.prev
        return False


class LazyCount:
    """num_rendered of an async forward: resolves (waits for the 8-byte D2H copy) on first use."""

    def __init__(self, st: "RasterState"):
        self._st = st

    def __int__(self):
        return self._st.resolve()

    __index__ = __int__

    def __repr__(self):
        return str(int(self))

    def __eq__(self, o):
        return int(self) == o

    def __hash__(self):
        return hash(int(self))

    def __gt__(self, o):
        return int(self) > o

    def __lt__(self, o):
        return int(self) < o

    def __add__(self, o):
        return int(self) + o

    __radd__ = __add__


_PINNED_RING = None      # [256,2] int32 pinned: (R, overflow) landing slots of eager async forwards
_PINNED_NEXT = 0
_CAPTURE_SLOTS: list = []  # dedicated pinned slots for forwards captured into CUDA graphs (allocated before capture)


_PINNED_FOREVER: list = []


def pinned_forever(shape, dtype) -> torch.Tensor:
    """Page-locked landing buffer for copies that get captured into CUDA graphs. It is never returned to torch's
    caching host allocator: the allocator records an event on every stream that used a block and queries it before
    re-using the block, and an event recorded on a CAPTURING stream cannot be queried (cudaErrorInvalidValue at some
    later, unrelated pin_memory() call). A few bytes per graph are kept for the life of the process instead."""
    t = torch.zeros(shape, dtype=dtype).pin_memory()
    _PINNED_FOREVER.append(t)
    return t


def prepare_capture(n_forwards: int = 1):
    """Call BEFORE capturing `n_forwards` async forwards into a CUDA graph: page-locked memory cannot be
    allocated while a stream is capturing, and a graph's landing slot must never be recycled."""
    for _ in range(n_forwards):
        _CAPTURE_SLOTS.append(pinned_forever((2,), torch.int32))


def _pinned_slot(capturing: bool) -> torch.Tensor:
    global _PINNED_RING, _PINNED_NEXT
    if capturing:
        if not _CAPTURE_SLOTS:
            raise RuntimeError("svgir_b200: call raster.prepare_capture() before capturing an async forward")
        return _CAPTURE_SLOTS.pop()
    if _PINNED_RING is None:
        _PINNED_RING = torch.zeros((256, 2), dtype=torch.int32).pin_memory()
    _PINNED_NEXT = (_PINNED_NEXT + 1) % 256
    return _PINNED_RING[_PINNED_NEXT]


def reserve(device, P: int, W: int, H: int, capacity: int):
    """Pre-size the binning buffers for (device, P, W, H): lets the very first forward run without a sync."""
    idx = torch.device(device).index
    _CAP_HINT[(idx if idx is not None else torch.cuda.current_device(), P, W, H)] = int(capacity)


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
        self.count_host = None    # pinned [2] int32: (R, overflow) of an async forward
        self.count_event = None
        self.hint_key = None
        self.pending = False
        self.owner_resolves = False
        self.cout = None
        self.out = {}
        self.settings = None
        self.dims = None
        self.src = None
        self.rendered = False
        self.binned = None        # event: start_binning() launched part 2a on the side stream

    def resolve(self) -> int:
        """Wait for the (R, overflow) copy of an async forward; raises CapacityOverflow if the bins
        were too small (after raising the capacity hint so that a re-run fits)."""
        if self.pending:
            if self.count_event is not None:
                self.count_event.synchronize()
            else:  # captured into a CUDA graph: the owner synchronises after the replay
                torch.cuda.current_stream().synchronize()
            R, overflow = self.count_host.tolist()
            self.pending = False
            self.num_rendered = int(R)
            cap = int(self.cstate.cap_R)
            _CAP_HINT[self.hint_key] = max(_CAP_HINT.get(self.hint_key, 0), int(R * ASYNC_SLACK) + ASYNC_MARGIN)
            if overflow:
                raise CapacityOverflow(int(R), cap)
        return int(self.num_rendered)


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


def preprocess(s: RasterSettings, means3D, opacities, scales=None, rotations=None, cov3D_precomp=None,
               shs=None, colors_precomp=None, want_vis_list=False) -> RasterState:
    """Forward, part 1 (svgir_raster_preprocess): per-surfel projection / culling / tile counting, which
    needs the geometry only. Returns the state that `forward(..., prestate=...)` completes. With
    want_vis_list the state also carries st.t["vis_list"] / ["vis_count"]: the surfels that survive culling,
    so that per-surfel work feeding `features` / `vfeatures` (the render_equation shading) can skip the rest."""
    L = _lib.lib()
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:65-67
    if not means3D.is_cuda:
        raise RuntimeError("svgir_b200 rasteriser needs CUDA tensors (no CPU fallback)")
    dev = means3D.device
    P = means3D.shape[0]
    H, W = int(s.image_height), int(s.image_width)
    shs_p = _prep(shs, dev)
    M = int(shs_p.shape[1]) if shs_p is not None else 0
    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    st = RasterState()
    st.settings, st.dims, st.src = s, (P, H, W, M), means3D
    st.out = dict(weights=torch.zeros((P, 1), **f32), radii=torch.zeros((P,), **i32))
    t = st.t
    if want_vis_list:
        t["vis_list"] = torch.empty((max(P, 1),), **i32)
        t["vis_count"] = torch.zeros((1,), **i32)
    if P == 0:
        return st
    keep = st.keep
    cin = _lib.RasterIn()
    tensors = dict(means3D=_prep(means3D, dev), opacities=_prep(opacities, dev), scales=_prep(scales, dev),
                   rotations=_prep(rotations, dev), cov3D_precomp=_prep(cov3D_precomp, dev), shs=shs_p,
                   colors_precomp=_prep(colors_precomp, dev))
    for k, v in tensors.items():
        setattr(cin, k, _ptr(v))
    keep.extend(v for v in tensors.values() if v is not None)
    cfg = _make_cfg(s, P, 0, 0, M, dev, keep)

    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    t["rec"] = torch.empty((P, _lib.REC_FLOATS), **f32)
    t["cov3D"] = torch.empty((P, 6), **f32)
    t["clamped"] = torch.empty((P,), dtype=torch.uint8, device=dev)
    t["rect"] = torch.empty((P, 4), dtype=torch.int16, device=dev)
    t["tiles_touched"] = torch.empty((P,), **i32)
    t["tile_count"] = torch.empty((T,), **i32)
    t["tile_cursor"] = torch.empty((T,), **i32)
    t["ranges"] = torch.empty((T, 2), **i32)
    t["big_tiles"] = torch.empty((3 * T + 4,), **i32)   # sorter work lists + the tiles' compositing order
    t["num_rendered"] = torch.zeros((2,), **i32)
    t["final_T"] = torch.empty((H * W,), **f32)
    t["final_D"] = torch.empty((H * W,), **f32)
    t["n_contrib"] = torch.empty((H * W,), **i32)

    cst = _lib.RasterState()
    for k in ("rec", "cov3D", "clamped", "rect", "tiles_touched", "tile_count", "tile_cursor", "ranges",
              "big_tiles", "num_rendered", "final_T", "final_D", "n_contrib"):
        setattr(cst, k, t[k].data_ptr())
    if want_vis_list:
        cst.vis_list, cst.vis_count = t["vis_list"].data_ptr(), t["vis_count"].data_ptr()
    cout = _lib.RasterOut()
    cout.weights, cout.radii = st.out["weights"].data_ptr(), st.out["radii"].data_ptr()
    st.cfg, st.cin, st.cstate, st.cout = cfg, cin, cst, cout
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _lib.check(L.svgir_raster_preprocess(C.byref(cfg), C.byref(cin), C.byref(cst), C.byref(cout), stream),
                   "raster_preprocess")
    return st


_BIN_STREAMS: dict = {}


def start_binning(st: RasterState) -> bool:
    """Forward, part 2a, early: launches the tile binning (scan, duplicateWithKeys, sort -- geometry only) of a
    `preprocess()`-ed state on a side stream, so that the caller's stream can run what produces `features` /
    `vfeatures` (the render_equation shading) meanwhile; `forward(..., prestate=st)` then waits for it and only
    composites. Fork and join are events, hence capturable. Needs a known binning capacity without a host read
    (count mode "async" with a capacity hint); otherwise it does nothing and forward() bins in order. Returns
    whether the binning was launched."""
    if st.cfg is None or st.rendered or st.binned is not None:
        return False
    P, H, W, M = st.dims
    dev = st.src.device
    mode = COUNT_MODE if SPECULATIVE else "exact"
    hint = _CAP_HINT.get((dev.index, P, W, H)) if mode == "async" else None
    if hint is None:
        return False
    t, cst = st.t, st.cstate
    cap = max(int(hint), 1)
    t["keys"] = torch.empty((cap,), dtype=torch.int64, device=dev)
    t["point_list"] = torch.empty((cap,), dtype=torch.int32, device=dev)
    t["sorted_keys"] = None
    cst.keys, cst.point_list, cst.sorted_keys, cst.cap_R = t["keys"].data_ptr(), t["point_list"].data_ptr(), None, cap
    side = _BIN_STREAMS.get(dev.index)
    if side is None:
        side = _BIN_STREAMS[dev.index] = torch.cuda.Stream(dev)
    cur = torch.cuda.current_stream(dev)
    fork = torch.cuda.Event()
    fork.record(cur)
    side.wait_event(fork)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().svgir_raster_bin(C.byref(st.cfg), C.byref(st.cin), C.byref(cst), C.byref(st.cout),
                                               C.c_void_p(side.cuda_stream)), "raster_bin")
    st.binned = torch.cuda.Event()
    st.binned.record(side)
    return True


def forward(s: RasterSettings, means3D, opacities, scales=None, rotations=None, cov3D_precomp=None,
            shs=None, colors_precomp=None, features=None, vfeatures=None, want_sorted_keys=False,
            prestate: Optional[RasterState] = None):
    """Returns (outputs dict, RasterState). Mirrors Rasterizer::forward (rasterizer_impl.cu:209-382).
    `prestate` = the result of `preprocess()` on the same geometry and settings: only part 2 (binning,
    sort, compositing) runs."""
    L = _lib.lib()
    if prestate is not None:
        st = prestate
        if st.src is not means3D or st.rendered:
            raise RuntimeError("svgir_b200: prestate belongs to another forward call")
    else:
        st = preprocess(s, means3D, opacities, scales, rotations, cov3D_precomp, shs, colors_precomp)
    st.rendered = True
    dev = means3D.device
    P, H, W, M = st.dims
    S = int(features.shape[1]) if features is not None and features.dim() == 2 else 0
    VS = int(vfeatures.shape[1]) if vfeatures is not None and vfeatures.dim() == 2 else 0
    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    out = st.out
    out.update(
        color=torch.empty((3, H, W), **f32), normal=torch.empty((3, H, W), **f32),
        depth=torch.empty((1, H, W), **f32), opacity=torch.empty((1, H, W), **f32),
        feature=torch.empty((S, H, W), **f32), vfeature=torch.empty((VS // 4, H, W), **f32))
    if s.variant == VARIANT_RGSS:
        out["pseudo_normal"] = torch.zeros((3, H, W), **f32)
        out["surface_xyz"] = torch.zeros((3, H, W), **f32)
    if P == 0:  # rasterize_points.cu:100: outputs stay zero
        for k in ("color", "normal", "depth", "opacity", "feature", "vfeature"):
            out[k].zero_()
        st.num_rendered = 0
        out["n_contrib"] = torch.zeros((H, W), **i32)
        return out, st

    t = st.t
    cfg, cin, cst, cout = st.cfg, st.cin, st.cstate, st.cout
    fe, vf = _prep(features, dev), _prep(vfeatures, dev)
    st.keep.extend(x for x in (fe, vf) if x is not None)
    cin.features, cin.vfeatures = _ptr(fe), _ptr(vf)
    cfg.S, cfg.VS = S, VS
    for k in ("color", "normal", "depth", "opacity", "feature", "vfeature"):
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
        hint_key = (dev.index, P, W, H)
        mode = COUNT_MODE if SPECULATIVE else "exact"
        hint = _CAP_HINT.get(hint_key) if mode != "exact" else None
        capturing = torch.cuda.is_current_stream_capturing()
        if hint is None and capturing:
            raise RuntimeError("svgir_b200: a forward captured into a CUDA graph needs a binning capacity "
                               "(run one eager forward first, or raster.reserve())")
        if hint is not None and mode == "async":
            if st.binned is not None and not want_sorted_keys:   # start_binning() ran it on the side stream
                torch.cuda.current_stream(dev).wait_event(st.binned)
                _lib.check(L.svgir_raster_composite(C.byref(cfg), C.byref(cin), C.byref(cst), C.byref(cout), stream),
                           "raster_composite")
            else:
                if st.binned is not None:
                    torch.cuda.current_stream(dev).wait_event(st.binned)
                alloc_bins(hint)
                _lib.check(L.svgir_raster_render(C.byref(cfg), C.byref(cin), C.byref(cst), C.byref(cout), stream),
                           "raster_render")
            st.count_host = _pinned_slot(capturing)
            st.count_host.copy_(t["num_rendered"], non_blocking=True)
            if not capturing:
                st.count_event = torch.cuda.Event()
                st.count_event.record()
            st.hint_key, st.pending, st.owner_resolves = hint_key, True, OWNER_RESOLVES
            st.num_rendered = LazyCount(st)
            out["n_contrib"] = t["n_contrib"].view(H, W)
            return out, st
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
        _CAP_HINT[hint_key] = int(R * (ASYNC_SLACK if mode == "async" else 1.25)) + 4096
    st.num_rendered = R
    out["n_contrib"] = t["n_contrib"].view(H, W)
    if R == 0:
        # nothing binned: the compositor still wrote background-only images
        pass
    return out, st


def backward(st: RasterState, radii, grads: dict, want_debug=False):
    """grads: dL_dcolor, dL_dnormal, dL_ddepth, dL_dopacity, dL_dfeature, dL_dvfeature (pixel space).
    Returns dict named like the reference's 13-tuple (rasterize_points.cu:264)."""
    L = _lib.lib()
    if st.pending and not st.owner_resolves:
        st.resolve()  # async forward whose count nobody checked yet: raises CapacityOverflow if it did not fit
    cfg = st.cfg
    if cfg is None:  # P == 0: nothing was launched
        P, S, VS, M = 0, int(st.out["feature"].shape[0]), 4 * int(st.out["vfeature"].shape[0]), st.dims[3]
    else:
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
