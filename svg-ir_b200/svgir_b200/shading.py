"""Host side of the fused PBR render_equation (shading) kernels.

Public entry points
  shade_surfels(...)        fused fast path: every [N,*] output of the reference's
                            rendering_equation4 plus the sample means render_view packs
                            (gaussian_renderer/svgss.py:116-166), differentiable.
  rendering_equation4(...)  same 9 arguments and return value `(pbr, extra_results)` as
                            gaussian_renderer/svgss.py:537-593 (drop-in).
  direct_light(...)         DirectLightMap.direct_light / EnvLight.direct_light
                            (scene/direct_light_map.py:70-83, scene/envmap.py:54-72), differentiable
                            w.r.t. the env parameter.
All arithmetic runs in libsvgir_b200.so (csrc/shading.cu); there is no torch fallback.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib

c_fp = C.c_void_p


class ShadeCfg(C.Structure):
    _fields_ = [("N", C.c_int32), ("Ns", C.c_int32), ("env_h", C.c_int32), ("env_w", C.c_int32),
                ("env_mode", C.c_int32), ("debug", C.c_int32), ("flags", C.c_int32), ("reserved_", C.c_int32)]


SHADE_ENV_READY = 1    # SVGIR_SHADE_ENV_READY: env_act_scratch already holds the activated map
SHADE_ACCUMULATE = 2   # SVGIR_SHADE_ACCUMULATE: backward adds into the d_* parameter-gradient buffers
SHADE_VIEW_4X4 = 4     # SVGIR_SHADE_VIEW_4X4: view3x3 points at the 4x4 world-view matrix itself


class ShadeIn(C.Structure):
    _fields_ = [(n, c_fp) for n in ("base_color", "roughness", "metallic", "normals", "viewdirs", "radiance",
                                    "visibility", "incident_dirs", "incident_areas", "env", "env_transform",
                                    "env_act_scratch", "view3x3", "surfel_list", "surfel_count",
                                    "means3D", "campos", "skip_flag", "env_taps")]


class ShadeOut(C.Structure):
    _fields_ = [(n, c_fp) for n in ("pbr", "diffuse_light", "specular", "direct", "indirect",
                                    "mean_visibility", "mean_local", "mean_incident", "mean_global", "pack",
                                    "sum_direct", "sum_indirect")] + \
               [(n, C.c_int32) for n in ("row_stride", "mean_vis_stride", "mean_stride", "reserved_")]


class ShadeGrads(C.Structure):
    _fields_ = [(n, c_fp) for n in ("g_pbr", "g_diffuse_light", "g_specular", "g_direct", "g_indirect",
                                    "g_mean_visibility", "g_mean_local", "g_mean_incident", "g_mean_global",
                                    "d_base_color", "d_roughness", "d_metallic", "d_normals", "d_viewdirs",
                                    "d_radiance", "d_visibility", "d_env", "g_pack", "sum_direct", "sum_indirect",
                                    "d_env_scratch")] + \
               [(n, C.c_int32) for n in ("g_row_stride", "g_mean_vis_stride", "g_mean_stride", "reserved_")] + \
               [("d_means3D", c_fp)]


_bound = False


def _L():
    global _bound
    L = _lib.lib()
    if not _bound:
        L.svgir_shade_forward.argtypes = [C.POINTER(ShadeCfg), C.POINTER(ShadeIn), C.POINTER(ShadeOut), C.c_void_p]
        L.svgir_shade_forward.restype = C.c_int
        L.svgir_shade_backward.argtypes = [C.POINTER(ShadeCfg), C.POINTER(ShadeIn), C.POINTER(ShadeGrads), C.c_void_p]
        L.svgir_shade_backward.restype = C.c_int
        L.svgir_direct_light_forward.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, C.c_void_p]
        L.svgir_direct_light_forward.restype = C.c_int
        L.svgir_direct_light_backward.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, C.c_void_p]
        L.svgir_direct_light_backward.restype = C.c_int
        L.svgir_env_taps.argtypes = [C.c_longlong, C.c_int, C.c_int, c_fp, c_fp, c_fp, C.c_void_p]
        L.svgir_env_taps.restype = C.c_int
        _bound = True
    return L


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _p(t):
    return None if t is None else t.data_ptr()


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


ENV_COPIES = 16       # SVGIR_SHADE_ENV_COPIES

# ---- env tap cache -------------------------------------------------------------------------------------------------
# The incident directions of a model are fixed between two update_radiace / update_visibility calls (training) and
# across all views and env maps of a relight sweep, so the texel corner + bilinear weights of every (surfel, sample)
# -- the acos / atan2 part of DirectLightMap.direct_light, 7-14 % of the shading kernels' instructions -- are computed
# once (svgir_env_taps) and handed to the kernels through svgir_shade_in.env_taps. Keyed by the directions' storage,
# shape and version, the env size and the transform; a changed version recomputes INTO THE SAME buffer, so CUDA graphs
# that baked the pointer stay valid (their owners call refresh_env_taps before a replay).
ENV_TAP_CACHE = True
_TAP_CACHE: "OrderedDict[tuple, list]" = OrderedDict()
_TAP_CACHE_SIZE = 4


def env_taps(dirs: torch.Tensor, env_h: int, env_w: int, transform: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[...,3] directions -> [...,3] taps (svgir_env_taps)."""
    L = _L()
    d = _c(dirs)
    n = d.numel() // 3
    taps = out if out is not None else torch.empty(tuple(d.shape), dtype=torch.float32, device=d.device)
    if n:
        with torch.cuda.device(d.device):
            _lib.check(L.svgir_env_taps(n, int(env_h), int(env_w), _p(_c(transform)), _p(d), _p(taps), _stream(d.device)),
                       "env_taps")
    return taps


def _cached_taps(dirs: Optional[torch.Tensor], env_h: int, env_w: int, transform: Optional[torch.Tensor]):
    if not ENV_TAP_CACHE or dirs is None or not dirs.is_cuda or dirs.dtype != torch.float32 or not dirs.is_contiguous():
        return None
    key = (dirs.data_ptr(), tuple(dirs.shape), int(env_h), int(env_w), None if transform is None else transform.data_ptr())
    ver = (dirs._version, None if transform is None else transform._version)
    ent = _TAP_CACHE.get(key)
    if ent is not None and ent[1] == ver:
        _TAP_CACHE.move_to_end(key)
        return ent[0]
    taps = env_taps(dirs, env_h, env_w, transform, out=None if ent is None else ent[0])
    _TAP_CACHE[key] = [taps, ver, dirs, transform]     # the entry keeps `dirs` alive: its address cannot be recycled
    _TAP_CACHE.move_to_end(key)
    while len(_TAP_CACHE) > _TAP_CACHE_SIZE:
        _TAP_CACHE.popitem(last=False)
    return taps


def clear_env_tap_cache():
    """Drops the cached taps (each entry keeps its directions tensor and a same-sized tap buffer alive)."""
    _TAP_CACHE.clear()


def refresh_env_taps(dirs: torch.Tensor, env_h: int, env_w: int, transform: Optional[torch.Tensor] = None):
    """Recomputes the cached taps of `dirs` in place if `dirs` (or the transform) was modified since; returns them."""
    return _cached_taps(dirs, env_h, env_w, transform)

MODE_LEARNABLE = 0  # softplus(param), x2   (DirectLightMap)
MODE_FIXED = 1      # linear map as given, x1 (EnvLight after its 32x64 resize)


def env_of(light):
    """(env tensor [He,We,3], mode, transform) of one of the reference's light objects, or of a
    (tensor, mode[, transform]) tuple."""
    if isinstance(light, (tuple, list)):
        env, mode = light[0], light[1]
        tr = light[2] if len(light) > 2 else None
    elif hasattr(light, "env"):  # DirectLightMap: nn.Parameter [1,H,W,3] (direct_light_map.py:11-16)
        env, mode, tr = light.env, MODE_LEARNABLE, None
    elif hasattr(light, "envmap"):  # EnvLight: [H,W,3] resized to 32x64 at every lookup (envmap.py:63-64)
        e = light.envmap.permute(2, 0, 1).unsqueeze(0)
        e = F.interpolate(e, size=(32, 64), mode="bilinear", align_corners=False)
        env, mode, tr = e[0].permute(1, 2, 0), MODE_FIXED, getattr(light, "transform", None)
    else:
        raise TypeError("unsupported env light object: %r" % type(light))
    if env.dim() == 4:
        env = env[0]
    return env, mode, tr


class _ShadeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, base_color, roughness, normals, viewdirs, radiance, visibility, incident_dirs,
                incident_areas, env, metallic, env_mode, transform, debug):
        if not base_color.is_cuda:
            raise RuntimeError("svgir_b200 shading needs CUDA tensors (no CPU fallback)")
        L = _L()
        dev = base_color.device
        N, Ns = incident_dirs.shape[0], incident_dirs.shape[1]
        t = dict(base_color=_c(base_color), roughness=_c(roughness), normals=_c(normals), viewdirs=_c(viewdirs),
                 radiance=_c(radiance), visibility=_c(visibility), incident_dirs=_c(incident_dirs),
                 incident_areas=_c(incident_areas), env=_c(env), metallic=_c(metallic), transform=_c(transform))
        He, We = t["env"].shape[0], t["env"].shape[1]
        f32 = dict(dtype=torch.float32, device=dev)
        outs = [torch.empty((N, 12), **f32) for _ in range(5)]
        mv, ml, mi, mg = (torch.empty((N, 1), **f32), torch.empty((N, 3), **f32), torch.empty((N, 3), **f32),
                          torch.empty((N, 3), **f32))
        scratch = torch.empty((He, We, 3), **f32)
        cfg = ShadeCfg(N, Ns, He, We, int(env_mode), int(bool(debug)))
        cin = ShadeIn(_p(t["base_color"]), _p(t["roughness"]), _p(t["metallic"]), _p(t["normals"]), _p(t["viewdirs"]),
                      _p(t["radiance"]), _p(t["visibility"]), _p(t["incident_dirs"]), _p(t["incident_areas"]),
                      _p(t["env"]), _p(t["transform"]), _p(scratch), None)
        cin.env_taps = _p(_cached_taps(t["incident_dirs"], He, We, t["transform"]))
        sums = torch.empty((2, N, 12), **f32)
        cout = ShadeOut(*[_p(o) for o in outs], _p(mv), _p(ml), _p(mi), _p(mg), None, _p(sums[0]), _p(sums[1]), 0, 0, 0, 0)
        if N > 0:
            with torch.cuda.device(dev):
                _lib.check(L.svgir_shade_forward(C.byref(cfg), C.byref(cin), C.byref(cout), _stream(dev)), "shade_forward")
        ctx.cfg = (N, Ns, He, We, int(env_mode), int(bool(debug)))
        ctx.has = (metallic is not None, transform is not None)
        ctx.save_for_backward(*[x for x in (t["base_color"], t["roughness"], t["normals"], t["viewdirs"], t["radiance"],
                                            t["visibility"], t["incident_dirs"], t["incident_areas"], t["env"], sums,
                                            t["metallic"], t["transform"]) if x is not None])
        return (*outs, mv, ml, mi, mg)

    @staticmethod
    def backward(ctx, g_pbr, g_diff, g_spec, g_dir, g_ind, g_mv, g_ml, g_mi, g_mg):
        L = _L()
        saved = list(ctx.saved_tensors)
        base_color, roughness, normals, viewdirs, radiance, visibility, dirs, areas, env, sums = saved[:10]
        rest = saved[10:]
        metallic = rest.pop(0) if ctx.has[0] else None
        transform = rest.pop(0) if ctx.has[1] else None
        N, Ns, He, We, env_mode, debug = ctx.cfg
        dev = base_color.device
        f32 = dict(dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad
        d_base = torch.empty((N, 12), **f32)
        d_rough = torch.empty((N, 4), **f32)
        d_norm = torch.empty((N, 4, 3), **f32)
        d_view = torch.empty((N, 3), **f32)
        d_rad = torch.empty((N, Ns, 3), **f32) if need[4] else None
        d_vis = torch.empty(tuple(visibility.shape), **f32) if need[5] else None
        d_env = torch.zeros((He, We, 3), **f32) if need[8] else None
        d_met = torch.empty((N, 4), **f32) if (metallic is not None and need[9]) else None
        scratch = torch.empty((He, We, 3), **f32)
        gs = [_c(x) for x in (g_pbr, g_diff, g_spec, g_dir, g_ind, g_mv, g_ml, g_mi, g_mg)]
        cfg = ShadeCfg(N, Ns, He, We, env_mode, debug)
        cin = ShadeIn(_p(base_color), _p(roughness), _p(metallic), _p(normals), _p(viewdirs), _p(radiance),
                      _p(visibility), _p(dirs), _p(areas), _p(env), _p(transform), _p(scratch), None)
        cin.env_taps = _p(_cached_taps(dirs, He, We, transform))
        cg = ShadeGrads(*[_p(x) for x in gs], _p(d_base), _p(d_rough), _p(d_met), _p(d_norm), _p(d_view), _p(d_rad),
                        _p(d_vis), _p(d_env), None, _p(sums[0]), _p(sums[1]),
                        _p(torch.empty((ENV_COPIES, He, We, 4), **f32)) if d_env is not None else None, 0, 0, 0, 0)
        if N > 0:
            with torch.cuda.device(dev):
                _lib.check(L.svgir_shade_backward(C.byref(cfg), C.byref(cin), C.byref(cg), _stream(dev)), "shade_backward")
        if need[6] or need[7]:
            raise NotImplementedError("svgir_b200 shading: gradients w.r.t. incident_dirs / incident_areas are not "
                                      "provided (they are precomputed buffers in the reference, gaussian_model.py:462-463)")
        return (d_base, d_rough, d_norm, d_view, d_rad, d_vis, None, None, d_env, d_met, None, None, None)


def shade_surfels(base_color, roughness, normals, viewdirs, radiance, env_light, visibility, incident_dirs,
                  incident_areas, metallic=None, debug=False) -> dict:
    """Fused render_equation. Shapes as in the reference: base_color [N,12] (channel-major),
    roughness [N,4], normals [N,4,3], viewdirs [N,3], radiance [N,Ns,3], visibility [N,Ns,1],
    incident_dirs [N,Ns,3], incident_areas [N,Ns,1]. Returns a dict of [N,*] tensors."""
    env, mode, tr = env_of(env_light)
    o = _ShadeFn.apply(base_color, roughness, normals, viewdirs, radiance, visibility, incident_dirs,
                       incident_areas, env, metallic, mode, tr, debug)
    keys = ("pbr", "diffuse_light", "specular", "direct", "indirect", "mean_visibility", "mean_local_lights",
            "mean_incident_lights", "mean_global_lights")
    return dict(zip(keys, o))


class _ShadePackedFn(torch.autograd.Function):
    """Shading fused with render_view's packing (gaussian_renderer/svgss.py:141-166): the kernel writes
    `features` [N,S] and `vfeatures` [N,VS] in the rasteriser's layout directly -- train: S=4
    [mean visibility, mean local light], VS=52 [pbr, base colour, view-space normal, roughness, diffuse
    light]; eval: S=7 [mean light, mean local light, mean visibility], VS=64 [.., direct, indirect] --
    and the backward kernel consumes their gradients with row strides, so no torch cat / split / matmul
    runs on the [N,*] tensors."""

    @staticmethod
    def forward(ctx, base_color, roughness, normals, viewdirs, radiance, visibility, incident_dirs,
                incident_areas, env, metallic, view3x3, env_mode, transform, is_training, debug, work=None,
                means3D=None, campos=None):
        if not base_color.is_cuda:
            raise RuntimeError("svgir_b200 shading needs CUDA tensors (no CPU fallback)")
        L = _L()
        dev = base_color.device
        N, Ns = incident_dirs.shape[0], incident_dirs.shape[1]
        if viewdirs is None and (means3D is None or campos is None):
            raise ValueError("shade_and_pack: pass viewdirs, or means3D + campos (the kernel then evaluates "
                             "normalize(campos - means3D), gaussian_renderer/svgss.py:95)")
        if work is not None:
            wl, wc = work
            if wl.dtype != torch.int32 or wc.dtype != torch.int32 or wl.numel() < N or not wl.is_contiguous():
                raise RuntimeError("shade_and_pack: work list must be contiguous int32 [N] plus an int32 [1] count")
        alloc = torch.zeros if work is not None else torch.empty   # rows outside the work list stay zero
        t = dict(base_color=_c(base_color), roughness=_c(roughness), normals=_c(normals), viewdirs=_c(viewdirs),
                 radiance=_c(radiance), visibility=_c(visibility), incident_dirs=_c(incident_dirs),
                 incident_areas=_c(incident_areas), env=_c(env), metallic=_c(metallic), transform=_c(transform),
                 view3x3=_c(view3x3), means3D=_c(means3D) if viewdirs is None else None,
                 campos=_c(campos) if viewdirs is None else None)
        He, We = t["env"].shape[0], t["env"].shape[1]
        f32 = dict(dtype=torch.float32, device=dev)
        S, VS = (4, 52) if is_training else (7, 64)
        feats = alloc((N, S), **f32)
        vfeats = alloc((N, VS), **f32)
        scratch = torch.empty((He, We, 3), **f32)
        cfg = ShadeCfg(N, Ns, He, We, int(env_mode), int(bool(debug)))
        cin = ShadeIn(_p(t["base_color"]), _p(t["roughness"]), _p(t["metallic"]), _p(t["normals"]), _p(t["viewdirs"]),
                      _p(t["radiance"]), _p(t["visibility"]), _p(t["incident_dirs"]), _p(t["incident_areas"]),
                      _p(t["env"]), _p(t["transform"]), _p(scratch), _p(t["view3x3"]),
                      _p(work[0]) if work is not None else None, _p(work[1]) if work is not None else None,
                      _p(t["means3D"]), _p(t["campos"]), None)
        cin.env_taps = _p(_cached_taps(t["incident_dirs"], He, We, t["transform"]))
        vp, fp = vfeats.data_ptr(), feats.data_ptr()
        # sums saved for backward: un-split total in training (only pbr / diffuse carry gradients there)
        sums = torch.empty((1 if is_training else 2, N, 12), **f32)
        if is_training:
            cout = ShadeOut(vp, vp + 4 * 40, None, None, None, fp, fp + 4, None, None, vp + 4 * 12, _p(sums[0]), None,
                            VS, S, S, 0)
        else:
            cout = ShadeOut(vp, None, None, vp + 4 * 40, vp + 4 * 52, fp + 4 * 6, fp + 4 * 3, fp, None, vp + 4 * 12,
                            _p(sums[0]), _p(sums[1]), VS, S, S, 0)
        if N > 0:
            with torch.cuda.device(dev):
                _lib.check(L.svgir_shade_forward(C.byref(cfg), C.byref(cin), C.byref(cout), _stream(dev)), "shade_forward")
        ctx.cfg = (N, Ns, He, We, int(env_mode), int(bool(debug)), bool(is_training))
        ctx.has = (metallic is not None, transform is not None)
        ctx.work = work
        ctx.fused_view = viewdirs is None
        vsrc = t["viewdirs"] if viewdirs is not None else t["means3D"]
        ctx.save_for_backward(*[x for x in (t["base_color"], t["roughness"], t["normals"], vsrc, t["radiance"],
                                            t["visibility"], t["incident_dirs"], t["incident_areas"], t["env"],
                                            t["view3x3"], sums, t["metallic"], t["transform"], t["campos"]) if x is not None])
        return feats, vfeats

    @staticmethod
    def backward(ctx, g_feats, g_vfeats):
        L = _L()
        saved = list(ctx.saved_tensors)
        base_color, roughness, normals, viewdirs, radiance, visibility, dirs, areas, env, view3x3, sums = saved[:11]
        rest = saved[11:]
        metallic = rest.pop(0) if ctx.has[0] else None
        transform = rest.pop(0) if ctx.has[1] else None
        means3D = campos = None
        if ctx.fused_view:   # the fourth saved tensor is means3D; the view direction is evaluated in the kernel
            means3D, viewdirs, campos = viewdirs, None, rest.pop(0)
        N, Ns, He, We, env_mode, debug, is_training = ctx.cfg
        dev = base_color.device
        f32 = dict(dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad
        S, VS = (4, 52) if is_training else (7, 64)
        g_feats = _c(g_feats) if g_feats is not None else torch.zeros((N, S), **f32)
        g_vfeats = _c(g_vfeats) if g_vfeats is not None else torch.zeros((N, VS), **f32)
        work = ctx.work
        alloc = torch.zeros if work is not None else torch.empty   # surfels outside the work list: zero gradient
        d_base = alloc((N, 12), **f32)
        d_rough = alloc((N, 4), **f32)
        d_norm = alloc((N, 4, 3), **f32)
        d_view = alloc((N, 3), **f32) if not ctx.fused_view else None
        d_xyz = torch.zeros((N, 3), **f32) if ctx.fused_view else None   # the kernel ADDS -d/d(campos - means3D)
        d_rad = alloc((N, Ns, 3), **f32) if need[4] else None
        d_vis = alloc(tuple(visibility.shape), **f32) if need[5] else None
        d_env = torch.zeros((He, We, 3), **f32) if need[8] else None
        d_met = alloc((N, 4), **f32) if (metallic is not None and need[9]) else None
        scratch = torch.empty((He, We, 3), **f32)
        cfg = ShadeCfg(N, Ns, He, We, env_mode, debug)
        cin = ShadeIn(_p(base_color), _p(roughness), _p(metallic), _p(normals), _p(viewdirs), _p(radiance),
                      _p(visibility), _p(dirs), _p(areas), _p(env), _p(transform), _p(scratch), _p(view3x3),
                      _p(work[0]) if work is not None else None, _p(work[1]) if work is not None else None,
                      _p(means3D), _p(campos), None)
        cin.env_taps = _p(_cached_taps(dirs, He, We, transform))
        vp, fp = g_vfeats.data_ptr(), g_feats.data_ptr()
        if is_training:
            gin = (vp, vp + 4 * 40, None, None, None, fp, fp + 4, None, None)
        else:
            gin = (vp, None, None, vp + 4 * 40, vp + 4 * 52, fp + 4 * 6, fp + 4 * 3, fp, None)
        cg = ShadeGrads(*gin, _p(d_base), _p(d_rough), _p(d_met), _p(d_norm), _p(d_view), _p(d_rad), _p(d_vis),
                        _p(d_env), vp + 4 * 12, _p(sums[0]), _p(sums[1]) if sums.shape[0] > 1 else None,
                        _p(torch.empty((ENV_COPIES, He, We, 4), **f32)) if d_env is not None else None, VS, S, S, 0, _p(d_xyz))
        if N > 0:
            with torch.cuda.device(dev):
                _lib.check(L.svgir_shade_backward(C.byref(cfg), C.byref(cin), C.byref(cg), _stream(dev)), "shade_backward")
        if need[6] or need[7]:
            raise NotImplementedError("svgir_b200 shading: no gradients w.r.t. incident_dirs / incident_areas")
        return (d_base, d_rough, d_norm, d_view, d_rad, d_vis, None, None, d_env, d_met, None, None, None, None, None, None,
                d_xyz, None)


def shade_and_pack(base_color, roughness, normals, viewdirs, radiance, env_light, visibility, incident_dirs,
                   incident_areas, view3x3, is_training=True, metallic=None, debug=False, work=None,
                   means3D=None, campos=None):
    """(features [N,S], vfeatures [N,VS]) exactly as render_view packs them (svgss.py:141-166), computed by
    one fused kernel. view3x3 = world_view_transform[:3,:3].
    work = (surfel_list int32 [N], count int32 [1]) restricts shading (and its backward) to the listed
    surfels -- e.g. the rasteriser's list of surfels that survive culling; all other rows are zero.
    viewdirs=None with means3D [N,3] + campos [3]: the kernels evaluate normalize(campos - means3D) (svgss.py:95)
    themselves and the backward returns the resulting gradient for means3D (no torch normalise / backward kernels)."""
    env, mode, tr = env_of(env_light)
    return _ShadePackedFn.apply(base_color, roughness, normals, viewdirs, radiance, visibility, incident_dirs,
                                incident_areas, env, metallic, view3x3, mode, tr, bool(is_training), debug, work,
                                means3D, campos)


class _DirectLightFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, env, dirs, env_mode, transform):
        L = _L()
        dev = dirs.device
        env_c, dirs_c, tr = _c(env), _c(dirs.reshape(-1, 3)), _c(transform)
        n = dirs_c.shape[0]
        He, We = env_c.shape[0], env_c.shape[1]
        out = torch.empty((n, 3), dtype=torch.float32, device=dev)
        scratch = torch.empty((He, We, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.svgir_direct_light_forward(n, He, We, int(env_mode), _p(env_c), _p(scratch), _p(tr), _p(dirs_c),
                                                    _p(out), _stream(dev)), "direct_light_forward")
        ctx.meta = (n, He, We, int(env_mode), tr is not None)
        ctx.save_for_backward(*[x for x in (env_c, dirs_c, tr) if x is not None])
        return out.reshape(dirs.shape)

    @staticmethod
    def backward(ctx, g):
        L = _L()
        n, He, We, mode, has_tr = ctx.meta
        sv = list(ctx.saved_tensors)
        env_c, dirs_c = sv[0], sv[1]
        tr = sv[2] if has_tr else None
        d_env = None
        if ctx.needs_input_grad[0]:
            d_env = torch.zeros((He, We, 3), dtype=torch.float32, device=g.device)
            gc = _c(g.reshape(-1, 3))
            with torch.cuda.device(g.device):
                _lib.check(L.svgir_direct_light_backward(n, He, We, mode, _p(env_c), _p(tr), _p(dirs_c), _p(gc), _p(d_env),
                                                         _stream(g.device)), "direct_light_backward")
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("svgir_b200.direct_light: no gradient w.r.t. directions")
        return d_env, None, None, None


def direct_light(env_light, dirs, transform=None):
    """Env radiance for arbitrary directions [...,3] -> [...,3]."""
    env, mode, tr = env_of(env_light)
    if transform is not None:
        tr = transform
    if not dirs.is_cuda:
        raise RuntimeError("svgir_b200.direct_light needs CUDA tensors (no CPU fallback)")
    return _DirectLightFn.apply(env, dirs, mode, tr)


class _LazyResults(dict):
    """extra_results of rendering_equation4. The [N,Ns,3] per-sample light tensors the reference
    materialises eagerly (svgss.py:544-553) are only built when somebody reads them."""

    def __init__(self, eager: dict, lazy: dict):
        super().__init__(eager)
        self._lazy = lazy

    def __missing__(self, key):
        if key in self._lazy:
            val = self._lazy[key]()
            self[key] = val
            return val
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def keys(self):
        return list(dict.keys(self)) + [k for k in self._lazy if not dict.__contains__(self, k)]


def rendering_equation4(base_color, roughness, normals, viewdirs, radiance, direct_light_env_light=None,
                        visibility_precompute=None, incident_dirs_precompute=None,
                        incident_areas_precompute=None, metallic=None):
    """Drop-in for gaussian_renderer/svgss.py:537-593 (same arguments, same `(pbr, extra_results)`)."""
    r = shade_surfels(base_color, roughness, normals, viewdirs, radiance, direct_light_env_light,
                      visibility_precompute, incident_dirs_precompute, incident_areas_precompute, metallic)

    def global_lights():
        return direct_light(direct_light_env_light, incident_dirs_precompute).clamp(0, 64) * visibility_precompute

    def incident_lights():
        return radiance + extra["global_incident_lights"]

    extra = _LazyResults(
        {"incident_dirs": incident_dirs_precompute, "local_incident_lights": radiance,
         "incident_visibility": visibility_precompute, "diffuse_light": r["diffuse_light"],
         "specular": r["specular"], "direct": r["direct"], "indirect": r["indirect"],
         # fused sample means (what render_view computes with .mean(-2), svgss.py:149-156)
         "mean_visibility": r["mean_visibility"], "mean_local_lights": r["mean_local_lights"],
         "mean_incident_lights": r["mean_incident_lights"]},
        {"global_incident_lights": global_lights, "incident_lights": incident_lights})
    return r["pbr"], extra
