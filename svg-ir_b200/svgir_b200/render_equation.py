"""SH-lit render_equation operators (csrc/render_equation_sh.cu).

Same names, argument order and result tuples as the pybind functions the reference declares in
rgss-rasterization/render_equation.h:7-46 (defined in render_equation.cu, never built by the
reference's setup.py -- SURVEY.md 8(a) a19):

  render_equation_forward(base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs,
                          visibility_shs, sample_num, is_training, debug) -> (pbr, incident_dirs, diffuse_light)
  render_equation_forward_complex(... , sample_num) -> 11-tuple
  render_equation_backward(base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs,
                           visibility_shs, sample_num, incident_dirs, dL_dpbr, dL_ddiffuse_light, debug) -> 8-tuple

plus `render_equation(...)`, an autograd wrapper in the style of the reference's other CUDA operators.
`legacy_exact=True` (default) reproduces the reference backward arithmetic including its slips; see
include/svgir_b200.h.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

c_fp = C.c_void_p


class ReqCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("P", "S_incident", "S_direct", "S_vis", "sample_num", "is_training",
                                         "legacy_exact", "debug")]


class ReqIn(C.Structure):
    _fields_ = [(n, c_fp) for n in ("base_color", "roughness", "metallic", "normals", "viewdirs", "incidents_shs",
                                    "direct_shs", "visibility_shs", "rand_float")]


class ReqOut(C.Structure):
    _fields_ = [(n, c_fp) for n in ("pbr", "incident_dirs", "diffuse_light", "incident_lights", "local_incident_lights",
                                    "global_incident_lights", "incident_visibility", "local_diffuse_light", "accum",
                                    "rgb_d", "rgb_s")]


class ReqGrads(C.Structure):
    _fields_ = [(n, c_fp) for n in ("incident_dirs", "dL_dpbr", "dL_ddiffuse_light", "dL_dbase_color", "dL_droughness",
                                    "dL_dmetallic", "dL_dnormals", "dL_dviewdirs", "dL_dincidents_shs", "dL_ddirect_shs",
                                    "dL_dvisibility_shs")]


_bound = False


def _L():
    global _bound
    L = _lib.lib()
    if not _bound:
        L.svgir_render_equation_sh_forward.argtypes = [C.POINTER(ReqCfg), C.POINTER(ReqIn), C.POINTER(ReqOut), C.c_void_p]
        L.svgir_render_equation_sh_forward.restype = C.c_int
        L.svgir_render_equation_sh_backward.argtypes = [C.POINTER(ReqCfg), C.POINTER(ReqIn), C.POINTER(ReqGrads), C.c_void_p]
        L.svgir_render_equation_sh_backward.restype = C.c_int
        _bound = True
    return L


def _c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _p(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _inputs(base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs, visibility_shs):
    if not base_color.is_cuda:
        raise RuntimeError("svgir_b200.render_equation needs CUDA tensors (no CPU fallback)")
    t = [_c(x) for x in (base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs, visibility_shs)]
    P = t[0].shape[0]
    return t, P, t[5].shape[1], t[6].shape[1], t[7].shape[1]


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def render_equation_forward(base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs,
                            visibility_shs, sample_num, is_training=False, debug=False, rand_float=None):
    """RenderEquationForwardCUDA (render_equation.cu:681-729). `rand_float` [P,sample_num,1] overrides the
    torch.rand draw the reference makes internally (:705)."""
    t, P, Si, Sd, Sv = _inputs(base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs, visibility_shs)
    dev = t[0].device
    f32 = dict(dtype=torch.float32, device=dev)
    pbr = torch.zeros((P, 3), **f32)
    dirs = torch.zeros((P, sample_num, 3), **f32)
    diffuse = torch.zeros((P, 3), **f32)
    rnd = rand_float if rand_float is not None else torch.rand((P, sample_num, 1), **f32)
    rnd = _c(rnd)
    cfg = ReqCfg(P, Si, Sd, Sv, int(sample_num), int(bool(is_training)), 1, int(bool(debug)))
    cin = ReqIn(*[_p(x) for x in t], _p(rnd))
    cout = ReqOut(_p(pbr), _p(dirs), _p(diffuse), *([None] * 8))
    if P:
        with torch.cuda.device(dev):
            _lib.check(_L().svgir_render_equation_sh_forward(C.byref(cfg), C.byref(cin), C.byref(cout), _stream(dev)),
                       "render_equation_forward")
    return pbr, dirs, diffuse


def render_equation_forward_complex(base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs,
                                    visibility_shs, sample_num):
    """RenderEquationForwardCUDA_complex (render_equation.cu:219-277): (pbr, incident_dirs, incident_lights,
    local_incident_lights, global_incident_lights, incident_visibility, diffuse_light, local_diffuse_light,
    accum, rgb_d, rgb_s)."""
    t, P, Si, Sd, Sv = _inputs(base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs, visibility_shs)
    dev = t[0].device
    f32 = dict(dtype=torch.float32, device=dev)
    z = lambda *s: torch.zeros(s, **f32)
    pbr, dirs = z(P, 3), z(P, sample_num, 3)
    il, ll, gl, iv = z(P, sample_num, 3), z(P, sample_num, 3), z(P, sample_num, 3), z(P, sample_num, 1)
    diffuse, ldiffuse, accum, rgb_d, rgb_s = z(P, 3), z(P, 3), z(P, 1), z(P, 3), z(P, 3)
    cfg = ReqCfg(P, Si, Sd, Sv, int(sample_num), 0, 1, 0)
    cin = ReqIn(*[_p(x) for x in t], None)
    cout = ReqOut(_p(pbr), _p(dirs), _p(diffuse), _p(il), _p(ll), _p(gl), _p(iv), _p(ldiffuse), _p(accum), _p(rgb_d), _p(rgb_s))
    if P:
        with torch.cuda.device(dev):
            _lib.check(_L().svgir_render_equation_sh_forward(C.byref(cfg), C.byref(cin), C.byref(cout), _stream(dev)),
                       "render_equation_forward_complex")
    return pbr, dirs, il, ll, gl, iv, diffuse, ldiffuse, accum, rgb_d, rgb_s


def render_equation_backward(base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs,
                             visibility_shs, sample_num, incident_dirs, dL_dpbr, dL_ddiffuse_light, debug=False,
                             legacy_exact=True):
    """RenderEquationBackwardCUDA (render_equation.cu:494-550): (dL_dbase_color, dL_droughness, dL_dmetallic,
    dL_dnormals, dL_dviewdirs, dL_dincidents_shs, dL_ddirect_shs, dL_dvisibility_shs)."""
    t, P, Si, Sd, Sv = _inputs(base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs, visibility_shs)
    dev = t[0].device
    f32 = dict(dtype=torch.float32, device=dev)
    z = lambda *s: torch.zeros(s, **f32)
    g = (z(P, 3), z(P, 1), z(P, 1), z(P, 3), z(P, 3), z(P, Si, 3), z(1, Sd, 3), z(P, Sv, 1))
    dirs, gp, gd = _c(incident_dirs), _c(dL_dpbr), _c(dL_ddiffuse_light)
    cfg = ReqCfg(P, Si, Sd, Sv, int(sample_num), 0, int(bool(legacy_exact)), int(bool(debug)))
    cin = ReqIn(*[_p(x) for x in t], None)
    cg = ReqGrads(_p(dirs), _p(gp), _p(gd), *[_p(x) for x in g])
    if P:
        with torch.cuda.device(dev):
            _lib.check(_L().svgir_render_equation_sh_backward(C.byref(cfg), C.byref(cin), C.byref(cg), _stream(dev)),
                       "render_equation_backward")
    return g


class _RenderEquation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs, visibility_shs,
                sample_num, is_training, debug, legacy_exact):
        pbr, dirs, diffuse = render_equation_forward(base_color, roughness, metallic, normals, viewdirs, incidents_shs,
                                                     direct_shs, visibility_shs, sample_num, is_training, debug)
        ctx.meta = (sample_num, debug, legacy_exact)
        ctx.save_for_backward(base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs,
                              visibility_shs, dirs)
        ctx.mark_non_differentiable(dirs)
        return pbr, dirs, diffuse

    @staticmethod
    def backward(ctx, g_pbr, g_dirs, g_diffuse):
        sample_num, debug, legacy_exact = ctx.meta
        *ins, dirs = ctx.saved_tensors
        P = ins[0].shape[0]
        z = lambda: torch.zeros((P, 3), dtype=torch.float32, device=ins[0].device)
        g = render_equation_backward(*ins, sample_num, dirs, g_pbr if g_pbr is not None else z(),
                                     g_diffuse if g_diffuse is not None else z(), debug, legacy_exact)
        return (*g, None, None, None, None)


def render_equation(base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs, visibility_shs,
                    sample_num=24, is_training=False, debug=False, legacy_exact=True):
    """Differentiable (pbr [P,3], incident_dirs [P,Ns,3], diffuse_light [P,3])."""
    return _RenderEquation.apply(base_color, roughness, metallic, normals, viewdirs, incidents_shs, direct_shs,
                                 visibility_shs, int(sample_num), bool(is_training), bool(debug), bool(legacy_exact))
