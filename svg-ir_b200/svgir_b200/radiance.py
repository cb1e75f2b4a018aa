"""Radiance cache and radiance-consistency loss (C ABI: svgir_radiance_* in include/svgir_b200.h).

Host-side mirror of the reference's stage-2 global-illumination regulariser:
  RadianceCache.update(...)   GaussianModel.update_radiace (scene/gaussian_model.py:469-522): sample the incident
                              hemisphere of every surfel, trace each ray from hit to hit through the surfel tree,
                              cache radiance / visibility / first hit / uv
                              (Renderer.render_radiance_with_sampling_SH, pbgi/renderer.py:596-615)
  RadianceCache.loss(...)     GaussianModel.get_radiance_loss (scene/gaussian_model.py:544-575): L1 between the one-bounce
                              irradiance of the most view-reflective occluded sample (render_irradiance_sample,
                              pbgi/renderer.py:181-227, 748-751) and the cached radiance, differentiable w.r.t. albedo,
                              roughness and the env map
The functional forms render_radiance_with_sampling_SH / radiance_loss take the tensors directly. PyTorch provides
device memory, the stream and autograd plumbing; all arithmetic is in libsvgir_b200.so (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib, bvh as _bvh, sampling as _sampling
from .shading import env_of

RADIANCE_ENV_READY = 1
RADIANCE_BWD_REFERENCE_GRID = 2
RADIANCE_NORMALS_VERTEX_MAJOR = 4
RECORD_FLOATS = 32
SCRATCH_FLOATS = 2048
ENV_COPIES = 32


class RadianceLossCfg(C.Structure):
    _fields_ = [("P", C.c_int32), ("S", C.c_int32), ("env_h", C.c_int32), ("env_w", C.c_int32), ("env_mode", C.c_int32),
                ("flags", C.c_int32), ("rough_stride", C.c_int32), ("reserved_", C.c_int32)]


class RadianceLossIn(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "means3D", "campos", "geo_normal", "incident_dirs", "incident_areas", "visibility", "hit_index", "uv", "radiances",
        "radiance_ratio", "normals", "albedo", "roughness", "env", "env_act_scratch", "skip_flag", "env_taps")]


_BOUND = False


def _L():
    global _BOUND
    L = _lib.lib()
    if not _BOUND:
        vp = C.c_void_p
        L.svgir_radiance_pack_surfels.argtypes = [C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
        L.svgir_radiance_pack_surfels.restype = C.c_int
        L.svgir_radiance_cache_build.argtypes = [C.POINTER(_bvh.BvhStruct), C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp,
                                                 vp, vp, vp, vp, vp]
        L.svgir_radiance_cache_build.restype = C.c_int
        L.svgir_radiance_loss_forward.argtypes = [C.POINTER(RadianceLossCfg), C.POINTER(RadianceLossIn), vp, vp, vp, vp, vp, vp]
        L.svgir_radiance_loss_forward.restype = C.c_int
        L.svgir_radiance_loss_backward.argtypes = [C.POINTER(RadianceLossCfg), C.POINTER(RadianceLossIn), vp, vp, vp, vp, vp,
                                                   vp, vp, vp]
        L.svgir_radiance_loss_backward.restype = C.c_int
        L.svgir_radiance_loss_forward_backward.argtypes = [C.POINTER(RadianceLossCfg), C.POINTER(RadianceLossIn)] + [vp] * 11
        L.svgir_radiance_loss_forward_backward.restype = C.c_int
        _BOUND = True
    return L


def _f(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("svgir_b200.radiance needs CUDA tensors (no CPU fallback)")
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def pack_surfels(xyz, scaling, rotation, normals, opacity, inverse_covariance) -> torch.Tensor:
    """[P,32] closest-hit records (svgir_radiance_pack_surfels)."""
    L = _L()
    xyz, scaling, rotation, normals = _f(xyz), _f(scaling), _f(rotation), _f(normals)
    opacity, ci = _f(opacity).reshape(-1), _f(inverse_covariance)
    P = xyz.shape[0]
    if scaling.shape[0] != P or scaling.dim() != 2 or scaling.shape[1] < 2 or rotation.shape != (P, 4) or \
            normals.shape != (P, 3) or opacity.numel() != P or ci.shape != (P, 6):
        raise RuntimeError("pack_surfels: xyz [P,3], scaling [P,>=2], rotation [P,4], normals [P,3], opacity [P], "
                           "inverse_covariance [P,6] expected")
    rec = torch.empty((P, RECORD_FLOATS), dtype=torch.float32, device=xyz.device)
    if P:
        with torch.cuda.device(xyz.device):
            _lib.check(L.svgir_radiance_pack_surfels(P, xyz.data_ptr(), scaling.data_ptr(), int(scaling.shape[1]),
                                                     rotation.data_ptr(), normals.data_ptr(), opacity.data_ptr(),
                                                     ci.data_ptr(), rec.data_ptr(), _stream(xyz.device)),
                       "radiance_pack_surfels")
    return rec


@torch.no_grad()
def render_radiance_with_sampling_SH(tree: _bvh.Bvh, records: torch.Tensor, features: torch.Tensor, ray_o: torch.Tensor,
                                     ray_d: torch.Tensor, sample_num: Optional[int] = None, first_index: int = 0,
                                     self_mod: int = 0):
    """Renderer.render_radiance_with_sampling_SH (pbgi/renderer.py:596-615): ray_o [N,3], ray_d [N,S,3] ->
    (radiance [N,S,3], visibility [N,S,1], hit_indices [N,S,1] int32, uvs [N,S,2]). `features` [P,16,3]."""
    L = _L()
    ray_o, ray_d, shs = _f(ray_o), _f(ray_d), _f(features)
    N, S = ray_d.shape[0], ray_d.shape[1]
    if sample_num is not None and sample_num != S:
        raise RuntimeError("render_radiance_with_sampling_SH: sample_num does not match ray_d")
    if shs.shape[1:] != (16, 3) or shs.shape[0] != tree.P or records.shape != (tree.P, RECORD_FLOATS):
        raise RuntimeError("render_radiance_with_sampling_SH: features [P,16,3] and records [P,32] of the tree's surfels expected")
    dev = ray_d.device
    rad = torch.empty((N, S, 3), dtype=torch.float32, device=dev)
    vis = torch.empty((N, S, 1), dtype=torch.float32, device=dev)
    hit = torch.empty((N, S, 1), dtype=torch.int32, device=dev)
    uv = torch.empty((N, S, 2), dtype=torch.float32, device=dev)
    if N and S:
        with torch.cuda.device(dev):
            _lib.check(L.svgir_radiance_cache_build(C.byref(tree.c), N, S, int(first_index), int(self_mod), ray_o.data_ptr(),
                                                    ray_d.data_ptr(), records.data_ptr(), shs.data_ptr(), rad.data_ptr(),
                                                    vis.data_ptr(), hit.data_ptr(), uv.data_ptr(), _stream(dev)),
                       "radiance_cache_build")
    return rad, vis, hit, uv


class _RadianceLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, albedo, roughness, env, env_mode, flags, xyz, campos, geo_normal, dirs, areas, vis, hit, uv, radiances,
                ratio, normals):
        L = _L()
        dev = albedo.device
        P, S = hit.shape[0], hit.shape[1]
        t = dict(albedo=_f(albedo), roughness=_f(roughness), env=_f(env), xyz=_f(xyz), campos=_f(campos).reshape(-1),
                 geo_normal=_f(geo_normal), dirs=_f(dirs), areas=_f(areas), vis=_f(vis), uv=_f(uv), radiances=_f(radiances),
                 normals=_f(normals), ratio=None if ratio is None else _f(ratio).reshape(-1))
        hit_c = hit.detach().contiguous()
        if hit_c.dtype != torch.int32:
            hit_c = hit_c.int()
        if t["normals"].dim() == 3 and t["normals"].shape == (P, 4, 3):
            flags = int(flags) | RADIANCE_NORMALS_VERTEX_MAJOR          # get_shading_normal as it is
            t["normals"] = t["normals"].reshape(P, 12)
        if t["albedo"].shape != (P, 12) or t["normals"].shape != (P, 12) or t["roughness"].dim() != 2 or \
                t["roughness"].shape[0] != P or t["dirs"].shape != (P, S, 3) or t["radiances"].shape != (P, S, 3) or \
                t["uv"].numel() != P * S * 2 or t["vis"].numel() != P * S or t["areas"].numel() != P * S:
            raise RuntimeError("radiance_loss: shapes do not match [P,S] caches / [P,12] materials")
        He, We = t["env"].shape[0], t["env"].shape[1]
        cfg = RadianceLossCfg(P, S, He, We, int(env_mode), int(flags), int(t["roughness"].shape[1]), 0)
        scratch_env = torch.empty((He * We * 3,), dtype=torch.float32, device=dev)
        cin = RadianceLossIn(t["xyz"].data_ptr(), t["campos"].data_ptr(), t["geo_normal"].data_ptr(), t["dirs"].data_ptr(),
                             t["areas"].data_ptr(), t["vis"].data_ptr(), hit_c.data_ptr(), t["uv"].data_ptr(),
                             t["radiances"].data_ptr(), None if t["ratio"] is None else t["ratio"].data_ptr(),
                             t["normals"].data_ptr(), t["albedo"].data_ptr(), t["roughness"].data_ptr(), t["env"].data_ptr(),
                             scratch_env.data_ptr(), None, None)
        from . import shading as _sh
        taps = _sh.refresh_env_taps(dirs if dirs.is_contiguous() and dirs.dtype == torch.float32 else t["dirs"], He, We)
        cin.env_taps = None if taps is None else taps.data_ptr()
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        irr = torch.empty((P, 3), dtype=torch.float32, device=dev)
        sel = torch.empty((P,), dtype=torch.int32, device=dev)
        scratch = torch.empty((SCRATCH_FLOATS,), dtype=torch.float32, device=dev)
        saved = torch.empty((P, 8), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.svgir_radiance_loss_forward(C.byref(cfg), C.byref(cin), loss.data_ptr(), irr.data_ptr(), sel.data_ptr(),
                                                     saved.data_ptr(), scratch.data_ptr(), _stream(dev)), "radiance_loss_forward")
        ctx.cfg, ctx.cin, ctx.keep = cfg, cin, (t, hit_c, scratch_env, taps)
        ctx.save_for_backward(irr, saved)
        ctx.mark_non_differentiable(irr, sel)
        return loss[0], irr, sel

    @staticmethod
    def backward(ctx, g_loss, g_irr, _g_sel):
        L = _L()
        irr, saved = ctx.saved_tensors
        t, _hit, _scr, _taps = ctx.keep
        cfg = ctx.cfg
        dev = irr.device
        need_a, need_r, need_e = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        d_alb = torch.zeros_like(t["albedo"]) if need_a else None
        d_rough = torch.zeros_like(t["roughness"]) if need_r else None
        d_env = torch.zeros_like(t["env"]) if need_e else None
        d_env_scr = torch.empty((ENV_COPIES * cfg.env_h * cfg.env_w * 4,), dtype=torch.float32, device=dev) if need_e else None
        g = g_loss.detach().reshape(1).float().contiguous()
        p = lambda x: None if x is None else x.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(L.svgir_radiance_loss_backward(C.byref(cfg), C.byref(ctx.cin), g.data_ptr(), irr.data_ptr(),
                                                      saved.data_ptr(), p(d_alb), p(d_rough), p(d_env), p(d_env_scr),
                                                      _stream(dev)), "radiance_loss_backward")
        return (d_alb, d_rough, d_env) + (None,) * 13


def radiance_loss(cam_center, direct_light, xyz, geo_normal, incident_dirs, incident_areas, visibility, hit_indices, uvs,
                  radiances, radiance_ratio, normals, albedo, roughness, reference_backward_grid: bool = False,
                  return_aux: bool = False):
    """get_radiance_loss (scene/gaussian_model.py:544-575) on explicit tensors. `normals` [P,12] with element 4*c + v
    (get_shading_normal.transpose(1,2).reshape(P,-1), as the reference passes it) or [P,4,3] (get_shading_normal itself,
    read in place); `albedo` [P,12] element 4*c + v (get_albedo); roughness [P,V]; `radiances` already detached
    like GaussianModel.get_radiances does (:323-324); radiance_ratio a scalar tensor or None."""
    env, mode, tr = env_of(direct_light)
    if tr is not None:
        raise NotImplementedError("radiance_loss: env transform is not part of the training path")
    flags = RADIANCE_BWD_REFERENCE_GRID if reference_backward_grid else 0
    loss, irr, sel = _RadianceLossFn.apply(albedo, roughness, env, mode, flags, xyz, cam_center, geo_normal, incident_dirs,
                                           incident_areas, visibility, hit_indices.reshape(hit_indices.shape[0], -1), uvs,
                                           radiances, radiance_ratio, normals)
    return (loss, irr, sel) if return_aux else loss


class RadianceCache:
    """State that GaussianModel keeps for the regulariser (_visibility_tracing, _incident_dirs, _incident_areas,
    _radiances, _radiance_ratio, renderer.hemi_index_buffers, renderer.uv_buffers)."""

    def __init__(self):
        self.tracer: Optional[_bvh.RayTracer] = None
        self.visibility_tracing = self.incident_dirs = self.incident_areas = None
        self.radiances = self.init_radiances = self.radiance_mean = None
        self.hemi_index_buffers = self.uv_buffers = None
        self.radiance_ratio: Optional[torch.Tensor] = None
        self.geo_normal = None

    @torch.no_grad()
    def update(self, xyz, scaling, rotation, opacity, geo_normal, inverse_covariance, features, sample_num: int = 64,
               reference_chunking: bool = True):
        """update_radiace: rebuild the tree, draw sample_num fibonacci directions per surfel with a random azimuth offset,
        trace. reference_chunking=True draws the offsets chunk by chunk as the reference does (chunk =
        P // ((sample_num-1)//24+1), :487; same torch.rand calls, so a seeded run draws the same numbers) and reproduces
        its chunk-local self test; everything still runs as one launch."""
        P = xyz.shape[0]
        dev = xyz.device
        self.tracer = _bvh.RayTracer(xyz, scaling, rotation)
        self.geo_normal = geo_normal.detach()
        records = pack_surfels(xyz, scaling, rotation, geo_normal, opacity, inverse_covariance)
        chunk = P // ((sample_num - 1) // 24 + 1)
        if reference_chunking and 0 < chunk < P:
            u = torch.cat([torch.rand(min(chunk, P - o), 1, device=dev) for o in range(0, P, chunk)], 0)
            self_mod = chunk
        else:
            u = torch.rand(P, 1, device=dev)
            self_mod = 0
        dirs, areas = _sampling.fibonacci_sphere_sampling(geo_normal, sample_num, random_rotate=True, rand_u=u)
        rad, vis, hit, uv = render_radiance_with_sampling_SH(self.tracer.bvh, records, features, xyz, dirs, sample_num,
                                                             first_index=0, self_mod=self_mod)
        self.visibility_tracing, self.incident_dirs, self.incident_areas = vis, dirs, areas
        if self.radiances is None or not bool(torch.any(self.radiances)) or self.radiances.shape[1] != sample_num:
            self.radiances = rad                                   # :515-516
        self.init_radiances = rad.clone()
        self.radiance_mean = rad.mean()
        self.hemi_index_buffers, self.uv_buffers = hit, uv
        if self.radiance_ratio is None:
            self.radiance_ratio = torch.tensor(1.0, device=dev)
        return self

    @property
    def get_radiances(self):
        """:323-324"""
        return torch.nan_to_num(self.radiances.detach() * self.radiance_ratio, nan=0.0)

    def loss(self, cam_center, direct_light, xyz, geo_normal, shading_normal, albedo, roughness, **kw):
        """get_radiance_loss. shading_normal [P,V,3] (get_shading_normal) or already [P,12]."""
        n12 = shading_normal
        return radiance_loss(cam_center, direct_light, xyz, geo_normal, self.incident_dirs, self.incident_areas,
                             self.visibility_tracing, self.hemi_index_buffers, self.uv_buffers, self.radiances.detach(),
                             self.radiance_ratio.detach(), n12, albedo, roughness, **kw)
