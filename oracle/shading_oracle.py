"""TEST INFRASTRUCTURE ONLY -- torch (CPU-runnable) restatement of the reference's live shading path.

Restates, function for function:
  direct_light_learnable  scene/direct_light_map.py:70-83,103-106  (softplus env, lat-long bilinear
                          grid_sample with align_corners=True + zero padding, result x 2.0)
  direct_light_hdr        scene/envmap.py:54-72 (F.interpolate to 32x64, optional 3x3 transform)
  ggx_specular4           gaussian_renderer/svgss.py:595-631
  rendering_equation4     gaussian_renderer/svgss.py:537-593
  pack_features           gaussian_renderer/svgss.py:141-166 (features / vfeatures packing)

Parity pin: `tests/golden/make_golden_shading.py` imports the reference's own
gaussian_renderer/svgss.py (with stub modules for its missing third-party imports) in the build
container, runs it on seeded inputs and stores inputs+outputs+gradients in
tests/golden/ref_shading_*.npz; tests/test_oracle_cpu.py checks this restatement against those
fixtures. The CUDA kernels are then checked against this restatement (and the fixtures).
The bilinear lookup is written out explicitly instead of calling F.grid_sample so that it is a
restatement of the arithmetic; it is compared with torch's grid_sample in the tests.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def grid_sample_bilinear_zeros(env_chw: torch.Tensor, gx: torch.Tensor, gy: torch.Tensor) -> torch.Tensor:
    """env_chw [3,He,We]; gx,gy [N] in [-1,1], align_corners=True, zero padding. Returns [N,3]."""
    C, He, We = env_chw.shape
    ix = (gx + 1) / 2 * (We - 1)
    iy = (gy + 1) / 2 * (He - 1)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    wx1, wy1 = ix - x0, iy - y0
    wx0, wy0 = 1 - wx1, 1 - wy1

    def tap(xi, yi):
        ok = (xi >= 0) & (xi <= We - 1) & (yi >= 0) & (yi <= He - 1)
        xc = xi.clamp(0, We - 1).long()
        yc = yi.clamp(0, He - 1).long()
        v = env_chw[:, yc, xc].t()  # [N,3]
        return v * ok[:, None].to(v.dtype)

    return (tap(x0, y0) * (wx0 * wy0)[:, None] + tap(x1, y0) * (wx1 * wy0)[:, None] +
            tap(x0, y1) * (wx0 * wy1)[:, None] + tap(x1, y1) * (wx1 * wy1)[:, None])


def _latlong_grid(dirs: torch.Tensor):
    phi = torch.arccos(dirs[:, 2]).reshape(-1) - 1e-6
    theta = torch.atan2(dirs[:, 1], dirs[:, 0]).reshape(-1)
    query_y = (phi / math.pi) * 2 - 1
    query_x = -theta / math.pi
    return query_x, query_y


def direct_light_learnable(env_param: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """env_param [1,He,We,3] raw parameter; dirs [...,3]. direct_light_map.py:70-83."""
    shape = dirs.shape
    d = dirs.reshape(-1, 3)
    env = F.softplus(env_param)[0].permute(2, 0, 1)  # get_env, [3,He,We]
    qx, qy = _latlong_grid(d)
    return grid_sample_bilinear_zeros(env, qx, qy).reshape(*shape) * 2.0


def direct_light_hdr(envmap_hw3: torch.Tensor, dirs: torch.Tensor, transform=None) -> torch.Tensor:
    """envmap [He,We,3] linear HDR; scene/envmap.py:54-72."""
    shape = dirs.shape
    d = dirs.reshape(-1, 3)
    if transform is not None:
        d = d @ transform.T
    env = envmap_hw3.permute(2, 0, 1).unsqueeze(0)
    env = F.interpolate(env, size=(32, 64), mode="bilinear", align_corners=False)[0]
    qx, qy = _latlong_grid(d)
    return grid_sample_bilinear_zeros(env, qx, qy).reshape(*shape)


def ggx_specular4(normal, pts2c, pts2l, roughness, fresnel):
    """svgss.py:595-631. normal [n,4,3], pts2c [n,3], pts2l [n,Ns,3], roughness [n,4] -> [n,Ns,4,1]."""
    L = F.normalize(pts2l, dim=-1).unsqueeze(-2)
    V = F.normalize(pts2c, dim=-1).unsqueeze(-2)
    Hh = F.normalize((L + V[:, None, :]) / 2.0, dim=-1)
    N = F.normalize(normal, dim=-1)
    NoV = torch.sum(V * N, dim=-1, keepdim=True)
    N = N * NoV.sign()
    NoL = torch.sum(N[:, None, :] * L, dim=-1, keepdim=True).clamp(1e-6, 1)
    NoV = torch.sum(N * V, dim=-1, keepdim=True).clamp(1e-6, 1)
    NoH = torch.sum(N[:, None, :] * Hh, dim=-1, keepdim=True).clamp(1e-6, 1)
    VoH = torch.sum(V[:, None, :] * Hh, dim=-1, keepdim=True).clamp(1e-6, 1)
    rough = roughness.unsqueeze(1).unsqueeze(-1)
    alpha = rough * rough
    alpha2 = alpha * alpha
    k = (alpha + 2 * rough + 1.0) / 8.0
    FMi = ((-5.55473) * VoH - 6.98316) * VoH
    frac0 = fresnel + (1 - fresnel) * torch.pow(2.0, FMi)
    frac = frac0 * alpha2
    nom0 = NoH * NoH * (alpha2 - 1) + 1
    nom1 = NoV.unsqueeze(1) * (1 - k) + k
    nom2 = NoL * (1 - k) + k
    nom = (4 * math.pi * nom0 * nom0 * nom1 * nom2).clamp(1e-6, 4 * math.pi)
    return frac / nom


def rendering_equation4(base_color, roughness, normals, viewdirs, radiance, global_light_fn,
                        visibility, incident_dirs, incident_areas):
    """svgss.py:537-593. global_light_fn(dirs[N,Ns,3]) -> [N,Ns,3] (before the [0,64] clamp)."""
    global_incident_lights = global_light_fn(incident_dirs).clamp(0, 64)
    local_incident_lights = radiance
    incident_visibility = visibility
    global_incident_lights = global_incident_lights * incident_visibility
    incident_lights = local_incident_lights + global_incident_lights
    n_d_i = (normals[:, None] * incident_dirs[:, :, None]).sum(-1, keepdim=True).clamp(min=0)
    f_d = base_color[:, None] / math.pi
    f_s = ggx_specular4(normals, viewdirs, incident_dirs, roughness, fresnel=0.04).squeeze(-1).repeat(1, 1, 3)

    def cm(t):  # [n,Ns,4,3] -> channel-major [n,Ns,12]
        return t.transpose(2, 3).reshape(t.shape[0], t.shape[1], -1)

    transport = cm(incident_lights[:, :, None] * incident_areas[:, :, None] * n_d_i)
    specular = (f_s * transport).mean(dim=-2)
    pbr = ((f_d + f_s) * transport).mean(dim=-2)
    diffuse_light = transport.mean(dim=-2)
    direct_pbr = ((f_d + f_s) * cm(global_incident_lights[:, :, None] * incident_areas[:, :, None] * n_d_i)).mean(dim=-2)
    indirect_pbr = ((f_d + f_s) * cm(local_incident_lights[:, :, None] * incident_areas[:, :, None] * n_d_i)).mean(dim=-2)
    extra = {
        "incident_dirs": incident_dirs,
        "incident_lights": incident_lights,
        "local_incident_lights": local_incident_lights,
        "global_incident_lights": global_incident_lights,
        "incident_visibility": incident_visibility,
        "diffuse_light": diffuse_light,
        "specular": specular,
        "direct": direct_pbr,
        "indirect": indirect_pbr,
    }
    return pbr, extra


def pack_features(pbr, extra, base_color, roughness, normals, view3x3, is_training):
    """svgss.py:141-166: features / vfeatures exactly as render_view packs them."""
    if is_training:
        features = torch.cat([extra["incident_visibility"].mean(-2), extra["local_incident_lights"].mean(-2)], dim=-1)
    else:
        features = torch.cat([extra["incident_lights"].mean(-2), extra["local_incident_lights"].mean(-2),
                              extra["incident_visibility"].mean(-2)], dim=-1)
    n = normals @ view3x3
    n = n.transpose(1, 2).reshape(n.shape[0], -1)
    if is_training:
        vfeatures = torch.cat([pbr, base_color, n, roughness, extra["diffuse_light"]], dim=-1)
    else:
        vfeatures = torch.cat([pbr, base_color, n, roughness, extra["direct"], extra["indirect"]], dim=-1)
    return features, vfeatures
