"""TEST INFRASTRUCTURE ONLY -- torch (CPU-capable) restatement of the reference's SH render_equation
operators (rgss-rasterization/render_equation.cu) and of its incident-ray sampler
(utils/graphics_utils.py:9-37, utils/sh_utils.py:36-68). Vectorised over [P,Ns]; every function cites
the reference lines it follows. Pinned against outputs of the reference kernels themselves
(tests/golden/ref_req_sh_*.npz, generated on a B200 by tests/golden/make_golden_req_gpu.py) and, for the
sampler, against the reference's own torch code run in this container (make_golden_sampling.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

PI_R = 3.14159  # the reference's literal
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def fibonacci_sphere_sampling(normals, sample_num, rand_u=None):
    """utils/graphics_utils.py:9-37 with rotation_between_z (utils/sh_utils.py:36-68) inlined; the
    reference hard-codes device='cuda', this runs on the device of `normals`. rand_u [...,1] or None."""
    pre_shape = normals.shape[:-1]
    n = normals.reshape(-1, 3).float()
    dev = n.device
    delta = np.pi * (3.0 - np.sqrt(5.0))
    idx = torch.arange(sample_num, dtype=torch.float, device=dev)[None]
    z = (1 - 2 * idx / (2 * sample_num - 1)).clamp_min(np.sin(10 / 180 * np.pi))
    rad = torch.sqrt(1 - z ** 2)
    theta = delta * idx
    if rand_u is not None:
        theta = rand_u.reshape(-1, 1).float() * 2 * np.pi + theta
    y = torch.cos(theta) * rad
    x = torch.sin(theta) * rad
    zs = torch.stack([x, y, z.expand_as(y)], dim=-2)
    v1, v2 = -n[:, 1], n[:, 0]
    c = (n[:, 2] + 1).clamp_min(1e-7)
    R = torch.zeros(n.shape[0], 3, 3, dtype=torch.float32, device=dev)
    R[:, 0, 0] = 1 + (-v2 * v2) / c
    R[:, 0, 1] = (v1 * v2) / c
    R[:, 0, 2] = v2
    R[:, 1, 0] = (v1 * v2) / c
    R[:, 1, 1] = 1 + (-v1 * v1) / c
    R[:, 1, 2] = -v1
    R[:, 2, 0] = -v2
    R[:, 2, 1] = v1
    R[:, 2, 2] = 1 + (-v2 * v2 - v1 * v1) / c
    R = torch.where((n[:, 2] + 1 > 0)[:, None, None], R, -torch.eye(3, dtype=torch.float32, device=dev).expand_as(R))
    dirs = torch.nn.functional.normalize(R @ zs, dim=-2).transpose(-1, -2)
    areas = torch.ones_like(dirs)[..., 0:1] * 2 * np.pi
    return dirs.reshape(*pre_shape, sample_num, 3), areas.reshape(*pre_shape, sample_num, 1)


def sh_basis(d):
    """computeSHcoef (render_equation.cu:19-53). d [...,3] -> [...,16]."""
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    c = [torch.full_like(x, SH_C0), -SH_C1 * y, SH_C1 * z, -SH_C1 * x,
         SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2.0 * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy),
         SH_C3[0] * y * (3.0 * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (4.0 * zz - xx - yy),
         SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy), SH_C3[4] * x * (4.0 * zz - xx - yy),
         SH_C3[5] * z * (xx - yy), SH_C3[6] * x * (xx - 3.0 * yy)]
    return torch.stack(c, dim=-1)


def kernel_dirs(normals, sample_num, rand_float=None):
    """In-kernel direction generation (render_equation.cu:90-117, 590-618). rand_float [P,Ns,1] or None."""
    P = normals.shape[0]
    ray = torch.arange(sample_num, dtype=torch.float32, device=normals.device)[None]
    delta = torch.tensor(PI_R, dtype=torch.float32) * (3.0 - torch.sqrt(torch.tensor(5.0)))
    z = 1 - 2 * ray / (2 * float(sample_num) - 1)
    rad = torch.sqrt(1 - z * z)
    theta = delta * ray
    if rand_float is not None:
        theta = rand_float.reshape(P, sample_num) * 2 * PI_R + theta
    y, x = torch.cos(theta) * rad, torch.sin(theta) * rad
    z = z.expand_as(y)
    n = normals
    v1, v2 = -n[:, 1:2], n[:, 0:1]
    c = (n[:, 2:3] + 1).clamp_min(1e-7)
    v12 = v1 * v2
    ox = (1 + (-v2 * v2) / c) * x + (v12 / c) * y + v2 * z
    oy = (v12 / c) * x + (1 + (-v1 * v1) / c) * y + (-v1) * z
    oz = (-v2) * x + v1 * y + (1 + (-v2 * v2 - v1 * v1) / c) * z
    nn = torch.sqrt((ox * ox + oy * oy + oz * oz).clamp_min(1e-7))
    return torch.stack([ox / nn, oy / nn, oz / nn], dim=-1)  # [P,Ns,3]


def _terms(base, rough, metal, n, v, inc, dsh, vsh, dirs):
    """Per-(surfel,sample) quantities shared by forward and backward (render_equation.cu:118-168)."""
    coef = sh_basis(dirs)                                   # [P,Ns,16]
    Si, Sd, Sv = inc.shape[1], dsh.shape[1], vsh.shape[1]
    lraw = torch.einsum("psk,pkc->psc", coef[..., :Si], inc)
    graw = 0.5 + torch.einsum("psk,kc->psc", coef[..., :Sd], dsh[0])
    vraw = 0.5 + torch.einsum("psk,pk->ps", coef[..., :Sv], vsh[..., 0])
    local, glob, vis = lraw.clamp_min(0), graw.clamp_min(0), vraw.clamp(0, 1)
    li = vis[..., None] * glob + local
    N, V = n[:, None, :], v[:, None, :]
    h = dirs + V
    hnorm = h.norm(dim=-1, keepdim=True).clamp_min(1e-7)
    hn = h / hnorm
    hdn = (hn * N).sum(-1).clamp_min(0)
    hdo = (hn * V).sum(-1).clamp_min(0)
    ndi = (N * dirs).sum(-1).clamp_min(0)
    ndo = (N * V).sum(-1).clamp_min(0).expand_as(ndi)
    r = rough.reshape(-1, 1)
    m = metal.reshape(-1, 1)
    r2 = (r * r).clamp_min(1e-7)
    amp = 1.0 / (r2 * PI_R)
    sharp = 2.0 / r2
    e = torch.exp(sharp * (hdn - 1.0))
    D = amp * e
    F0 = 0.04 * (1.0 - m)[..., None] + base[:, None, :] * m[..., None]
    pw5 = torch.pow(1.0 - hdo, 5.0)
    F = F0 + (1.0 - F0) * pw5[..., None]
    r2v = torch.pow(1.0 + r, 2.0) / 8.0
    den1 = (ndi * (1 - r2v) + r2v).clamp_min(1e-7)
    den2 = (ndo * (1 - r2v) + r2v).clamp_min(1e-7)
    g1, g2 = 0.5 / den1, 0.5 / den2
    Vt = g1 * g2
    f_d = ((1 - m) * base / PI_R)[:, None, :]
    f_s = (D * Vt)[..., None] * F
    return dict(coef=coef, lraw=lraw, graw=graw, vraw=vraw, local=local, glob=glob, vis=vis, li=li, hnorm=hnorm, hn=hn,
                hdn=hdn, hdo=hdo, ndi=ndi, ndo=ndo, r2=r2, amp=amp, sharp=sharp, e=e, D=D, F0=F0, pw5=pw5, F=F, r2v=r2v,
                den1=den1, den2=den2, g1=g1, g2=g2, V=Vt, f_d=f_d, f_s=f_s, r=r, m=m)


def forward_complex(base, rough, metal, n, v, inc, dsh, vsh, sample_num, rand_float=None, dirs=None):
    """render_equation_forward_complex_kernel (:55-190); with rand_float it is the simple kernel's
    training variant (:555-655). Returns a dict with every output of the _complex operator."""
    if dirs is None:
        dirs = kernel_dirs(n, sample_num, rand_float)
    t = _terms(base, rough, metal, n, v, inc, dsh, vsh, dirs)
    tmp = (2.0 * PI_R * t["ndi"] / float(sample_num))[..., None]
    transport = t["li"] * tmp
    diffuse_light = transport.sum(1)
    local_diffuse = (t["local"] * tmp).sum(1)
    rgb_d = (t["f_d"] * transport).sum(1)
    rgb_s = (t["f_s"] * transport).sum(1)
    acc = diffuse_light / PI_R + rgb_s
    return dict(pbr=rgb_d + rgb_s, incident_dirs=dirs, incident_lights=t["li"], local_incident_lights=t["local"],
                global_incident_lights=t["vis"][..., None] * t["glob"], incident_visibility=t["vis"][..., None],
                diffuse_light=diffuse_light, local_diffuse_light=local_diffuse, accum=acc.sum(-1, keepdim=True) / 3,
                rgb_d=rgb_d, rgb_s=rgb_s)


def backward_legacy(base, rough, metal, n, v, inc, dsh, vsh, sample_num, dirs, g_pbr, g_dl):
    """render_equation_backward_kernel (:280-470), slips included: dL_dn_d_i overwritten (:406), incident-SH
    gradient looped over S_direct (:453), no clamp masks (:438-452), half-vector normalisation Jacobian
    ignored (:432-433). dL_ddirect_shs is the exact sum (the reference accumulates it with a data race)."""
    t = _terms(base, rough, metal, n, v, inc, dsh, vsh, dirs)
    Ns = float(sample_num)
    tt = (2.0 * PI_R * t["ndi"] / Ns)[..., None]
    gP, gD = g_pbr[:, None, :], g_dl[:, None, :]
    fds = t["f_d"] + t["f_s"]
    g_f = gP * t["li"] * tt
    g_li = gP * fds * tt + gD * tt
    g_base = g_f * (1 - t["m"])[..., None] / PI_R
    g_metal = -(g_f * base[:, None, :]).sum(-1) / PI_R
    g_D = (g_f * t["V"][..., None] * t["F"]).sum(-1)
    g_F = g_f * (t["D"] * t["V"])[..., None]
    g_V = (g_f * t["D"][..., None] * t["F"]).sum(-1)
    g_amp, g_e = g_D * t["e"], g_D * t["amp"]
    g_sharp = (t["hdn"] - 1.0) * t["e"] * g_e
    g_hdn = t["sharp"] * t["e"] * g_e
    g_r2 = -2.0 / (t["r2"] * t["r2"]) * g_sharp - 1.0 / (t["r2"] * t["r2"] * PI_R) * g_amp
    g_rough = g_r2 * 2.0 * t["r"]
    g_F0 = (1.0 - t["pw5"])[..., None] * g_F
    g_hdo = ((1.0 - t["F0"]) * g_F).sum(-1) * -5.0 * torch.pow(1.0 - t["hdo"], 4.0)
    g_base = g_base + t["m"][..., None] * g_F0
    g_metal = g_metal + ((base[:, None, :] - 0.04) * g_F0).sum(-1)
    g_g1, g_g2 = g_V * t["g2"], g_V * t["g1"]
    g_den1 = -0.5 / (t["den1"] * t["den1"]) * g_g1
    g_den2 = -0.5 / (t["den2"] * t["den2"]) * g_g2
    g_ndi = g_den1 * (1 - t["r2v"])  # overwrite
    g_ndo = g_den2 * (1 - t["r2v"])
    g_r2v = (1.0 - t["ndi"]) * g_den1 + (1.0 - t["ndo"]) * g_den2
    g_rough = g_rough + (1.0 + t["r"]) / 4.0 * g_r2v
    N, V = n[:, None, :], v[:, None, :]
    m_hdn, m_hdo = (t["hdn"] > 0)[..., None], (t["hdo"] > 0)[..., None]
    m_ndi, m_ndo = (t["ndi"] > 0)[..., None], (t["ndo"] > 0)[..., None]
    g_h = m_hdn * N * g_hdn[..., None] + m_hdo * V * g_hdo[..., None]
    g_n = m_hdn * t["hn"] * g_hdn[..., None] + m_ndi * dirs * g_ndi[..., None] + m_ndo * V * g_ndo[..., None]
    g_v = m_hdo * t["hn"] * g_hdo[..., None] + m_ndo * N * g_ndo[..., None] + g_h / t["hnorm"]
    g_vis = (g_li * t["glob"]).sum(-1)
    g_glob = g_li * t["vis"][..., None]
    coef = t["coef"]
    Si, Sd, Sv = inc.shape[1], dsh.shape[1], vsh.shape[1]
    d_vsh = torch.einsum("ps,psk->pk", g_vis, coef[..., :Sv])[..., None]
    d_dsh = torch.einsum("psc,psk->kc", g_glob.double(), coef[..., :Sd].double()).float()[None]
    d_inc = torch.zeros_like(inc)
    k = min(Si, Sd)
    d_inc[:, :k] = torch.einsum("psc,psk->pkc", g_li, coef[..., :k])
    return dict(dL_dbase_color=g_base.sum(1), dL_droughness=g_rough.sum(1, keepdim=True),
                dL_dmetallic=g_metal.sum(1, keepdim=True), dL_dnormals=g_n.sum(1), dL_dviewdirs=g_v.sum(1),
                dL_dincidents_shs=d_inc, dL_ddirect_shs=d_dsh, dL_dvisibility_shs=d_vsh)


def backward_analytic(base, rough, metal, n, v, inc, dsh, vsh, sample_num, dirs, g_pbr, g_dl):
    """Autograd gradient of forward_complex w.r.t. the eight inputs (incident_dirs held fixed), float64."""
    ins = [x.detach().double().requires_grad_(True) for x in (base, rough, metal, n, v, inc, dsh, vsh)]
    o = forward_complex(*ins, sample_num, dirs=dirs.double())
    loss = (o["pbr"] * g_pbr.double()).sum() + (o["diffuse_light"] * g_dl.double()).sum()
    gs = torch.autograd.grad(loss, ins)
    names = ("dL_dbase_color", "dL_droughness", "dL_dmetallic", "dL_dnormals", "dL_dviewdirs", "dL_dincidents_shs",
             "dL_ddirect_shs", "dL_dvisibility_shs")
    return {k: g.float() for k, g in zip(names, gs)}


def make_inputs(P, Si=16, Sd=16, Sv=16, seed=0, device="cpu"):
    """Seeded random inputs of the operators' shapes (render_equation.h:7-46)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    nrm = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1)
    view = torch.nn.functional.normalize(nrm + 0.8 * torch.randn(P, 3, generator=g), dim=-1)
    d = dict(base_color=0.05 + 0.9 * r(P, 3), roughness=0.1 + 0.85 * r(P, 1), metallic=r(P, 1), normals=nrm, viewdirs=view,
             incidents_shs=0.3 * torch.randn(P, Si, 3, generator=g), direct_shs=0.4 * torch.randn(1, Sd, 3, generator=g),
             visibility_shs=0.4 * torch.randn(P, Sv, 1, generator=g))
    return {k: t.float().to(device).contiguous() for k, t in d.items()}
