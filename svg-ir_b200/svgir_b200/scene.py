"""Seeded synthetic surfel clouds, cameras and materials of the shapes BASELINE.json names.

Everything is generated with numpy on the CPU (so the CPU oracle and the GPU path see identical
bits) following SURVEY.md section 8(d). Camera conventions restate the reference's
`scene/cameras.py:63-80` and `utils/graphics_utils.py:127-168`: matrices are stored in the
row-vector convention (`world_view_transform = W2C.T`, `full_proj = view @ proj`), i.e. the
kernels read them column-major exactly like `cuda_rasterizer/auxiliary.h:65-84`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

FOVX_TENSOIR = 0.6911112  # relighting.py:150 in the reference


@dataclass
class Camera:
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray  # [4,4] float32, row-vector convention
    projmatrix: np.ndarray  # [4,4] float32, view @ proj
    campos: np.ndarray  # [3]
    patch_bbox: np.ndarray  # [4] = (h0, w0, h1, w1)
    prcppoint: np.ndarray  # [2]


@dataclass
class SurfelCloud:
    means3D: np.ndarray  # [P,3]
    scales: np.ndarray  # [P,3]
    rotations: np.ndarray  # [P,4] (r,x,y,z), normalised
    opacity: np.ndarray  # [P,1]
    shs: np.ndarray  # [P,16,3]
    normals: np.ndarray  # [P,3] geometric normal (3rd column of R)
    extras: dict = field(default_factory=dict)

    @property
    def P(self) -> int:
        return self.means3D.shape[0]


def _fibonacci_sphere(n: int) -> np.ndarray:
    i = np.arange(n, dtype=np.float64) + 0.5
    z = 1.0 - 2.0 * i / n
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    th = np.pi * (3.0 - math.sqrt(5.0)) * i
    return np.stack([r * np.cos(th), r * np.sin(th), z], -1)


def _rotmat_to_quat(R: np.ndarray) -> np.ndarray:
    """Batched rotation matrix [N,3,3] -> quaternion (r,x,y,z) with r >= 0."""
    m00, m11, m22 = R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]
    q = np.empty((R.shape[0], 4), np.float64)
    q[:, 0] = np.sqrt(np.maximum(0, 1 + m00 + m11 + m22)) / 2
    q[:, 1] = np.sqrt(np.maximum(0, 1 + m00 - m11 - m22)) / 2
    q[:, 2] = np.sqrt(np.maximum(0, 1 - m00 + m11 - m22)) / 2
    q[:, 3] = np.sqrt(np.maximum(0, 1 - m00 - m11 + m22)) / 2
    q[:, 1] = np.copysign(q[:, 1], R[:, 2, 1] - R[:, 1, 2])
    q[:, 2] = np.copysign(q[:, 2], R[:, 0, 2] - R[:, 2, 0])
    q[:, 3] = np.copysign(q[:, 3], R[:, 1, 0] - R[:, 0, 1])
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def make_surfels(P: int, seed: int = 1234, scale_mult: float = 1.6) -> SurfelCloud:
    """P surfels on a bumpy unit sphere, facing outwards (SURVEY 8(d) 'Geometry')."""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((P, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = 1.0 + 0.05 * rng.standard_normal((P, 1))
    means = d * r
    n = d + 0.2 * rng.standard_normal((P, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    # tangent frame with a random in-plane angle
    helper = np.where(np.abs(n[:, 2:3]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
    t0 = np.cross(helper, n)
    t0 /= np.linalg.norm(t0, axis=1, keepdims=True)
    t1 = np.cross(n, t0)
    ang = rng.uniform(0, 2 * np.pi, (P, 1))
    a = np.cos(ang) * t0 + np.sin(ang) * t1
    b = np.cross(n, a)
    R = np.stack([a, b, n], axis=2)  # columns a, b, n
    quat = _rotmat_to_quat(R)
    s0 = scale_mult * math.sqrt(4 * math.pi / P)
    sxy = np.exp(rng.normal(math.log(s0), 0.35, (P, 2)))
    scales = np.concatenate([sxy, np.full((P, 1), 1e-6)], 1)
    opacity = 1.0 / (1.0 + np.exp(-rng.normal(1.5, 1.5, (P, 1))))
    shs = np.concatenate([rng.uniform(-1, 1, (P, 1, 3)), 0.05 * rng.standard_normal((P, 15, 3))], 1)
    f32 = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    return SurfelCloud(f32(means), f32(scales), f32(quat), f32(opacity), f32(shs), f32(n))


def look_at_camera(W: int, H: int, view_index: int = 0, n_views: int = 8, distance: float = 4.0,
                   fovx: float = FOVX_TENSOIR) -> Camera:
    """Camera on a Fibonacci sphere of view positions looking at the origin."""
    eye = _fibonacci_sphere(max(n_views, 1))[view_index % max(n_views, 1)] * distance
    fwd = -eye / np.linalg.norm(eye)  # camera +z looks at the origin
    up = np.array([0.0, 0.0, 1.0]) if abs(fwd[2]) < 0.95 else np.array([0.0, 1.0, 0.0])
    right = np.cross(up, fwd)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, eye
    w2c = np.linalg.inv(c2w)
    view = np.float32(w2c).T.copy()  # world_view_transform
    fovy = 2 * math.atan(math.tan(fovx / 2) * H / W)
    znear, zfar = 0.01, 100.0
    tx, ty = math.tan(fovx / 2), math.tan(fovy / 2)
    Pm = np.zeros((4, 4), np.float32)
    Pm[0, 0] = 2.0 * znear / (2 * tx * znear)
    Pm[1, 1] = 2.0 * znear / (2 * ty * znear)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    proj = Pm.T.copy()
    full = (view @ proj).astype(np.float32)
    campos = np.linalg.inv(view.astype(np.float64))[3, :3].astype(np.float32)
    return Camera(W, H, tx, ty, view, full, campos,
                  np.array([0, 0, H, W], np.float32), np.array([0.5, 0.5], np.float32))


def fibonacci_hemisphere_dirs(normals: np.ndarray, sample_num: int) -> tuple[np.ndarray, np.ndarray]:
    """Restates fibonacci_sphere_sampling(random_rotate=False) + rotation_between_z
    (reference utils/graphics_utils.py:9-37, utils/sh_utils.py:36-68) in numpy float32."""
    n = normals.astype(np.float32)
    delta = np.float32(np.pi * (3.0 - np.sqrt(5.0)))
    idx = np.arange(sample_num, dtype=np.float32)[None]
    z = np.maximum(1 - 2 * idx / (2 * sample_num - 1), np.float32(np.sin(10 / 180 * np.pi))).astype(np.float32)
    rad = np.sqrt(1 - z ** 2).astype(np.float32)
    theta = (delta * idx).astype(np.float32)
    y = np.cos(theta) * rad
    x = np.sin(theta) * rad
    zs = np.stack([x, y, z], axis=-2).astype(np.float32)  # [1,3,S]
    v1, v2 = -n[:, 1], n[:, 0]
    cp1 = np.maximum(n[:, 2] + 1, np.float32(1e-7))
    R = np.zeros((n.shape[0], 3, 3), np.float32)
    R[:, 0, 0] = 1 + (-v2 * v2) / cp1
    R[:, 0, 1] = (v1 * v2) / cp1
    R[:, 0, 2] = v2
    R[:, 1, 0] = (v1 * v2) / cp1
    R[:, 1, 1] = 1 + (-v1 * v1) / cp1
    R[:, 1, 2] = -v1
    R[:, 2, 0] = -v2
    R[:, 2, 1] = v1
    R[:, 2, 2] = 1 + (-v2 * v2 - v1 * v1) / cp1
    flip = (n[:, 2] + 1 > 0)[:, None, None]
    R = np.where(flip, R, -np.eye(3, dtype=np.float32)[None])
    dirs = R @ zs  # [N,3,S]
    dirs = dirs / np.maximum(np.linalg.norm(dirs, axis=-2, keepdims=True), 1e-12)
    dirs = np.ascontiguousarray(np.transpose(dirs, (0, 2, 1)), dtype=np.float32)
    areas = np.full((n.shape[0], sample_num, 1), 2 * np.pi, np.float32)
    return dirs, areas


def make_materials(cloud: SurfelCloud, sample_num: int, seed: int = 4321, env_hw=(32, 64)) -> dict:
    """Stage-2 spatially-varying materials and lighting inputs (SURVEY 8(d) 'Materials')."""
    rng = np.random.default_rng(seed)
    P = cloud.P
    f32 = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    sn = cloud.normals[:, None, :] + 0.1 * rng.standard_normal((P, 4, 3))
    sn /= np.linalg.norm(sn, axis=-1, keepdims=True)
    dirs, areas = fibonacci_hemisphere_dirs(cloud.normals, sample_num)
    return dict(
        base_color=f32(0.03 + 0.77 * rng.uniform(0, 1, (P, 12))),
        roughness=f32(0.09 + 0.9 * rng.uniform(0, 1, (P, 4))),
        shading_normals=f32(sn),
        incident_dirs=dirs,
        incident_areas=areas,
        visibility=f32(rng.uniform(0, 1, (P, sample_num, 1)) > 0.3),
        radiance=f32(0.2 * rng.uniform(0, 1, (P, sample_num, 3))),
        env_param=f32(3.0 * rng.uniform(0, 1, (1, env_hw[0], env_hw[1], 3))),
    )


def make_materials_torch(cloud: SurfelCloud, sample_num: int, seed: int, device, env_hw=(32, 64)) -> dict:
    """`make_materials` with the big per-sample buffers ([P,Ns,*]: 32 B per sample) generated directly on `device`
    by torch -- same distributions and the same incident-direction formula, different random bits. For workloads
    where no CPU oracle looks at the inputs (the 1M-surfel / Ns=384 bench configurations), where building
    2-4 GB with numpy and copying it would dominate the run."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    P = cloud.P
    f32 = dict(dtype=torch.float32, device=device)
    n = torch.from_numpy(cloud.normals).to(device)
    sn = n[:, None, :] + 0.1 * torch.randn((P, 4, 3), generator=g, **f32)
    sn = sn / sn.norm(dim=-1, keepdim=True)
    # fibonacci_sphere_sampling(random_rotate=False) + rotation_between_z, as fibonacci_hemisphere_dirs above
    delta = math.pi * (3.0 - math.sqrt(5.0))
    idx = torch.arange(sample_num, **f32)[None]
    z = torch.clamp_min(1 - 2 * idx / (2 * sample_num - 1), math.sin(10 / 180 * math.pi))
    rad = torch.sqrt(1 - z * z)
    theta = delta * idx
    zs = torch.stack([torch.sin(theta) * rad, torch.cos(theta) * rad, z], dim=-2)  # [1,3,S]
    v1, v2 = -n[:, 1], n[:, 0]
    cp1 = torch.clamp_min(n[:, 2] + 1, 1e-7)
    R = torch.zeros((P, 3, 3), **f32)
    R[:, 0, 0] = 1 + (-v2 * v2) / cp1
    R[:, 0, 1] = (v1 * v2) / cp1
    R[:, 0, 2] = v2
    R[:, 1, 0] = (v1 * v2) / cp1
    R[:, 1, 1] = 1 + (-v1 * v1) / cp1
    R[:, 1, 2] = -v1
    R[:, 2, 0] = -v2
    R[:, 2, 1] = v1
    R[:, 2, 2] = 1 + (-v2 * v2 - v1 * v1) / cp1
    flip = (n[:, 2] + 1 > 0)[:, None, None]
    R = torch.where(flip, R, -torch.eye(3, **f32)[None])
    dirs = R @ zs
    dirs = dirs / dirs.norm(dim=-2, keepdim=True).clamp_min(1e-12)
    dirs = dirs.transpose(1, 2).contiguous()
    return dict(
        base_color=0.03 + 0.77 * torch.rand((P, 12), generator=g, **f32),
        roughness=0.09 + 0.9 * torch.rand((P, 4), generator=g, **f32),
        shading_normals=sn.contiguous(),
        incident_dirs=dirs,
        incident_areas=torch.full((P, sample_num, 1), 2 * math.pi, **f32),
        visibility=(torch.rand((P, sample_num, 1), generator=g, **f32) > 0.3).float(),
        radiance=0.2 * torch.rand((P, sample_num, 3), generator=g, **f32),
        env_param=3.0 * torch.rand((1, env_hw[0], env_hw[1], 3), generator=g, **f32),
    )
