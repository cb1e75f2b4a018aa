"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of oracle/bvh_oracle.c (the CPU restatement of
submodules/bvh: RayTracer.__init__ leaf boxes, construct_bvh, trace_bvh_opacity)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .svgss import lib, _p, _f32


def init(means3D, scales, rotations):
    """submodules/bvh/__init__.py:29-57 -> (nodes [2P-1,5] int32, aabbs [2P-1,6] f32)."""
    means3D, scales, rotations = map(_f32, (means3D, scales, rotations))
    P = means3D.shape[0]
    nodes = np.empty((2 * P - 1, 5), np.int32)
    aabbs = np.empty((2 * P - 1, 6), np.float32)
    lib().oracle_bvh_init(C.c_int(P), _p(means3D), _p(scales), _p(rotations), _p(nodes), _p(aabbs))
    return nodes, aabbs


def build(nodes, aabbs):
    """construct_bvh (construct.cu:147-265) in place -> (nodes, aabbs, morton [P] uint64)."""
    P = (nodes.shape[0] + 1) // 2
    nodes = np.ascontiguousarray(nodes, np.int32).copy()
    aabbs = np.ascontiguousarray(aabbs, np.float32).copy()
    morton = np.zeros((P,), np.uint64)
    lib().oracle_bvh_build(C.c_int(P), _p(nodes), _p(aabbs), _p(morton))
    return nodes, aabbs, morton


def create(means3D, scales, rotations):
    return build(*init(means3D, scales, rotations))


def trace_opacity(nodes, aabbs, rays_o, rays_d, means3D, cov_inv, opacity, normals, brute=False):
    """trace_bvh_opacity (bvh.cu:89-116 + trace.cu:196-286) -> (contributes int32, visibility f32),
    both shaped like rays_o without its last axis."""
    rays_o, rays_d, means3D, cov_inv, opacity, normals = map(_f32, (rays_o, rays_d, means3D, cov_inv, opacity, normals))
    shape = rays_o.shape[:-1]
    n = int(np.prod(shape)) if shape else 1
    P = means3D.shape[0]
    contrib = np.zeros(shape, np.int32)
    vis = np.ones(shape, np.float32)
    fn = lib().oracle_bvh_trace_brute if brute else lib().oracle_bvh_trace_opacity
    fn(C.c_int(n), C.c_int(P), _p(np.ascontiguousarray(nodes, np.int32)), _p(_f32(aabbs)), _p(rays_o), _p(rays_d),
       _p(means3D), _p(cov_inv), _p(opacity.reshape(-1)), _p(normals), _p(contrib), _p(vis))
    return contrib, vis


def inverse_covariance(scales, rotations):
    """gaussian_model.py:379-382 with build_scaling_rotation / strip_symmetric
    (utils/general_utils.py:66-79,151-160): L = R diag(1/s); Sigma^-1 = L L^T, upper triangle."""
    s = 1.0 / np.asarray(scales, np.float32)
    q = np.asarray(rotations, np.float32)
    q = q / np.sqrt((q * q).sum(-1, keepdims=True, dtype=np.float32))
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                  2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                  2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    L = R * s[:, None, :]
    M = L @ L.transpose(0, 2, 1)
    return np.stack([M[:, 0, 0], M[:, 0, 1], M[:, 0, 2], M[:, 1, 1], M[:, 1, 2], M[:, 2, 2]], -1).astype(np.float32)
