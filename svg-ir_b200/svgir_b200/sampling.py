"""Incident-ray sampling on the GPU (csrc/sampling.cu).

Drop-in for the reference's torch helpers, same names / arguments / return values:
  fibonacci_sphere_sampling(normals, sample_num, random_rotate=True)   utils/graphics_utils.py:9-37
  sample_incident_rays(normals, is_training=False, sample_num=24)      scene/gaussian_model.py:23-31
The per-surfel random azimuth offset comes from `torch.rand(*pre_shape, 1, device=normals.device)`
exactly as in the reference (graphics_utils.py:21), so a seeded run draws the same numbers.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_bound = False


def _L():
    global _bound
    L = _lib.lib()
    if not _bound:
        L.svgir_sample_incident_rays.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.svgir_sample_incident_rays.restype = C.c_int
        _bound = True
    return L


def fibonacci_sphere_sampling(normals: torch.Tensor, sample_num: int, random_rotate: bool = True, rand_u=None):
    """normals [...,3] -> (incident_dirs [...,sample_num,3], incident_areas [...,sample_num,1])."""
    if not normals.is_cuda:
        raise RuntimeError("svgir_b200.sampling needs CUDA tensors (no CPU fallback)")
    L = _L()
    pre_shape = normals.shape[:-1]
    n = normals.reshape(-1, 3)
    if n.dtype != torch.float32:
        n = n.float()
    n = n.contiguous()
    N = n.shape[0]
    dev = n.device
    u = None
    if random_rotate:
        u = rand_u if rand_u is not None else torch.rand(*pre_shape, 1, device=dev)
        u = u.reshape(-1).float().contiguous()
    dirs = torch.empty((N, sample_num, 3), dtype=torch.float32, device=dev)
    areas = torch.empty((N, sample_num, 1), dtype=torch.float32, device=dev)
    if N:
        with torch.cuda.device(dev):
            _lib.check(L.svgir_sample_incident_rays(N, int(sample_num), n.data_ptr(), None if u is None else u.data_ptr(),
                                                    dirs.data_ptr(), areas.data_ptr(),
                                                    C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                       "sample_incident_rays")
    return dirs.reshape(*pre_shape, sample_num, 3), areas.reshape(*pre_shape, sample_num, 1)


def sample_incident_rays(normals, is_training=False, sample_num=24):
    """scene/gaussian_model.py:23-31."""
    return fibonacci_sphere_sampling(normals, sample_num, random_rotate=bool(is_training))
