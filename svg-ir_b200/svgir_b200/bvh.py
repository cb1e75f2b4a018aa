"""Host side of the surfel LBVH and the visibility trace (C ABI: svgir_bvh_* in include/svgir_b200.h).

New equivalent of the reference's torch glue for submodules/bvh: `create_bvh` / `trace_bvh_opacity`
(src/bvh.cu:9-27, 89-116) and the tensor code of `RayTracer` (submodules/bvh/__init__.py:28-71).
PyTorch provides device memory and the current stream; all arithmetic is in libsvgir_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib


class BvhStruct(C.Structure):
    _fields_ = [("P", C.c_int32), ("reserved_", C.c_int32), ("nodes", C.c_void_p), ("aabbs", C.c_void_p),
                ("morton", C.c_void_p), ("packed", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t)]


_BOUND = False


def _L():
    global _BOUND
    L = _lib.lib()
    if not _BOUND:
        vp = C.c_void_p
        L.svgir_bvh_workspace_bytes.argtypes = [C.c_int]
        L.svgir_bvh_workspace_bytes.restype = C.c_size_t
        L.svgir_bvh_leaf_aabbs.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp]
        L.svgir_bvh_leaf_aabbs.restype = C.c_int
        L.svgir_bvh_build.argtypes = [C.POINTER(BvhStruct), vp]
        L.svgir_bvh_build.restype = C.c_int
        L.svgir_bvh_pack_leaves.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp]
        L.svgir_bvh_pack_leaves.restype = C.c_int
        L.svgir_bvh_trace_opacity.argtypes = [C.POINTER(BvhStruct), C.c_longlong, vp, vp, C.c_int, C.c_float, vp, vp,
                                              vp, vp]
        L.svgir_bvh_trace_opacity.restype = C.c_int
        _BOUND = True
    return L


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("svgir_b200 BVH needs CUDA tensors (no CPU fallback)")
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def leaf_aabbs(means3D, scales, rotations) -> Tuple[torch.Tensor, torch.Tensor]:
    """RayTracer.__init__ lines 31-57: returns (nodes [2P-1,5] int32, aabbs [2P-1,6]) with the node
    table initialised (-1 / counts 0,1) and the 8-corner leaf boxes in rows P-1.."""
    L = _L()
    means3D, scales, rotations = _f32c(means3D), _f32c(scales), _f32c(rotations)
    P, dev = means3D.shape[0], means3D.device
    nodes = torch.empty((max(2 * P - 1, 0), 5), dtype=torch.int32, device=dev)
    aabbs = torch.empty((max(2 * P - 1, 0), 6), dtype=torch.float32, device=dev)
    if P:
        with torch.cuda.device(dev):
            _lib.check(L.svgir_bvh_leaf_aabbs(P, means3D.data_ptr(), scales.data_ptr(), rotations.data_ptr(),
                                              nodes.data_ptr(), aabbs.data_ptr(), _stream(dev)), "bvh_leaf_aabbs")
    return nodes, aabbs


class Bvh:
    """A built tree: reference-layout `nodes`, `aabbs`, `morton` + the packed traversal records."""

    def __init__(self, nodes: torch.Tensor, aabbs: torch.Tensor):
        L = _L()
        if nodes.dtype != torch.int32 or aabbs.dtype != torch.float32 or not nodes.is_cuda:
            raise RuntimeError("create_bvh: nodes must be CUDA int32 [2P-1,5], aabbs CUDA float32 [2P-1,6]")
        if nodes.dim() != 2 or nodes.shape[1] != 5 or aabbs.shape != (nodes.shape[0], 6) or nodes.shape[0] % 2 == 0:
            raise RuntimeError("create_bvh: nodes must be [2P-1,5] and aabbs [2P-1,6]")
        self.nodes = nodes if nodes.is_contiguous() else nodes.contiguous()
        self.aabbs = aabbs if aabbs.is_contiguous() else aabbs.contiguous()
        dev = nodes.device
        P = (nodes.shape[0] + 1) // 2
        self.P = P
        self.morton = torch.zeros((P,), dtype=torch.int64, device=dev)
        self.packed = torch.zeros((max(P - 1, 1), 16), dtype=torch.float32, device=dev)
        ws_bytes = int(L.svgir_bvh_workspace_bytes(P))
        ws = torch.empty((ws_bytes + 256,), dtype=torch.uint8, device=dev)
        base = (ws.data_ptr() + 255) // 256 * 256
        self.c = BvhStruct(P, 0, self.nodes.data_ptr(), self.aabbs.data_ptr(), self.morton.data_ptr(),
                           self.packed.data_ptr(), base, ws_bytes)
        with torch.cuda.device(dev):
            _lib.check(L.svgir_bvh_build(C.byref(self.c), _stream(dev)), "bvh_build")
        ws.record_stream(torch.cuda.current_stream(dev))
        self.c.workspace, self.c.workspace_bytes = None, 0
        self._leaf_key = None
        self._leaf_rec = None

    def pack_leaves(self, means3D, symm_inv, opacity, normals) -> torch.Tensor:
        """[P,16] leaf records; cached while the four tensors are unchanged (same storage + version)."""
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in (means3D, symm_inv, opacity, normals))
        if key == self._leaf_key and self._leaf_rec is not None:
            return self._leaf_rec
        L = _L()
        m, ci, op, n = _f32c(means3D), _f32c(symm_inv), _f32c(opacity).reshape(-1), _f32c(normals)
        P = self.P
        if m.shape != (P, 3) or ci.shape != (P, 6) or op.numel() != P or n.shape != (P, 3):
            raise RuntimeError("trace_bvh_opacity: means3D [P,3], covs3D [P,6], opacities [P], normals [P,3] expected")
        rec = torch.empty((P, 16), dtype=torch.float32, device=m.device)
        with torch.cuda.device(m.device):
            _lib.check(L.svgir_bvh_pack_leaves(P, m.data_ptr(), ci.data_ptr(), op.data_ptr(), n.data_ptr(),
                                               rec.data_ptr(), _stream(m.device)), "bvh_pack_leaves")
        self._leaf_key, self._leaf_rec = key, rec
        return rec

    def trace_opacity(self, rays_o, rays_d, means3D, symm_inv, opacity, normals, origin_offset: float = 0.0):
        """trace_bvh_opacity (src/bvh.cu:89-116): returns (contributes int32, visibility float32) shaped
        like rays_o without its last axis. A rays_o that is an expand() of per-surfel origins over the
        sample axis ([N,1,3] -> [N,Ns,3], as gaussian_model.py:449 passes it) is read in place."""
        L = _L()
        shape = tuple(rays_o.shape[:-1])
        dev = rays_d.device
        rays_d = _f32c(rays_d)
        n_rays = rays_d.numel() // 3
        per_origin = 1
        if rays_o.dim() == 3 and rays_o.stride(1) == 0 and rays_o.shape[1] > 1 and rays_o.stride(2) == 1 \
                and rays_o.stride(0) == 3:
            per_origin = int(rays_o.shape[1])
            ro = rays_o[:, 0, :]
            if ro.dtype != torch.float32:
                ro = ro.float()
            ro = ro.detach()
        else:
            ro = _f32c(rays_o)
        contrib = torch.empty(shape, dtype=torch.int32, device=dev)
        vis = torch.empty(shape, dtype=torch.float32, device=dev)
        if n_rays == 0:
            return contrib, vis
        rec = self.pack_leaves(means3D, symm_inv, opacity, normals)
        with torch.cuda.device(dev):
            _lib.check(L.svgir_bvh_trace_opacity(C.byref(self.c), n_rays, ro.data_ptr(), rays_d.data_ptr(), per_origin,
                                                 float(origin_offset), rec.data_ptr(), contrib.data_ptr(),
                                                 vis.data_ptr(), _stream(dev)), "bvh_trace_opacity")
        return contrib, vis


class RayTracer:
    """submodules/bvh/__init__.py:28-71 (the class gaussian_model.py:17 imports). `ray_offset` is the
    0.05 the application's copy adds to the ray origins (:63); the installed `bvh_tracing.RayTracer`
    (submodules/bvh/bvh_tracing/__init__.py:48-58) uses 0."""
    ray_offset = 0.05

    def __init__(self, means3D, scales, rotations):
        nodes, aabbs = leaf_aabbs(means3D, scales, rotations)
        self.bvh = Bvh(nodes, aabbs)
        self.tree, self.aabb, self.morton = self.bvh.nodes, self.bvh.aabbs, self.bvh.morton

    @torch.no_grad()
    def trace_visibility(self, rays_o, rays_d, means3D, symm_inv, opacity, normals):
        cotrib, opa = self.bvh.trace_opacity(rays_o, rays_d, means3D, symm_inv, opacity, normals,
                                             origin_offset=self.ray_offset)
        return {"visibility": opa.unsqueeze(-1), "contribute": cotrib.unsqueeze(-1)}
