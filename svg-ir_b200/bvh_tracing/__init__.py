"""Drop-in `bvh_tracing` package (the name submodules/bvh/setup.py installs and
submodules/bvh/__init__.py:9 imports `_C` from).

`_C` exposes the three functions of submodules/bvh/src/bindings.cpp:9-11 with the reference's
positional signatures (src/bvh.h:5-18) over torch tensors; `RayTracer` mirrors
submodules/bvh/bvh_tracing/__init__.py:11-58 (no origin offset -- the application's own copy of
the class, submodules/bvh/__init__.py:28-71, adds 0.05 d and is `svgir_b200.bvh.RayTracer`).
Everything runs in libsvgir_b200.so; there is no CPU fallback.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from svgir_b200 import bvh as _bvh


class _CModule:
    """bvh_tracing._C"""

    def __init__(self, capacity: int = 8):
        # The reference's RayTracer keeps only the `nodes` / `aabbs` tensors (submodules/bvh/__init__.py:59-60), so
        # trace_bvh_opacity(nodes, aabbs, ...) must find the packed traversal records create_bvh built from the
        # tensor alone: a small LRU of strongly held trees keyed by the storage address, each remembering the
        # tensor it was built for. Holding `nodes` alive means its address cannot be recycled while the entry
        # exists; a hit is validated by identity of the storage, shape and version.
        self._trees: "OrderedDict[int, object]" = OrderedDict()
        self._capacity = capacity

    def create_bvh(self, means3D, scales, rotations, nodes, aabbs):
        """src/bvh.cu:9-27: nodes / aabbs are updated in place and returned with the Morton codes."""
        tree = _bvh.Bvh(nodes, aabbs)
        if tree.nodes.data_ptr() != nodes.data_ptr():  # caller passed a non-contiguous view: copy back
            nodes.copy_(tree.nodes)
            aabbs.copy_(tree.aabbs)
        tree._owner_nodes = nodes
        tree._owner_version = nodes._version
        self._trees.pop(nodes.data_ptr(), None)
        self._trees[nodes.data_ptr()] = tree
        while len(self._trees) > self._capacity:
            self._trees.popitem(last=False)
        return nodes, aabbs, tree.morton

    def _find(self, nodes, aabbs):
        tree = self._trees.get(nodes.data_ptr())
        if (tree is None or tree.P != (nodes.shape[0] + 1) // 2 or tuple(nodes.shape) != tuple(tree._owner_nodes.shape) or
                nodes._version != tree._owner_version):
            raise RuntimeError("trace_bvh_opacity: `nodes` was not produced by this module's create_bvh (or was modified "
                               "since, or more than %d trees were built after it)" % self._capacity)
        self._trees.move_to_end(nodes.data_ptr())
        return tree

    def trace_bvh_opacity(self, nodes, aabbs, rays_o, rays_d, means3D, covs3D, opacities, normals):
        """src/bvh.cu:89-116 -> (num_contributes int32, rendered_opacity float32), shaped like rays_o[..., 0]."""
        return self._find(nodes, aabbs).trace_opacity(rays_o, rays_d, means3D, covs3D, opacities, normals)

    def trace_bvh(self, nodes, aabbs, rays_o, rays_d, means3D, covs3D, opacities):
        """src/bvh.cu:29-87 (per-ray hit lists). Unused by the application and not on the hot
        path; not provided."""
        raise NotImplementedError("bvh_tracing._C.trace_bvh is not part of the svgir_b200 hot path")


_C = _CModule()


class RayTracer(_bvh.RayTracer):
    ray_offset = 0.0
