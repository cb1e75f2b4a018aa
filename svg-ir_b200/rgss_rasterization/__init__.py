"""Drop-in replacement of the reference's stage-1 rasteriser package (`rgss_rasterization`).

The reference application resolves this rasteriser by JIT-compiling `rgss-rasterization/` through
`torch.utils.cpp_extension.load(name='rgss_rasterization', ...)` (gaussian_renderer/rgss_rasterization.py:7-24);
for a no-edit drop-in, patch that `load` to return this module's `_C`, or
import `GaussianRasterizationSettings` / `GaussianRasterizer` from here directly.

Exports with the reference's names, argument order and error behaviour:
  * `_C`: `rasterize_gaussians` (23 positional args -> 14-tuple), `rasterize_gaussians_backward`
    (27 args -> 9-tuple), `mark_visible` (rgss-rasterization/rasterize_points.cu:36-60,142,145-173,242);
  * `GaussianRasterizationSettings`, `GaussianRasterizer`, `rasterize_gaussians`, `_RasterizeGaussians`
    (gaussian_renderer/rgss_rasterization.py:30-262).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import NamedTuple

import torch
import torch.nn as nn

from svgir_b200 import raster as _raster
from svgir_b200._lib import VARIANT_RGSS


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def _settings(bg, scale_modifier, viewmatrix, projmatrix, tanfovx, tanfovy, cx, cy, H, W, degree, campos,
              prefiltered, computer_pseudo_normal, debug, backward_geometry=True):
    return _raster.RasterSettings(
        image_height=H, image_width=W, tanfovx=tanfovx, tanfovy=tanfovy, bg=bg, scale_modifier=scale_modifier,
        viewmatrix=viewmatrix, projmatrix=projmatrix, sh_degree=degree, campos=campos, prefiltered=prefiltered,
        debug=debug, variant=VARIANT_RGSS, backward_geometry=backward_geometry,
        computer_pseudo_normal=computer_pseudo_normal, cx=cx, cy=cy)


class _CompatC:
    def __init__(self, capacity: int = 16):
        self._states: "OrderedDict[int, _raster.RasterState]" = OrderedDict()
        self._next = 1
        self._capacity = capacity

    def _register(self, st) -> torch.Tensor:
        h = self._next
        self._next += 1
        self._states[h] = st
        while len(self._states) > self._capacity:
            self._states.popitem(last=False)
        return torch.tensor([h], dtype=torch.int64)

    def rasterize_gaussians(self, background, means3D, features, colors, opacity, scales, rotations,
                            scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, cx, cy,
                            image_height, image_width, sh, degree, campos, prefiltered, computer_pseudo_normal,
                            debug):
        s = _settings(background, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, cx, cy,
                      image_height, image_width, degree, campos, prefiltered, computer_pseudo_normal, debug)
        out, st = _raster.forward(s, means3D, opacity, scales, rotations, cov3D_precomp, sh, colors, features, None)
        handle = self._register(st)
        empty = torch.empty((0,), dtype=torch.uint8)
        # the reference returns n_contrib as a from_blob alias into imgBuffer
        # (rgss-rasterization/rasterize_points.cu:138-141); here it is a real tensor
        return (st.num_rendered, out["n_contrib"], out["color"], out["normal"], out["opacity"], out["depth"],
                out["feature"], out["pseudo_normal"], out["surface_xyz"], out["weights"], out["radii"], handle,
                empty, empty.clone())

    def rasterize_gaussians_backward(self, background, means3D, features, radii, colors, scales, rotations,
                                     scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                     dL_dout_color, dL_dout_normal, dL_dout_opacity, dL_dout_depth,
                                     dL_dout_feature, sh, degree, campos, geomBuffer, R, binningBuffer,
                                     imageBuffer, backward_geometry, debug):
        h = int(geomBuffer.reshape(-1)[0].item())
        st = self._states.get(h)
        if st is None:
            raise RuntimeError("rgss_rasterization: forward state for this backward call was evicted")
        st.cfg.backward_geometry = int(bool(backward_geometry))
        r = _raster.backward(st, radii, dict(dL_dcolor=dL_dout_color, dL_dnormal=dL_dout_normal,
                                             dL_ddepth=dL_dout_depth, dL_dopacity=dL_dout_opacity,
                                             dL_dfeature=dL_dout_feature))
        return (r["dL_dmeans2D"], r["dL_dcolors"], r["dL_dopacity"], r["dL_dmeans3D"], r["dL_dfeatures"],
                r["dL_dcov3D"], r["dL_dsh"], r["dL_dscales"], r["dL_drotations"])

    def mark_visible(self, means3D, viewmatrix, projmatrix):
        return _raster.mark_visible(VARIANT_RGSS, means3D, viewmatrix, projmatrix)


_C = _CompatC()


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    cx: float
    cy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    backward_geometry: bool
    computer_pseudo_normal: bool
    debug: bool


def rasterize_gaussians(means3D, means2D, features, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, features, sh, colors_precomp, opacities, scales,
                                     rotations, cov3Ds_precomp, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, features, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        rs = raster_settings
        s = _settings(rs.bg, rs.scale_modifier, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.cx, rs.cy,
                      rs.image_height, rs.image_width, rs.sh_degree, rs.campos, rs.prefiltered,
                      rs.computer_pseudo_normal, rs.debug, rs.backward_geometry)
        args = (means3D, opacities, scales, rotations, cov3Ds_precomp, sh, colors_precomp, features, None)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args + tuple(rs))
            try:
                out, st = _raster.forward(s, *args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            out, st = _raster.forward(s, *args)
        ctx.raster_settings = rs
        ctx.num_rendered = st.num_rendered
        ctx.state = st
        ctx.save_for_backward(out["radii"])
        ctx.mark_non_differentiable(out["n_contrib"], out["pseudo_normal"], out["surface_xyz"], out["weights"],
                                    out["radii"])
        return (st.num_rendered, out["n_contrib"], out["color"], out["normal"], out["opacity"], out["depth"],
                out["feature"], out["pseudo_normal"], out["surface_xyz"], out["weights"], out["radii"])

    @staticmethod
    def backward(ctx, grad_num_rendered, grad_num_contrib, grad_out_color, grad_out_normal, grad_out_opacity,
                 grad_out_depth, grad_out_feature, grad_out_pseudo_normal, grad_out_surface_xyz, grad_out_weights,
                 grad_out_radii):
        (radii,) = ctx.saved_tensors
        grads_in = dict(dL_dcolor=grad_out_color, dL_dnormal=grad_out_normal, dL_ddepth=grad_out_depth,
                        dL_dopacity=grad_out_opacity, dL_dfeature=grad_out_feature)
        if ctx.raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(tuple(v for v in grads_in.values() if v is not None))
            try:
                r = _raster.backward(ctx.state, radii, grads_in)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            r = _raster.backward(ctx.state, radii, grads_in)
        has = ctx.state.cin
        return (r["dL_dmeans3D"], r["dL_dmeans2D"], r["dL_dfeatures"], r["dL_dsh"] if has.shs else None,
                r["dL_dcolors"] if has.colors_precomp else None, r["dL_dopacity"],
                r["dL_dscales"] if has.scales else None, r["dL_drotations"] if has.rotations else None,
                r["dL_dcov3D"] if has.cov3D_precomp else None, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, features=None):
        raster_settings = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        if features is None:
            features = torch.empty_like(means3D[..., :0])
        return rasterize_gaussians(means3D, means2D, features, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, raster_settings)
