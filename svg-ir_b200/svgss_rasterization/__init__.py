"""Drop-in replacement of the reference's stage-2 rasteriser package.

Put the directory `svg-ir_b200/` on sys.path and the unmodified reference application
(`gaussian_renderer/svgss.py:12,69,171`) resolves

    from svgss_rasterization import _C                        (gaussian_renderer/svgss_rasterization.py:8)
    GaussianRasterizationSettings / GaussianRasterizer          (same module, :331-411)

to this package.  Exports, with the reference's names, argument order and error behaviour:

  * `_C` -- object with `rasterize_gaussians` (24 positional args -> 12-tuple),
    `rasterize_gaussians_backward` (31 args -> 13-tuple) and `mark_visible`
    (svgss_rasterization/rasterize_points.h:18-87);
  * `GaussianRasterizationSettings`, `GaussianRasterizer`, `rasterize_gaussians`,
    `_RasterizeGaussians` (gaussian_renderer/svgss_rasterization.py:57-411).

Everything is computed by libsvgir_b200.so through svgir_b200.raster; there is no fallback.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import NamedTuple

import torch
import torch.nn as nn

from svgir_b200 import raster as _raster
from svgir_b200._lib import VARIANT_SVGSS


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


class _CompatC:
    """`_C`-compatible seam: same positional signatures as the reference's pybind module.

    The reference returns three opaque byte buffers that Python saves for backward; here
    `geomBuffer` is a one-element int64 CPU tensor holding a handle into a small LRU registry of
    typed state objects (multiple outstanding forwards are supported, cf. train.py:173-174).
    """

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

    def rasterize_gaussians(self, background, means3D, features, vfeatures, colors, opacity, scales,
                            rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix, prcppoint,
                            patchbbox, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                            prefiltered, debug, config):
        s = _raster.RasterSettings(
            image_height=image_height, image_width=image_width, tanfovx=tan_fovx, tanfovy=tan_fovy,
            bg=background, scale_modifier=scale_modifier, viewmatrix=viewmatrix, projmatrix=projmatrix,
            sh_degree=degree, campos=campos, prefiltered=prefiltered, debug=debug, variant=VARIANT_SVGSS,
            patch_bbox=patchbbox, config=config)
        out, st = _raster.forward(s, means3D, opacity, scales, rotations, cov3D_precomp, sh, colors,
                                  features, vfeatures)
        handle = self._register(st)
        empty = torch.empty((0,), dtype=torch.uint8)
        # C++ tuple order (rasterize_points.cu:144): depth before opacity
        return (st.num_rendered, out["color"], out["normal"], out["depth"], out["opacity"], out["feature"],
                out["vfeature"], out["weights"], out["radii"], handle, empty, empty.clone())

    def rasterize_gaussians_backward(self, background, means3D, features, vfeatures, radii, colors, scales,
                                     rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
                                     prcppoint, patchbbox, tan_fovx, tan_fovy, dL_dout_color, dL_dout_normal,
                                     dL_dout_depth, dL_dout_opac, dL_dout_feature, dL_dout_vfeature, sh, degree,
                                     campos, geomBuffer, R, binningBuffer, imageBuffer, debug, config):
        h = int(geomBuffer.reshape(-1)[0].item())
        st = self._states.get(h)
        if st is None:
            raise RuntimeError("svgss_rasterization: forward state for this backward call was evicted "
                               "(more than %d outstanding forwards)" % self._capacity)
        r = _raster.backward(st, radii, dict(dL_dcolor=dL_dout_color, dL_dnormal=dL_dout_normal,
                                             dL_ddepth=dL_dout_depth, dL_dopacity=dL_dout_opac,
                                             dL_dfeature=dL_dout_feature, dL_dvfeature=dL_dout_vfeature))
        return (r["dL_dmeans2D"], r["dL_dcolors"], r["dL_dopacity"], r["dL_dmeans3D"], r["dL_dfeatures"],
                r["dL_dvfeatures"], r["dL_dcov3D"], r["dL_dsh"], r["dL_dscales"], r["dL_drotations"],
                r["dL_dviewmat"], r["dL_dprojmat"], r["dL_dcampos"])

    def mark_visible(self, means3D, viewmatrix, projmatrix):
        return _raster.mark_visible(VARIANT_SVGSS, means3D, viewmatrix, projmatrix)


_C = _CompatC()


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    patch_bbox: torch.Tensor
    prcppoint: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    config: torch.Tensor
    # Extension (keyword, optional -- the reference's 15 fields above are unchanged): the state returned by
    # `preprocess_geometry()` for the same geometry, so that forward() runs only binning + compositing.
    prestate: object = None


def preprocess_geometry(raster_settings, means3D, opacities, scales=None, rotations=None, cov3D_precomp=None,
                        shs=None, colors_precomp=None):
    """Runs the rasteriser's per-surfel preprocess ahead of the forward call (it needs the geometry only) and
    returns (prestate, (vis_list, vis_count)): the device-side list of the surfels that survive culling.
    Shade only those (svgir_b200.shading.shade_and_pack(work=...)), then pass
    `raster_settings._replace(prestate=prestate)` to GaussianRasterizer. Culled surfels are never read by
    the compositor and receive zero gradients, so images and gradients are unchanged."""
    rs = raster_settings
    s = _raster.RasterSettings(
        image_height=rs.image_height, image_width=rs.image_width, tanfovx=rs.tanfovx, tanfovy=rs.tanfovy,
        bg=rs.bg, scale_modifier=rs.scale_modifier, viewmatrix=rs.viewmatrix, projmatrix=rs.projmatrix,
        sh_degree=rs.sh_degree, campos=rs.campos, prefiltered=rs.prefiltered, debug=rs.debug,
        variant=VARIANT_SVGSS, patch_bbox=rs.patch_bbox, config=rs.config)
    with torch.no_grad():
        st = _raster.preprocess(s, means3D, opacities, scales, rotations, cov3D_precomp, shs, colors_precomp,
                                want_vis_list=True)
    return st, (st.t["vis_list"], st.t["vis_count"])


def rasterize_gaussians(means3D, means2D, sh, features, vfeatures, colors_precomp, opacities, scales,
                        rotations, cov3Ds_precomp, viewmatrix, projmatrix, campos, raster_settings):
    # The reference's wrapper names its parameters (.., sh, features, vfeatures, ..) but forwards
    # them positionally into an apply() whose slots are (.., features, vfeatures, sh, ..); the
    # caller passes them in the latter order, so the mislabel cancels out
    # (gaussian_renderer/svgss_rasterization.py:57-88, 396-411). Same here.
    return _RasterizeGaussians.apply(means3D, means2D, sh, features, vfeatures, colors_precomp, opacities,
                                     scales, rotations, cov3Ds_precomp, viewmatrix, projmatrix, campos,
                                     raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, features, vfeatures, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, viewmatrix, projmatrix, campos, raster_settings):
        rs = raster_settings
        s = _raster.RasterSettings(
            image_height=rs.image_height, image_width=rs.image_width, tanfovx=rs.tanfovx, tanfovy=rs.tanfovy,
            bg=rs.bg, scale_modifier=rs.scale_modifier, viewmatrix=viewmatrix, projmatrix=projmatrix,
            sh_degree=rs.sh_degree, campos=campos, prefiltered=rs.prefiltered, debug=rs.debug,
            variant=VARIANT_SVGSS, patch_bbox=rs.patch_bbox, config=rs.config)
        args = (means3D, opacities, scales, rotations, cov3Ds_precomp, sh, colors_precomp, features, vfeatures)
        pre = getattr(rs, "prestate", None)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args + tuple(rs)[:15])  # copy before they can be corrupted
            try:
                out, st = _raster.forward(s, *args, prestate=pre)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            out, st = _raster.forward(s, *args, prestate=pre)
        ctx.raster_settings = rs
        ctx.num_rendered = st.num_rendered
        ctx.state = st
        ctx.save_for_backward(out["radii"])
        ctx.mark_non_differentiable(out["weights"], out["radii"])
        return (st.num_rendered, out["color"], out["normal"], out["opacity"], out["depth"], out["feature"],
                out["vfeature"], out["weights"], out["radii"])

    @staticmethod
    def backward(ctx, grad_num_rendered, grad_out_color, grad_out_normal, grad_out_opacity, grad_out_depth,
                 grad_out_feature, grad_out_vfeature, grad_out_weights, grad_out_radii):
        (radii,) = ctx.saved_tensors
        grads_in = dict(dL_dcolor=grad_out_color, dL_dnormal=grad_out_normal, dL_ddepth=grad_out_depth,
                        dL_dopacity=grad_out_opacity, dL_dfeature=grad_out_feature,
                        dL_dvfeature=grad_out_vfeature)
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
        # same slot order as the reference (svgss_rasterization.py:293-308)
        return (r["dL_dmeans3D"], r["dL_dmeans2D"], r["dL_dfeatures"], r["dL_dvfeatures"],
                r["dL_dsh"] if has.shs else None, r["dL_dcolors"] if has.colors_precomp else None,
                r["dL_dopacity"], r["dL_dscales"] if has.scales else None,
                r["dL_drotations"] if has.rotations else None,
                r["dL_dcov3D"] if has.cov3D_precomp else None,
                r["dL_dviewmat"], r["dL_dprojmat"], r["dL_dcampos"], None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, features=None, vfeatures=None):
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
        if vfeatures is None:
            vfeatures = torch.empty_like(means3D[..., :0])
        return rasterize_gaussians(means3D, means2D, features, vfeatures, shs, colors_precomp, opacities,
                                   scales, rotations, cov3D_precomp, raster_settings.viewmatrix,
                                   raster_settings.projmatrix, raster_settings.campos, raster_settings)
