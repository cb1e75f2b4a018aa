"""The narrowest drop-in seam, EXECUTED: `svgss_rasterization._C` / `rgss_rasterization._C` called with the positional
tuples the reference's own wrapper builds (gaussian_renderer/svgss_rasterization.py:139-176 -> 24 arguments, 12 results
in the C++ order `depth` before `opacity`; :211-262 -> 31 arguments, 13 results), including the round trip of the three
opaque state buffers through torch.autograd's save_for_backward, and compared with svgir_b200.raster.forward /
backward on the same inputs. INTEGRATION.md section 1 tells a maintainer to bind exactly this object."""
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


class _RefShapedFunction(torch.autograd.Function):
    """What the reference's _RasterizeGaussians does with `_C` (svgss_rasterization.py:92-310), tuple for tuple."""

    @staticmethod
    def forward(ctx, _C, means3D, means2D, features, vfeatures, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, viewmatrix, projmatrix, campos, rs):
        args = (rs["bg"], means3D, features, vfeatures, colors_precomp, opacities, scales, rotations, rs["scale_modifier"],
                cov3Ds_precomp, viewmatrix, projmatrix, rs["prcppoint"], rs["patch_bbox"], rs["tanfovx"], rs["tanfovy"],
                rs["image_height"], rs["image_width"], sh, rs["sh_degree"], campos, rs["prefiltered"], rs["debug"], rs["config"])
        assert len(args) == 24
        res = _C.rasterize_gaussians(*args)
        assert len(res) == 12
        (num_rendered, color, normal, depth, opacity, feature, vfeature, weights, radii, geomBuffer, binningBuffer,
         imgBuffer) = res
        ctx.rs, ctx.num_rendered, ctx._C = rs, num_rendered, _C
        ctx.save_for_backward(colors_precomp, means3D, features, vfeatures, scales, rotations, cov3Ds_precomp, radii, sh,
                              geomBuffer, binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(weights, radii)
        return color, normal, opacity, depth, feature, vfeature, weights, radii

    @staticmethod
    def backward(ctx, g_color, g_normal, g_opacity, g_depth, g_feature, g_vfeature, g_weights, g_radii):
        rs = ctx.rs
        (colors_precomp, means3D, features, vfeatures, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
         binningBuffer, imgBuffer) = ctx.saved_tensors
        args = (rs["bg"], means3D, features, vfeatures, radii, colors_precomp, scales, rotations, rs["scale_modifier"],
                cov3Ds_precomp, rs["viewmatrix"], rs["projmatrix"], rs["prcppoint"], rs["patch_bbox"], rs["tanfovx"],
                rs["tanfovy"], g_color, g_normal, g_depth, g_opacity, g_feature, g_vfeature, sh, rs["sh_degree"],
                rs["campos"], geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, rs["debug"], rs["config"])
        assert len(args) == 31
        res = ctx._C.rasterize_gaussians_backward(*args)
        assert len(res) == 13
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_features, grad_vfeatures, grad_cov3Ds_precomp,
         grad_sh, grad_scales, grad_rotations, grad_viewmat, grad_projmat, grad_campos) = res
        return (None, grad_means3D, grad_means2D, grad_features, grad_vfeatures, grad_sh, None, grad_opacities, grad_scales,
                grad_rotations, None, grad_viewmat, grad_projmat, grad_campos, None)


def test_svgss_C_seam_forward_backward_matches_raster_module():
    import svgss_rasterization as pkg
    from svgir_b200 import raster
    case = util.make_case(20000, 256, 192, S=4, VS=52, seed=5)
    g = util.pixel_grads(case)
    t = util.to_cuda(case)
    cam = case["cam"]
    rs = dict(bg=t["bg"], scale_modifier=1.0, prcppoint=t["prcppoint"], patch_bbox=t["patch_bbox"], tanfovx=cam.tanfovx,
              tanfovy=cam.tanfovy, image_height=cam.H, image_width=cam.W, sh_degree=3, prefiltered=False, debug=False,
              config=t["config"], viewmatrix=t["viewmatrix"], projmatrix=t["projmatrix"], campos=t["campos"])
    leaves = {k: t[k].clone().requires_grad_(True) for k in ("means3D", "opacity", "scales", "rotations", "shs", "features",
                                                            "vfeatures")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    empty = torch.Tensor([]).cuda()
    outs = _RefShapedFunction.apply(pkg._C, leaves["means3D"], m2d, leaves["features"], leaves["vfeatures"], leaves["shs"],
                                    empty, leaves["opacity"], leaves["scales"], leaves["rotations"], empty, t["viewmatrix"],
                                    t["projmatrix"], t["campos"], rs)
    color, normal, opacity, depth, feature, vfeature, weights, radii = outs
    gt = {k: torch.from_numpy(v).cuda() for k, v in g.items()}
    torch.autograd.backward([color, normal, opacity, depth, feature, vfeature],
                            [gt["dL_dcolor"], gt["dL_dnormal"], gt["dL_dopacity"], gt["dL_ddepth"], gt["dL_dfeature"],
                             gt["dL_dvfeature"]])
    # the same inputs through the module the packages are built on
    out, st, bw = util.run_ours(case, grads=g)
    for k, v in (("color", color), ("normal", normal), ("opacity", opacity), ("depth", depth), ("feature", feature),
                 ("vfeature", vfeature)):
        assert torch.equal(out[k], v.detach()), k
    assert torch.equal(out["radii"], radii)
    for k, leaf in (("dL_dmeans3D", leaves["means3D"]), ("dL_dopacity", leaves["opacity"]), ("dL_dscales", leaves["scales"]),
                    ("dL_drotations", leaves["rotations"]), ("dL_dsh", leaves["shs"]), ("dL_dfeatures", leaves["features"]),
                    ("dL_dvfeatures", leaves["vfeatures"]), ("dL_dmeans2D", m2d)):
        a, b = leaf.grad.reshape(-1), bw[k].reshape(-1)
        assert float((a - b).norm() / b.norm().clamp_min(1e-20)) < 1e-3, k
    # two outstanding forwards, backward of the OLDER one (train.py:173-174 renders twice per visualisation step)
    r1 = pkg._C.rasterize_gaussians(rs["bg"], t["means3D"], t["features"], t["vfeatures"], empty, t["opacity"], t["scales"],
                                    t["rotations"], 1.0, empty, t["viewmatrix"], t["projmatrix"], rs["prcppoint"],
                                    rs["patch_bbox"], cam.tanfovx, cam.tanfovy, cam.H, cam.W, t["shs"], 3, t["campos"], False,
                                    False, rs["config"])
    r2 = pkg._C.rasterize_gaussians(rs["bg"], t["means3D"] * 1.01, t["features"], t["vfeatures"], empty, t["opacity"],
                                    t["scales"], t["rotations"], 1.0, empty, t["viewmatrix"], t["projmatrix"],
                                    rs["prcppoint"], rs["patch_bbox"], cam.tanfovx, cam.tanfovy, cam.H, cam.W, t["shs"], 3,
                                    t["campos"], False, False, rs["config"])
    assert int(r1[9].reshape(-1)[0]) != int(r2[9].reshape(-1)[0])
    res = pkg._C.rasterize_gaussians_backward(
        rs["bg"], t["means3D"], t["features"], t["vfeatures"], r1[8], empty, t["scales"], t["rotations"], 1.0, empty,
        t["viewmatrix"], t["projmatrix"], rs["prcppoint"], rs["patch_bbox"], cam.tanfovx, cam.tanfovy, gt["dL_dcolor"],
        gt["dL_dnormal"], gt["dL_ddepth"], gt["dL_dopacity"], gt["dL_dfeature"], gt["dL_dvfeature"], t["shs"], 3, t["campos"],
        r1[9], r1[0], r1[10], r1[11], False, rs["config"])
    assert float((res[3] - bw["dL_dmeans3D"]).norm() / bw["dL_dmeans3D"].norm()) < 1e-3
    assert pkg._C.mark_visible(t["means3D"], t["viewmatrix"], t["projmatrix"]).sum() == 0   # rasterizer_impl.cu:54-66


def test_rgss_C_seam_arity_and_results():
    """rgss: 23 arguments -> 14 results, 27 -> 9 (rgss-rasterization/rasterize_points.cu:36-60,142,145-173,242)."""
    import rgss_rasterization as pkg
    case = util.make_case(8000, 160, 128, S=5, VS=0, seed=6)
    g = util.pixel_grads(case)
    t = util.to_cuda(case)
    cam = case["cam"]
    empty = torch.Tensor([]).cuda()
    res = pkg._C.rasterize_gaussians(t["bg"], t["means3D"], t["features"], empty, t["opacity"], t["scales"], t["rotations"], 1.0,
                                     empty, t["viewmatrix"], t["projmatrix"], cam.tanfovx, cam.tanfovy, cam.W / 2.0, cam.H / 2.0,
                                     cam.H, cam.W, t["shs"], 3, t["campos"], False, True, False)
    assert len(res) == 14
    (num_rendered, num_contrib, color, normal, opacity, depth, feature, pseudo_normal, surface_xyz, weights, radii, geomBuffer,
     binningBuffer, imgBuffer) = res
    assert int(num_rendered) > 0 and num_contrib.shape == (cam.H, cam.W)
    gt = {k: torch.from_numpy(v).cuda() for k, v in g.items()}
    bw = pkg._C.rasterize_gaussians_backward(
        t["bg"], t["means3D"], t["features"], radii, empty, t["scales"], t["rotations"], 1.0, empty, t["viewmatrix"],
        t["projmatrix"], cam.tanfovx, cam.tanfovy, gt["dL_dcolor"], gt["dL_dnormal"], gt["dL_dopacity"], gt["dL_ddepth"],
        gt["dL_dfeature"], t["shs"], 3, t["campos"], geomBuffer, num_rendered, binningBuffer, imgBuffer, True, False)
    assert len(bw) == 9
    for x in bw:
        assert torch.isfinite(x).all()
    assert float(bw[3].abs().sum()) > 0   # dL_dmeans3D
