"""Host-side mirror of the reference's stage-2 render function for the hot path.

`render_view` follows gaussian_renderer/svgss.py:15-262 (`render_view`) step for step -- shade every
surfel, pack `features` / `vfeatures`, rasterise, un-premultiply, split the G-buffer -- but calls
the fused CUDA kernels (svgir_b200.shading, svgss_rasterization) and takes plain tensors instead
of the reference's GaussianModel / Camera objects (those stay in the application).
`training_step` adds the image loss and the backward pass; it is the unit bench.py times.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import shading
from ._lib import launch_count


@dataclass
class SurfelModel:
    """Tensors the reference reads through GaussianModel's getters (gaussian_model.py:270-351)."""
    xyz: torch.Tensor             # [P,3]
    opacity: torch.Tensor         # [P,1]
    scaling: torch.Tensor         # [P,3]
    rotation: torch.Tensor        # [P,4]
    shs: torch.Tensor             # [P,16,3]
    base_color: torch.Tensor      # [P,12] channel-major (R v0..3, G v0..3, B v0..3)
    roughness: torch.Tensor       # [P,4]
    shading_normal: torch.Tensor  # [P,4,3]
    radiance: torch.Tensor        # [P,Ns,3]
    visibility: torch.Tensor      # [P,Ns,1]
    incident_dirs: torch.Tensor   # [P,Ns,3]
    incident_areas: torch.Tensor  # [P,Ns,1]
    active_sh_degree: int = 3
    config: tuple = (1.0, 1.0, 1.0)

    def trainable(self):
        return [self.xyz, self.opacity, self.scaling, self.rotation, self.shs, self.base_color, self.roughness,
                self.shading_normal]


@dataclass
class ViewCamera:
    """Fields of scene/cameras.py:Camera the render function reads."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    world_view_transform: torch.Tensor  # [4,4]
    full_proj_transform: torch.Tensor   # [4,4]
    camera_center: torch.Tensor         # [3]
    patch_bbox: torch.Tensor            # [4]
    prcppoint: torch.Tensor             # [2]


def rgb_to_srgb(img):
    """utils/graphics_utils.py:198-213."""
    t = torch.tensor(0.0031308, device=img.device)
    return torch.where(img > 0.0031308, torch.pow(torch.max(img, t), 1.0 / 2.4) * 1.055 - 0.055, 12.92 * img).clamp(0, 1)


def camera_from_scene(cam, device) -> ViewCamera:
    d = lambda a: torch.from_numpy(a).to(device)
    return ViewCamera(cam.H, cam.W, cam.tanfovx, cam.tanfovy, d(cam.viewmatrix), d(cam.projmatrix), d(cam.campos),
                      d(cam.patch_bbox), d(cam.prcppoint))


def model_from_scene(cloud, mats, device, requires_grad=True) -> SurfelModel:
    d = lambda a: torch.from_numpy(a).to(device)
    m = SurfelModel(d(cloud.means3D), d(cloud.opacity), d(cloud.scales), d(cloud.rotations), d(cloud.shs),
                    d(mats["base_color"]), d(mats["roughness"]), d(mats["shading_normals"]), d(mats["radiance"]),
                    d(mats["visibility"]), d(mats["incident_dirs"]), d(mats["incident_areas"]))
    if requires_grad:
        for t in m.trainable():
            t.requires_grad_(True)
    return m


def render_view(cam: ViewCamera, pc: SurfelModel, env_light, bg_color: torch.Tensor, scaling_modifier=1.0,
                is_training=True, debug=False) -> dict:
    """svgss.py:15-262 without the application objects. Returns the same result keys the hot path
    produces (render, depth, pbr, normal, opacity, base_color, roughness, diffuse, ...)."""
    from svgss_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    means3D = pc.xyz
    screenspace_points = torch.zeros_like(means3D, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    raster_settings = GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width), tanfovx=cam.tanfovx,
        tanfovy=cam.tanfovy, bg=bg_color, scale_modifier=scaling_modifier, viewmatrix=cam.world_view_transform,
        projmatrix=cam.full_proj_transform, patch_bbox=cam.patch_bbox, prcppoint=cam.prcppoint,
        sh_degree=pc.active_sh_degree, campos=cam.camera_center, prefiltered=False, debug=debug,
        config=torch.tensor(pc.config, dtype=torch.float32, device=means3D.device))
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)

    viewdirs = torch.nn.functional.normalize(cam.camera_center - means3D, dim=-1)
    # shading + the features / vfeatures packing of svgss.py:116-166 in one fused kernel
    features, vfeatures = shading.shade_and_pack(
        pc.base_color, pc.roughness, pc.shading_normal, viewdirs, pc.radiance, env_light, pc.visibility,
        pc.incident_dirs, pc.incident_areas, cam.world_view_transform[:3, :3], is_training=is_training, debug=debug)

    (num_rendered, rendered_image, rendered_normal, rendered_opacity, rendered_depth, rendered_feature,
     rendered_vfeature, weights, radii) = rasterizer(
        means3D=means3D, means2D=screenspace_points, shs=pc.shs, colors_precomp=None, opacities=pc.opacity,
        scales=pc.scaling, rotations=pc.rotation, cov3D_precomp=None, features=features, vfeatures=vfeatures)

    inv_o = 1.0 / rendered_opacity.clamp_min(1e-5)
    rendered_feature = rendered_feature * inv_o
    rendered_vfeature = rendered_vfeature * inv_o

    def opacity_filter(r):
        return r * rendered_opacity + (1 - rendered_opacity) * bg_color[:, None, None]

    res = {}
    if is_training:
        vis, local = rendered_feature.split([1, 3], dim=0)
        pbr, base, shn, rough, diffuse = rendered_vfeature.split([3, 3, 3, 1, 3], dim=0)
        res["diffuse"] = opacity_filter(rgb_to_srgb(diffuse))
    else:
        light, local, vis = rendered_feature.split([3, 3, 1], dim=0)
        pbr, base, shn, rough, direct, indirect = rendered_vfeature.split([3, 3, 3, 1, 3, 3], dim=0)
        res["lights"] = opacity_filter(rgb_to_srgb(light))
        res["direct"] = rgb_to_srgb(direct)
        res["indirect"] = rgb_to_srgb(indirect)
    res.update({
        "render": rendered_image, "depth": rendered_depth,
        "pbr": rgb_to_srgb(pbr * rendered_opacity + (1 - rendered_opacity) * bg_color[:, None, None]),
        "normal": shn, "geo_normal": rendered_normal, "opacity": rendered_opacity,
        "base_color": opacity_filter(rgb_to_srgb(base)), "roughness": opacity_filter(rough),
        "local_lights": opacity_filter(rgb_to_srgb(local)), "visibility": opacity_filter(vis),
        "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
        "num_rendered": num_rendered, "weights": weights, "diffuse_light": vfeatures[:, 40:52] if is_training else None})
    return res


def image_loss(res: dict, gt_image: torch.Tensor, lambda_pbr: float = 1.0) -> torch.Tensor:
    """The L1 part of calculate_loss (svgss.py:280-294) on the splatted colour and the PBR image;
    SSIM / smoothness terms are application-side torch code."""
    return (res["render"] - gt_image).abs().mean() + lambda_pbr * (res["pbr"] - gt_image).abs().mean() + \
        0.02 * (1.0 - (res["normal"] * res["geo_normal"]).sum(0)).mean()


def training_step(cam: ViewCamera, pc: SurfelModel, env_param: torch.Tensor, bg, gt_image, zero_grad=True):
    """One stage-2 iteration of the hot path: shade + rasterise forward, loss, backward. Gradients
    are left in .grad of pc.trainable() and env_param. Returns (loss tensor, result dict)."""
    if zero_grad:
        for t in pc.trainable() + [env_param]:
            t.grad = None
    res = render_view(cam, pc, (env_param, shading.MODE_LEARNABLE), bg, is_training=True)
    loss = image_loss(res, gt_image)
    loss.backward()
    return loss, res


def smoke_step(P=3000, W=96, H=64, Ns=16) -> dict:
    from . import scene
    dev = torch.device("cuda:0")
    cloud = scene.make_surfels(P, seed=3)
    mats = scene.make_materials(cloud, Ns, seed=4, env_hw=(16, 32))
    pc = model_from_scene(cloud, mats, dev)
    cam = camera_from_scene(scene.look_at_camera(W, H, 1), dev)
    env = torch.from_numpy(mats["env_param"]).to(dev).requires_grad_(True)
    bg = torch.zeros(3, device=dev)
    gt = torch.rand(3, H, W, device=dev)
    launch_count(reset=True)
    loss, res = training_step(cam, pc, env, bg, gt)
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    for t in pc.trainable() + [env]:
        assert t.grad is not None and torch.isfinite(t.grad).all()
    return {"loss": float(loss), "launches": launch_count(), "num_rendered": res["num_rendered"]}
