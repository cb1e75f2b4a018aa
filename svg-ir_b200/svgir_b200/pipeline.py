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

from . import _lib, losses, raster, shading
from ._lib import launch_count

_CONFIG_CACHE: dict = {}
# render_view default: True = shade every surfel like the reference (svgss.py:116-135); False = shade only
# the surfels that survive the rasteriser's culling (identical images and gradients, see render_view).
SHADE_CULLED = True


def _config_tensor(config, device) -> torch.Tensor:
    """pc.config as a device tensor (svgss.py:66 builds it with torch.tensor(...).cuda() every call: a
    pageable H2D copy per view; cached here, which also keeps render_view capturable into a CUDA graph)."""
    key = (tuple(float(x) for x in config), str(device))
    t = _CONFIG_CACHE.get(key)
    if t is None:
        t = _CONFIG_CACHE[key] = torch.tensor(key[0], dtype=torch.float32, device=device)
    return t


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
    # optional: ONE flat 48-float tensor the five tensors above are views of (`blocked_camera`), so a view's
    # per-step constants move with a single copy instead of five
    block: Optional[torch.Tensor] = None


_CAM_SLICES = ((0, 16, (4, 4)), (16, 32, (4, 4)), (32, 35, (3,)), (36, 40, (4,)), (40, 42, (2,)))


def blocked_camera(image_height, image_width, tanfovx, tanfovy, world_view_transform, full_proj_transform, camera_center,
                   patch_bbox, prcppoint, device=None, pin=False) -> ViewCamera:
    """ViewCamera whose tensors are 16-byte-aligned views into one contiguous fp32 block."""
    src = (world_view_transform, full_proj_transform, camera_center, patch_bbox, prcppoint)
    dev = device if device is not None else src[0].device
    block = torch.zeros(48, dtype=torch.float32, device=dev)
    if pin:
        block = block.pin_memory()
    views = []
    for (lo, hi, shape), t in zip(_CAM_SLICES, src):
        v = block[lo:hi].view(shape)
        v.copy_(t.to(torch.float32))
        views.append(v)
    return ViewCamera(image_height, image_width, tanfovx, tanfovy, *views, block=block)


def rgb_to_srgb(img):
    """utils/graphics_utils.py:198-213."""
    return torch.where(img > 0.0031308, torch.pow(img.clamp_min(0.0031308), 1.0 / 2.4) * 1.055 - 0.055, 12.92 * img).clamp(0, 1)


def camera_from_scene(cam, device) -> ViewCamera:
    d = lambda a: torch.from_numpy(a)
    return blocked_camera(cam.H, cam.W, cam.tanfovx, cam.tanfovy, d(cam.viewmatrix), d(cam.projmatrix), d(cam.campos),
                          d(cam.patch_bbox), d(cam.prcppoint), device=device)


def model_from_scene(cloud, mats, device, requires_grad=True) -> SurfelModel:
    """`mats` from scene.make_materials (numpy) or scene.make_materials_torch (tensors already on the device)."""
    d = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(a)).to(device)
    m = SurfelModel(d(cloud.means3D), d(cloud.opacity), d(cloud.scales), d(cloud.rotations), d(cloud.shs),
                    d(mats["base_color"]), d(mats["roughness"]), d(mats["shading_normals"]), d(mats["radiance"]),
                    d(mats["visibility"]), d(mats["incident_dirs"]), d(mats["incident_areas"]))
    if requires_grad:
        for t in m.trainable():
            t.requires_grad_(True)
    return m


FUSED_VIEWDIRS = True

# render_view, eval branch: True = the G-buffer resolve tail runs as one CUDA kernel (losses.resolve_eval) when no
# gradient is being recorded; False = the torch mirror of gaussian_renderer/svgss.py:187-262.
FUSED_RESOLVE = True


def _shade_and_rasterize(cam: ViewCamera, pc: SurfelModel, env_light, bg_color: torch.Tensor, scaling_modifier, is_training,
                         debug, shade_culled):
    """svgss.py:15-184: shading, feature packing and the rasteriser call; returns the rasteriser's raw 9-tuple
    plus (screenspace_points, vfeatures)."""
    from svgss_rasterization import GaussianRasterizationSettings, GaussianRasterizer, preprocess_geometry
    if shade_culled is None:
        shade_culled = SHADE_CULLED
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
        config=_config_tensor(pc.config, means3D.device))
    work, binning_beside = None, False
    if not shade_culled:
        prestate, work = preprocess_geometry(raster_settings, means3D, pc.opacity, pc.scaling, pc.rotation, None,
                                             pc.shs, None)
        raster_settings = raster_settings._replace(prestate=prestate)
        binning_beside = OVERLAP_BINNING and raster.start_binning(prestate)   # side stream, under the shading (no-op
        #                                                                        without a capacity hint)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)

    # shading + the features / vfeatures packing of svgss.py:116-166 in one fused kernel. FUSED_VIEWDIRS: the view
    # direction F.normalize(camera_center - means3D) of svgss.py:95 is evaluated inside the kernel (and its gradient
    # returned for means3D) instead of by torch ops around it -- the same code path fused_step.FusedTrainStep uses,
    # so both give bit-identical images; False = torch normalise, then the kernel (the reference's op order)
    reserve = BINNING_RESERVE_SMS if (work is not None and binning_beside) else 0
    if reserve:   # the shading grid is persistent: leave a few SMs to the binning kernels running beside it
        _lib.lib().svgir_shade_reserve_sms(int(reserve))
    if FUSED_VIEWDIRS:
        features, vfeatures = shading.shade_and_pack(
            pc.base_color, pc.roughness, pc.shading_normal, None, pc.radiance, env_light, pc.visibility,
            pc.incident_dirs, pc.incident_areas, cam.world_view_transform[:3, :3], is_training=is_training, debug=debug,
            work=work, means3D=means3D, campos=cam.camera_center)
    else:
        viewdirs = torch.nn.functional.normalize(cam.camera_center - means3D, dim=-1)
        features, vfeatures = shading.shade_and_pack(
            pc.base_color, pc.roughness, pc.shading_normal, viewdirs, pc.radiance, env_light, pc.visibility,
            pc.incident_dirs, pc.incident_areas, cam.world_view_transform[:3, :3], is_training=is_training, debug=debug,
            work=work)

    if reserve:
        _lib.lib().svgir_shade_reserve_sms(0)
    raw = rasterizer(
        means3D=means3D, means2D=screenspace_points, shs=pc.shs, colors_precomp=None, opacities=pc.opacity,
        scales=pc.scaling, rotations=pc.rotation, cov3D_precomp=None, features=features, vfeatures=vfeatures)
    return raw, screenspace_points, vfeatures


def render_view(cam: ViewCamera, pc: SurfelModel, env_light, bg_color: torch.Tensor, scaling_modifier=1.0,
                is_training=True, debug=False, shade_culled: Optional[bool] = None) -> dict:
    """svgss.py:15-262 without the application objects. Returns the same result keys the hot path
    produces (render, depth, pbr, normal, opacity, base_color, roughness, diffuse, ...).

    shade_culled=False runs the rasteriser's per-surfel preprocess FIRST and shades only the surfels that
    survive culling (radii > 0): the compositor never reads the others and their gradients are zero, so
    every image and every gradient is unchanged; only the per-surfel by-product `diffuse_light` is zero
    for culled surfels (it feeds the optional lambda_light regulariser, svgss.py:359-364, which the TensoIR
    recipe disables: script/run_tensoir.sh:37-38). Default: the module switch SHADE_CULLED."""
    raw, screenspace_points, vfeatures = _shade_and_rasterize(cam, pc, env_light, bg_color, scaling_modifier,
                                                              is_training, debug, shade_culled)
    (num_rendered, rendered_image, rendered_normal, rendered_opacity, rendered_depth, rendered_feature,
     rendered_vfeature, weights, radii) = raw

    if (not is_training and FUSED_RESOLVE and rendered_feature.is_cuda and rendered_feature.shape[0] == 7 and
            rendered_vfeature.shape[0] == 16 and not (torch.is_grad_enabled() and rendered_vfeature.requires_grad)):
        # eval / relighting frame: the whole torch tail below as one kernel (losses.resolve_eval)
        res = losses.resolve_eval(rendered_opacity, rendered_feature, rendered_vfeature, bg_color)
        res.update({
            "render": rendered_image, "depth": rendered_depth, "geo_normal": rendered_normal, "opacity": rendered_opacity,
            "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
            "num_rendered": num_rendered, "weights": weights,
            "raster_state": num_rendered._st if isinstance(num_rendered, raster.LazyCount) else None, "diffuse_light": None})
        return res

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
        "num_rendered": num_rendered, "weights": weights,
        "raster_state": num_rendered._st if isinstance(num_rendered, raster.LazyCount) else None, "diffuse_light": vfeatures[:, 40:52] if is_training else None})
    return res


def image_loss(res: dict, gt_image: torch.Tensor, lambda_pbr: float = 1.0, cam: Optional[ViewCamera] = None,
               image_mask: Optional[torch.Tensor] = None, lambda_normal: float = 0.02) -> torch.Tensor:
    """The part of calculate_loss (gaussian_renderer/svgss.py:265-313) that drives the hot path: the L1 terms on the
    splatted colour and the PBR image (:280-294) and the 0.02-weighted surface term
    cos_loss(rendered_normal, depth2normal(rendered_depth, image_mask, camera)) (:300-313) -- torch mirror of the
    reference's ops (losses.depth2normal_torch / cos_loss_torch). With cam=None the round-1 stand-in
    mean(1 - <normal, geo_normal>) is used instead of the surface term (A/B only). The SSIM, edge-aware and TV terms
    are losses.fused_ssim / fused_edge_aware / fused_tv."""
    l = (res["render"] - gt_image).abs().mean() + lambda_pbr * (res["pbr"] - gt_image).abs().mean()
    if cam is None:
        return l + lambda_normal * (1.0 - (res["normal"] * res["geo_normal"]).sum(0)).mean()
    H, W = int(cam.image_height), int(cam.image_width)
    d2n = losses.depth2normal_torch(res["depth"], image_mask, H, W, camera_d2n_terms(cam))
    return l + lambda_normal * losses.cos_loss_torch(res["normal"], d2n)


def camera_d2n_terms(cam: ViewCamera):
    """depth2normal's camera terms (losses.d2n_camera_terms) of a ViewCamera; its principal point is read on the host
    once per camera object (prcppoint is (0.5, 0.5) for every TensoIR camera)."""
    t = getattr(cam, "_d2n_terms", None)
    if t is None:
        pp = [float(v) for v in cam.prcppoint.detach().cpu().tolist()] if not torch.cuda.is_current_stream_capturing() else [0.5, 0.5]
        t = losses.d2n_camera_terms(int(cam.image_height), int(cam.image_width), cam.tanfovx, cam.tanfovy, pp)
        try:
            cam._d2n_terms = t
        except Exception:
            pass
    return t


def render_raw(cam: ViewCamera, pc: SurfelModel, env_light, bg_color: torch.Tensor, scaling_modifier=1.0,
               is_training=True, debug=False, shade_culled: Optional[bool] = None) -> dict:
    """render_view up to and including the rasteriser (svgss.py:15-184): the raw, opacity-premultiplied G-buffer
    that `losses.fused_train_loss` consumes directly."""
    raw, screenspace_points, vfeatures = _shade_and_rasterize(cam, pc, env_light, bg_color, scaling_modifier,
                                                              is_training, debug, shade_culled)
    (num_rendered, rendered_image, rendered_normal, rendered_opacity, rendered_depth, rendered_feature,
     rendered_vfeature, weights, radii) = raw
    return {"render": rendered_image, "depth": rendered_depth, "geo_normal": rendered_normal, "opacity": rendered_opacity,
            "raw_feature": rendered_feature, "raw_vfeature": rendered_vfeature, "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0, "radii": radii, "num_rendered": num_rendered, "weights": weights,
            "raster_state": num_rendered._st if isinstance(num_rendered, raster.LazyCount) else None,
            "diffuse_light": vfeatures[:, 40:52] if is_training else None}


# training_step default: True = the resolve + loss tail runs as one fused CUDA kernel per direction
# (losses.fused_train_loss); False = the torch mirror of the reference's tail (render_view + image_loss).
FUSED_LOSS = True


def training_step(cam: ViewCamera, pc: SurfelModel, env_param: torch.Tensor, bg, gt_image, zero_grad=True,
                  fused_loss: Optional[bool] = None, overlap_bucket=None):
    """One stage-2 iteration of the hot path: shade + rasterise forward, loss, backward. Gradients
    are left in .grad of pc.trainable() and env_param. Returns (loss tensor, result dict).
    Both tails compute the same loss and gradients (tests/test_fused_loss_gpu.py).
    With `overlap_bucket` (dist.FlatGradBucket whose views are the .grad tensors) the per-surfel gradient
    all-reduce is issued segment by segment from inside the backward pass and has completed (on the current
    stream) when this returns."""
    if zero_grad:
        for t in pc.trainable() + [env_param]:
            t.grad = None
    if fused_loss is None:
        fused_loss = FUSED_LOSS
    if fused_loss:
        res = render_raw(cam, pc, (env_param, shading.MODE_LEARNABLE), bg, is_training=True)
        loss, terms = losses.fused_train_loss(res["render"], res["geo_normal"], res["opacity"], res["raw_vfeature"],
                                              gt_image, bg, lambda_pbr=1.0, lambda_normal=0.02, depth=res["depth"],
                                              cam_terms=camera_d2n_terms(cam))
        res["loss_terms"] = terms
    else:
        res = render_view(cam, pc, (env_param, shading.MODE_LEARNABLE), bg, is_training=True)
        loss = image_loss(res, gt_image, cam=cam)
    if overlap_bucket is not None:
        st = res.get("raster_state")
        if overlap_bucket.extra is not None and st is not None:
            # this rank's binning-overflow flag joins the last gradient segment: after the all-reduce every rank
            # sees "some rank overflowed" and takes the same re-capture decision (GraphedTrainingStep.finish)
            overlap_bucket.extra[0:1].copy_(st.t["num_rendered"][1:2])
        overlap_bucket.begin_overlap()
    loss.backward()
    if overlap_bucket is not None:
        overlap_bucket.finish_overlap()
    return loss, res


def reduce_segments(pc: SurfelModel) -> list:
    """Bucket segments for `pc.trainable() + [env]` in the order the stage-2 backward finishes them: the
    rasteriser's preprocess backward completes opacity / scaling / rotation / SH first (56 floats per surfel),
    the shading backward then completes xyz (it also receives the view-direction gradient), base colour,
    roughness, shading normals and the env map (31 floats per surfel + the map)."""
    return [[1, 2, 3, 4], [0, 5, 6, 7, 8]]


# GraphedTrainingStep default: True = the step is the fixed kernel sequence of fused_step.FusedTrainStep (no autograd
# between the kernels, one arena memset, binning / parameter backward on a side stream); False = `training_step`
# (the autograd mirror of the reference's control flow) is what gets captured.
FUSED_STEP = True
OVERLAP_BINNING = True    # render_view / training_step: bin on a side stream while the shading kernel runs
# SMs left out of the (persistent) shading grid for the binning kernels beside it. In the relight frame the shading kernel
# is 8x longer than the binning, so every reserved SM costs more than it gives (0 / 4 / 8 / 16 SMs: 1.410 / 1.417 / 1.443 /
# 1.492 ms per frame); the training step, where binning is the longer chain, reserves 20 (fused_step.FWD_RESERVE_SMS).
BINNING_RESERVE_SMS = 0


class BinOverflow(RuntimeError):
    """Raised by GraphedTrainingStep(on_overflow="raise") after a replay whose binning buffers were too small on some
    rank: the buffers have been grown and the graph re-captured; the caller re-runs its whole step."""


class GraphedTrainingStep:
    """One training iteration captured ONCE into a CUDA graph and replayed: one cudaGraphLaunch per iteration
    instead of ~150 host-side launches and tensor allocations, and no mid-step device->host read
    (the reference blocks on num_rendered inside every forward, rasterizer_impl.cu:311, and again at
    svgss.py:186). The binning buffers have a fixed capacity (2x the largest num_rendered seen), the
    (num_rendered, overflow) pair lands in pinned host memory, and __call__ checks it after the replay -- on
    overflow the capacity is raised, the graph re-captured and the SAME step re-run, so the results are always
    those of a complete render. Per-step inputs (camera matrices, ground-truth image) are copied into static
    device buffers; results and .grad tensors are static buffers overwritten by the next call.

    What is captured: `fused_step.FusedTrainStep` (default, `fused=True`: a fixed sequence of svgir kernels, see
    that module) or `training_step` (`fused=False`: the autograd path, ~55 extra torch kernels per step).

    The camera intrinsics (H, W, tan_fov) are baked into the graph; a different value re-captures.
    With `bucket` (dist.FlatGradBucket) the .grad views are zeroed inside the graph and accumulated
    in place, ready for bucket.all_reduce() after the call.
    """

    def __init__(self, pc: SurfelModel, env_param: torch.Tensor, bg: torch.Tensor, cam: ViewCamera,
                 gt_image: torch.Tensor, bucket=None, warmup: int = 2, reduce_in_graph: bool = False,
                 zero_in_graph: bool = True, fused: Optional[bool] = None, radiance_cache=None,
                 lambda_radiance: float = 0.05, on_overflow: str = "rerun"):
        self.pc, self.env, self.bg, self.bucket = pc, env_param, bg, bucket
        # radiance_cache (fused path only): the step also carries lambda_radiance * get_radiance_loss (svgss.py:319-320)
        self.radiance_cache, self.lambda_radiance = radiance_cache, lambda_radiance
        # zero_in_graph=False (needs `bucket`): the graph ACCUMULATES into the bucket instead of zeroing it first, so a
        # rank that renders several views per step replays it once per view and reduces once (C4: 8 views / step);
        # the caller zeroes the bucket at the start of the step. A (re-)capture leaves the bucket's content untouched,
        # and a replay whose binning overflowed adds exactly zero (every backward kernel returns at once when the
        # overflow flag is set), so re-running the step after the re-capture gives the complete sum.
        # on_overflow="raise": after an overflowed replay the step is NOT re-run here; BinOverflow is raised (on every rank
        # when the reduction is in the graph: the flag is summed with the gradients) once the bins are grown and the graph
        # re-captured. That is what makes accumulate + reduce_in_graph usable: the LAST local view of a multi-view step
        # carries the exchange under its shading backward; after the in-place all-reduce the bucket holds all ranks' sums,
        # so a re-run of that one view would count them twice -- the caller zeroes the bucket and redoes the step instead.
        self.zero_in_graph = bool(zero_in_graph)
        if on_overflow not in ("rerun", "raise"):
            raise ValueError("on_overflow: 'rerun' or 'raise'")
        self.on_overflow = on_overflow
        if not self.zero_in_graph and (bucket is None or (reduce_in_graph and on_overflow != "raise")):
            raise ValueError("zero_in_graph=False needs a bucket and a reduction outside the graph (or on_overflow='raise')")
        # reduce_in_graph: the segment-wise all-reduce of `bucket` is captured INSIDE the graph, overlapped
        # with the shading backward; the caller must not all-reduce again.
        self.reduce_in_graph = bool(reduce_in_graph and bucket is not None)
        if self.reduce_in_graph and bucket.extra is None:
            raise ValueError("reduce_in_graph needs FlatGradBucket(..., extra_floats>=1) for the overflow flag")
        self.fused = FUSED_STEP if fused is None else bool(fused)
        if self.fused and self.reduce_in_graph and getattr(bucket, "segment_peer", None) is None:
            self.fused = False   # NCCL collectives inside the step are issued through the autograd hooks
        if radiance_cache is not None and not self.fused:
            raise ValueError("radiance_cache is carried by the fused step only")
        self.flag_host = raster.pinned_forever((1,), torch.float32) if self.reduce_in_graph else None
        dev = pc.xyz.device
        self.dev = dev
        self.cam = blocked_camera(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                                  cam.world_view_transform, cam.full_proj_transform, cam.camera_center,
                                  cam.patch_bbox, cam.prcppoint, device=dev)
        self.gt = gt_image.clone()
        self.graph = None
        self.loss = None
        self.res = None
        self.fs = None
        self.captures = 0
        self.launches_per_step = 0
        self.warmup = warmup
        self._copy_stream = None
        self._staged = self._consumed = None

    def _params(self):
        return self.pc.trainable() + [self.env]

    def _zero(self):
        if self.bucket is not None:
            self.bucket.zero()
        else:
            for t in self._params():
                t.grad = None

    # ---- fused path ---------------------------------------------------------------------------------------------
    def _capture_fused(self):
        from . import fused_step
        cur = torch.cuda.current_stream(self.dev)
        keep = None if self.zero_in_graph else self.bucket.flat.clone()   # accumulated gradients survive the capture
        if self.fs is None:
            self.fs = fused_step.FusedTrainStep(self.pc, self.env, self.bg, self.cam, self.gt, bucket=self.bucket,
                                                zero_grads=self.zero_in_graph, reduce_in_step=self.reduce_in_graph,
                                                radiance_cache=self.radiance_cache, lambda_radiance=self.lambda_radiance)
            if self.bucket is None:
                self.bucket_private = self.fs.bucket
        fs = self.fs
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            fs.calibrate()                      # sizes the bins for this camera (one host sync)
            for _ in range(max(self.warmup, 1)):  # eager: lazy initialisations (function attributes, streams)
                fs.enqueue()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = fs.enqueue()
        if self.flag_host is not None:
            self.flag_host = fs.flag_host
        self.res = fs.result
        self.launches_per_step = fs.launches
        self.captures += 1
        if keep is not None:
            self.bucket.flat.copy_(keep)

    def _finish_fused(self) -> int:
        fs = self.fs
        for _ in range(4):
            torch.cuda.current_stream(self.dev).synchronize()
            R, overflow = fs.read_count()
            if not overflow and not (self.reduce_in_graph and float(self.flag_host[0]) > 0):
                return R
            # this rank or (flag summed over ranks) another one overflowed: every rank re-captures and re-runs
            # together -- the step's all-reduce lives in the graph, so the replays must stay matched across ranks
            if overflow:
                fs.grow(R)
            self.graph = None
            self._capture_fused()
            if self.on_overflow == "raise":
                raise BinOverflow("binning buffers overflowed; re-captured with larger ones -- redo the step")
            self.graph.replay()
        raise RuntimeError("GraphedTrainingStep: binning capacity did not converge")

    # ---- autograd path ------------------------------------------------------------------------------------------
    def _capture(self):
        if self.fused:
            return self._capture_fused()
        cur = torch.cuda.current_stream(self.dev)
        keep = None if self.zero_in_graph else self.bucket.flat.clone()   # accumulated gradients survive the capture
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(self.warmup):  # eager: sizes the binning hint, warms allocators / lazy inits
                self._zero()
                with raster.count_mode("speculative"):
                    _, r = training_step(self.cam, self.pc, self.env, self.bg, self.gt, zero_grad=False,
                                         overlap_bucket=self.bucket if self.reduce_in_graph else None)
            R = int(r["num_rendered"])
            raster.reserve(self.dev, self.pc.xyz.shape[0], self.cam.image_width, self.cam.image_height,
                           int(R * raster.ASYNC_SLACK) + raster.ASYNC_MARGIN)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self._zero()
        raster.prepare_capture(1)
        self.graph = torch.cuda.CUDAGraph()
        n0 = launch_count()
        with torch.cuda.graph(self.graph):
            if self.bucket is not None and self.zero_in_graph:
                self.bucket.zero()
            with raster.count_mode("async", owner_resolves=True):
                self.loss, self.res = training_step(self.cam, self.pc, self.env, self.bg, self.gt, zero_grad=False,
                                                    overlap_bucket=self.bucket if self.reduce_in_graph else None)
            if self.reduce_in_graph:
                self.flag_host.copy_(self.bucket.extra[0:1], non_blocking=True)
        self.launches_per_step = launch_count() - n0  # svgir kernels inside the graph
        self.captures += 1
        if keep is not None:
            self.bucket.flat.copy_(keep)

    def _check_intrinsics(self, cam: ViewCamera):
        c = self.cam
        if (cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy) != (c.image_height, c.image_width, c.tanfovx, c.tanfovy):
            self.cam = ViewCamera(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                                  c.world_view_transform, c.full_proj_transform, c.camera_center, c.patch_bbox,
                                  c.prcppoint, block=c.block)
            self.graph = None
            self.fs = None   # intrinsics are baked into the fused step's kernel arguments
        return self.cam

    def load_inputs(self, cam: ViewCamera, gt_image: torch.Tensor):
        c = self._check_intrinsics(cam)
        if self.fs is not None:
            self.fs.refresh_taps()   # incident directions re-sampled in place since the capture? (a version compare)
        if cam.block is not None and c.block is not None:
            c.block.copy_(cam.block, non_blocking=True)   # one copy for all per-view constants
            if gt_image is not self.gt:
                self.gt.copy_(gt_image, non_blocking=True)
            return
        c.world_view_transform.copy_(cam.world_view_transform, non_blocking=True)
        c.full_proj_transform.copy_(cam.full_proj_transform, non_blocking=True)
        c.camera_center.copy_(cam.camera_center, non_blocking=True)
        c.patch_bbox.copy_(cam.patch_bbox, non_blocking=True)
        c.prcppoint.copy_(cam.prcppoint, non_blocking=True)
        if gt_image is not self.gt:
            self.gt.copy_(gt_image, non_blocking=True)

    def __call__(self, cam: ViewCamera, gt_image: torch.Tensor, check: bool = True):
        """Runs one step. With check=True (default) waits for the step and validates the binning
        capacity (re-running on overflow); check=False returns right after the launch and leaves the
        validation to a later `finish()`."""
        self.load_inputs(cam, gt_image)
        if self.graph is None:
            self._capture()
        self.graph.replay()
        if check:
            self.finish()
        return self.loss, self.res

    # ---- host inputs one step ahead -----------------------------------------------------------------------------
    def prefetch(self, cam: ViewCamera, gt_image: torch.Tensor):
        """Uploads the NEXT step's inputs (camera block + ground-truth image, typically pinned host memory) into
        a staging slot on a copy stream, so the host->device transfer overlaps the step that is running.
        `replay_prefetched()` then moves the slot into the graph's static inputs (a device-to-device copy
        at the head of the step) and replays."""
        if cam.block is None:
            raise ValueError("prefetch needs a blocked camera (pipeline.blocked_camera)")
        self._check_intrinsics(cam)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.dev)
            self._stage_cam = torch.empty_like(self.cam.block)
            self._stage_gt = torch.empty_like(self.gt)
        cs = self._copy_stream
        if self._consumed is not None:
            cs.wait_event(self._consumed)   # the slot's previous content has been moved into the static inputs
        with torch.cuda.stream(cs):
            self._stage_cam.copy_(cam.block, non_blocking=True)
            self._stage_gt.copy_(gt_image, non_blocking=True)
            self._staged = torch.cuda.Event()
            self._staged.record(cs)

    def replay_prefetched(self, check: bool = False):
        if self._staged is None:
            raise RuntimeError("replay_prefetched: call prefetch() first")
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self._staged)
        self.cam.block.copy_(self._stage_cam, non_blocking=True)
        self.gt.copy_(self._stage_gt, non_blocking=True)
        self._consumed = torch.cuda.Event()
        self._consumed.record(cur)
        self._staged = None
        if self.fs is not None:
            self.fs.refresh_taps()
        if self.graph is None:
            self._capture()
        self.graph.replay()
        if check:
            self.finish()
        return self.loss, self.res

    def finish(self) -> int:
        """Waits for the last replay and returns its num_rendered; on overflow re-captures with larger
        bins and re-runs the step."""
        if self.fused:
            return self._finish_fused()
        st = self.res["raster_state"]
        for _ in range(4):
            torch.cuda.current_stream(self.dev).synchronize()
            st.pending = True  # the graph re-wrote count_host
            try:
                R = st.resolve()   # raises the capacity hint on a local overflow
                if not (self.reduce_in_graph and float(self.flag_host[0]) > 0):
                    return R
                # another rank overflowed: every rank re-captures and re-runs together (the step's
                # all-reduce lives in the graph, so the replays must stay matched across ranks)
            except raster.CapacityOverflow:
                pass
            self.graph = None
            self._capture()
            self.graph.replay()
            st = self.res["raster_state"]
        raise RuntimeError("GraphedTrainingStep: binning capacity did not converge")


class GraphedRelightFrame:
    """One relighting frame (eval_relighting_tensoIR.py:303-331: `render_view(..., is_training=False)` under a fixed
    HDR env map) captured ONCE into a CUDA graph and replayed per (view, env map): the eager frame issues ~70 launches
    and blocks on `num_rendered`, which costs as much host time as the kernels take. Same capacity protocol as
    GraphedTrainingStep: fixed binning capacity, (num_rendered, overflow) lands in pinned memory, `finish()` checks it
    after the replay and re-captures + re-runs on overflow. Per-frame inputs (camera block, env map) are copied into
    static device buffers; the result dict holds static tensors overwritten by the next frame."""

    def __init__(self, pc: SurfelModel, env_map: torch.Tensor, bg: torch.Tensor, cam: ViewCamera, warmup: int = 2,
                 env_mode=None):
        self.pc, self.bg = pc, bg
        dev = pc.xyz.device
        self.dev = dev
        self.env_mode = shading.MODE_FIXED if env_mode is None else env_mode
        self.env = env_map.detach().clone()
        self.cam = blocked_camera(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                                  cam.world_view_transform, cam.full_proj_transform, cam.camera_center,
                                  cam.patch_bbox, cam.prcppoint, device=dev)
        self.graph = None
        self.res = None
        self.captures = 0
        self.launches_per_frame = 0
        self.warmup = warmup

    def _frame(self):
        with torch.no_grad():
            return render_view(self.cam, self.pc, (self.env, self.env_mode), self.bg, is_training=False)

    def _capture(self):
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                with raster.count_mode("speculative"):
                    r = self._frame()
            R = int(r["num_rendered"])
            raster.reserve(self.dev, self.pc.xyz.shape[0], self.cam.image_width, self.cam.image_height,
                           int(R * raster.ASYNC_SLACK) + raster.ASYNC_MARGIN)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        raster.prepare_capture(1)
        self.graph = torch.cuda.CUDAGraph()
        n0 = launch_count()
        with torch.cuda.graph(self.graph):
            with raster.count_mode("async", owner_resolves=True):
                self.res = self._frame()
        self.launches_per_frame = launch_count() - n0
        self.captures += 1

    def load_inputs(self, cam: ViewCamera, env_map: Optional[torch.Tensor]):
        c = self.cam
        if (cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy) != (c.image_height, c.image_width, c.tanfovx, c.tanfovy):
            self.cam = ViewCamera(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, c.world_view_transform,
                                  c.full_proj_transform, c.camera_center, c.patch_bbox, c.prcppoint, block=c.block)
            c = self.cam
            self.graph = None
        if cam.block is not None:
            c.block.copy_(cam.block, non_blocking=True)
        else:
            for k in ("world_view_transform", "full_proj_transform", "camera_center", "patch_bbox", "prcppoint"):
                getattr(c, k).copy_(getattr(cam, k), non_blocking=True)
        if env_map is not None and env_map is not self.env:
            self.env.copy_(env_map, non_blocking=True)
        # the captured shading kernel points at the cached env taps of pc.incident_dirs: recomputed in place if those changed
        shading.refresh_env_taps(self.pc.incident_dirs, int(self.env.shape[-3]), int(self.env.shape[-2]))

    def __call__(self, cam: ViewCamera, env_map: Optional[torch.Tensor] = None, check: bool = True) -> dict:
        self.load_inputs(cam, env_map)
        if self.graph is None:
            self._capture()
        self.graph.replay()
        if check:
            self.finish()
        return self.res

    def finish(self) -> int:
        st = self.res["raster_state"]
        for _ in range(4):
            torch.cuda.current_stream(self.dev).synchronize()
            st.pending = True  # the graph re-wrote count_host
            try:
                return st.resolve()
            except raster.CapacityOverflow:
                pass
            self.graph = None
            self._capture()
            self.graph.replay()
            st = self.res["raster_state"]
        raise RuntimeError("GraphedRelightFrame: binning capacity did not converge")


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
    n_launch = launch_count()
    # the same step as the bench runs it: FusedTrainStep captured into a CUDA graph, replayed twice
    want = [t.grad.detach().clone() for t in pc.trainable() + [env]]
    runner = GraphedTrainingStep(pc, env, bg, cam, gt)
    for _ in range(2):
        gloss, gres = runner(cam, gt)
    torch.cuda.synchronize()
    assert abs(float(gloss) - float(loss)) <= 1e-5 * abs(float(loss)), (float(gloss), float(loss))
    for t, w in zip(pc.trainable() + [env], want):
        assert float((t.grad - w).norm()) <= 1e-3 * float(w.norm()) + 1e-12
    return {"loss": float(loss), "launches": n_launch, "num_rendered": res["num_rendered"],
            "graph_launches_per_step": runner.launches_per_step}
