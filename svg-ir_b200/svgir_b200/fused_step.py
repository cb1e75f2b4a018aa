"""One stage-2 training iteration issued as a FIXED sequence of C-ABI calls (no autograd, no allocations).

`pipeline.training_step` mirrors the reference's control flow -- `render_view` (gaussian_renderer/svgss.py:15-262)
through torch.autograd.Functions, `loss.backward()` -- which is what a drop-in caller needs, but between the svgir
kernels autograd launches ~55 small torch kernels per step (zero-fills of fresh gradient tensors, the view-direction
normalise and its backward, gradient accumulation adds): ~7 % of a 2.2 ms step, and every kernel runs serialised on
one stream. `FusedTrainStep` is the same iteration for a caller that owns the whole step (bench.py, a trainer):

  * every buffer is allocated once; the per-step accumulators live in ONE arena cleared by ONE memset, the parameter
    gradients in one flat buffer (dist.FlatGradBucket: `.grad` of each parameter is a view of it) cleared by another;
  * the view direction normalize(camera_center - xyz) (svgss.py:95) is evaluated inside the shading kernels and its
    gradient is added to xyz.grad by the shading backward (svgir_shade_in.means3D / campos);
  * the kernels ADD parameter gradients into the flat buffer for the visible surfels only
    (svgir_raster_backward_params, SVGIR_SHADE_ACCUMULATE) -- no zero-filling of culled rows, and a rank that renders
    several views per step accumulates them without extra kernels;
  * independent work runs concurrently: tile binning (scan, duplicateWithKeys, sort) on a side stream under the
    shading forward; the rasteriser's parameter backward (+ the all-reduce of its gradient segment over NVLink peer
    memory, N > 1) on the side stream under the shading backward. Forks and joins are events, so the whole step
    captures into one CUDA graph (pipeline.GraphedTrainingStep).

Same arithmetic as the autograd path (pipeline.FUSED_VIEWDIRS: it evaluates the view direction inside the shading
kernel too): identical kernels on identical inputs, bit-identical images; gradients differ by the order of the atomic
additions only. tests/test_fused_step_gpu.py compares loss, images and every gradient with `pipeline.training_step`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib, losses, raster, shading
from . import dist as svdist


# The shading forward is a persistent grid that takes every SM's registers, so the tile binning issued beside it on the
# side stream only gets the slots its first kernel grabbed before the grid filled up and then waits for CTAs to retire;
# binning is the longer of the two chains (0.27 vs 0.20 ms). Leaving 20 of the 148 SMs out of the shading grid
# (svgir_shade_reserve_sms) shortens the phase: step 2.068 -> 2.029 ms; 8 / 16 / 24 / 40 SMs: 2.038 / 2.031 / 2.032 / 2.059
# (stream priorities made no difference).
FWD_RESERVE_SMS = 20


def _al4(n: int) -> int:
    return (int(n) + 3) // 4 * 4


class StepResult(dict):
    """Result dict of a fused step; `visibility_filter` / `num_rendered` are produced on first access."""

    def __init__(self, step: "FusedTrainStep", eager: dict):
        super().__init__(eager)
        self._step = step

    def __missing__(self, key):
        if key == "visibility_filter":
            return self["radii"] > 0
        if key == "num_rendered":
            return self._step.read_count()[0]
        raise KeyError(key)

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


class FusedTrainStep:
    S, VS = 4, 52   # training G-buffer: features [vis 1, local light 3], vfeatures [pbr 12, base 12, normal 12, rough 4, diffuse 12]

    def __init__(self, pc, env_param: torch.Tensor, bg: torch.Tensor, cam, gt_image: torch.Tensor, bucket=None,
                 lambda_pbr: float = 1.0, lambda_normal: float = 0.02, zero_grads: bool = True,
                 reduce_in_step: bool = False, capacity: Optional[int] = None, surface_term: str = "depth2normal",
                 image_mask: Optional[torch.Tensor] = None, radiance_cache=None, lambda_radiance: float = 0.05):
        """pc: pipeline.SurfelModel; cam: pipeline.ViewCamera whose tensors are the step's STATIC camera inputs (copy a
        new view into cam.block before each step); gt_image: the static ground-truth buffer [3,H,W].
        bucket: dist.FlatGradBucket over pc.trainable() + [env_param] (default: a private one, so that all parameter
        gradients are one memset); zero_grads=False leaves clearing it to the caller (multi-view accumulation).
        reduce_in_step (needs a bucket built on dist.PeerAllReduce): the gradient all-reduce over NVLink peer memory
        is part of the step -- segment 0 on the side stream under the shading backward, the rest after it.
        radiance_cache (radiance.RadianceCache after update()): adds lambda_radiance * get_radiance_loss
        (gaussian_renderer/svgss.py:319-320, arguments/__init__.py:130) to the step: its kernels run on the side stream
        under the rasteriser and add their albedo / roughness / env gradients before the shading backward adds its own."""
        self.L = _lib.lib()
        shading._L()
        losses._bind()
        if not pc.xyz.is_cuda:
            raise RuntimeError("FusedTrainStep needs CUDA tensors (there is no CPU fallback)")
        self.pc, self.env, self.bg, self.cam, self.gt = pc, env_param, bg.contiguous(), cam, gt_image
        dev = pc.xyz.device
        self.dev = dev
        P = int(pc.xyz.shape[0])
        H, W = int(cam.image_height), int(cam.image_width)
        self.P, self.H, self.W = P, H, W
        M = int(pc.shs.shape[1])
        S, VS = self.S, self.VS
        NV = VS // 4
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        params = pc.trainable() + [env_param]
        for t in params + [pc.radiance, pc.visibility, pc.incident_dirs, pc.incident_areas]:
            if t.dtype != torch.float32 or not t.is_contiguous() or t.data_ptr() % 16:
                raise RuntimeError("FusedTrainStep: parameters and light buffers must be contiguous, 16-byte aligned fp32")
        self.params = params
        self.own_bucket = bucket is None
        self.bucket = bucket if bucket is not None else svdist.FlatGradBucket(params)
        if [id(p) for p in self.bucket.params] != [id(p) for p in params]:
            raise ValueError("FusedTrainStep: the bucket must be built over pc.trainable() + [env_param], in that order")
        self.zero_grads = bool(zero_grads)
        self.reduce_in_step = bool(reduce_in_step)
        self.peer = getattr(self.bucket, "segment_peer", None)
        if self.reduce_in_step and (self.peer is None or self.bucket.extra is None):
            raise ValueError("reduce_in_step needs FlatGradBucket(..., segment_peer=PeerAllReduce, extra_floats>=1)")
        if self.reduce_in_step and len(self.bucket.seg_bounds) > _lib.PEER_BANKS:
            raise ValueError("at most %d gradient segments" % _lib.PEER_BANKS)
        self.side = torch.cuda.Stream(dev)
        self.fwd_reserve = int(FWD_RESERVE_SMS)

        # ---- per-step accumulators: one arena, one memset ------------------------------------------------------
        sizes = [("geo", P * _lib.GEO_GRAD_FLOATS), ("dfeat", P * S), ("dvfeat", P * VS), ("weights", P), ("dmeans2D", P * 3)]
        off, total = {}, 0
        for k, n in sizes:
            off[k] = total
            total += _al4(n)
        self.arena = torch.zeros(max(total, 4), **f32)
        a = self.arena
        self.geo = a[off["geo"]:off["geo"] + P * _lib.GEO_GRAD_FLOATS].view(P, _lib.GEO_GRAD_FLOATS)
        self.dfeat = a[off["dfeat"]:off["dfeat"] + P * S].view(P, S)
        self.dvfeat = a[off["dvfeat"]:off["dvfeat"] + P * VS].view(P, VS)
        self.weights = a[off["weights"]:off["weights"] + P].view(P, 1)
        self.dmeans2D = a[off["dmeans2D"]:off["dmeans2D"] + P * 3].view(P, 3)

        # ---- static buffers ---------------------------------------------------------------------------------------
        gx, gy = (W + 15) // 16, (H + 15) // 16
        T = gx * gy
        t = self.t = {}
        t["rec"] = torch.empty((P, _lib.REC_FLOATS), **f32)
        t["cov3D"] = torch.empty((P, 6), **f32)
        t["clamped"] = torch.zeros((P,), dtype=torch.uint8, device=dev)
        t["rect"] = torch.empty((P, 4), dtype=torch.int16, device=dev)
        t["tiles_touched"] = torch.empty((P,), **i32)
        t["tile_count"] = torch.empty((T,), **i32)
        t["tile_cursor"] = torch.empty((T,), **i32)
        t["ranges"] = torch.empty((T, 2), **i32)
        t["big_tiles"] = torch.empty((3 * T + 4,), **i32)
        t["num_rendered"] = torch.zeros((2,), **i32)
        t["final_T"] = torch.empty((H * W,), **f32)
        t["final_D"] = torch.empty((H * W,), **f32)
        t["n_contrib"] = torch.empty((H * W,), **i32)
        t["vis_list"] = torch.empty((max(P, 1),), **i32)
        t["vis_count"] = torch.zeros((1,), **i32)
        self.radii = torch.zeros((P,), **i32)
        # images: forward outputs and their gradients (the depth / flat-feature gradients of this loss are zero)
        self.img = {"color": torch.zeros((3, H, W), **f32), "normal": torch.zeros((3, H, W), **f32),
                    "depth": torch.zeros((1, H, W), **f32), "opacity": torch.zeros((1, H, W), **f32),
                    "feature": torch.zeros((S, H, W), **f32), "vfeature": torch.zeros((NV, H, W), **f32)}
        self.gimg = {"color": torch.empty((3, H, W), **f32), "normal": torch.empty((3, H, W), **f32),
                     "depth": torch.zeros((1, H, W), **f32), "opacity": torch.empty((1, H, W), **f32),  # depth: surface term
                     "feature": torch.zeros((S, H, W), **f32), "vfeature": torch.empty((NV, H, W), **f32)}
        self.feats = torch.zeros((P, S), **f32)
        self.vfeats = torch.zeros((P, VS), **f32)
        He, We = int(env_param.shape[-3]), int(env_param.shape[-2])
        self.env_hw = (He, We)
        self.env_act = torch.empty((He, We, 3), **f32)
        self.env_scratch = torch.empty((shading.ENV_COPIES, He, We, 4), **f32)
        self.sums = torch.empty((P, 12), **f32)
        self.loss_out = torch.zeros(8, **f32)
        nblk = int(self.L.svgir_train_loss_blocks(W, H))
        self.loss_partials = torch.empty(4 * nblk, **f32)
        self.loss_counter = torch.zeros(1, **i32)
        # landing buffers of copies that are captured into CUDA graphs: never recycled (raster.pinned_forever)
        self.count_host = raster.pinned_forever((2,), torch.int32)
        self.flag_host = raster.pinned_forever((1,), torch.float32) if self.reduce_in_step else None

        # ---- C structs (pointers are static, so they are built once) ------------------------------------------------
        from .pipeline import _config_tensor
        self.keep = []
        s = raster.RasterSettings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=self.bg,
                                  scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
                                  sh_degree=pc.active_sh_degree, campos=cam.camera_center, prefiltered=False, debug=False,
                                  variant=_lib.VARIANT_SVGSS, patch_bbox=cam.patch_bbox,
                                  config=_config_tensor(pc.config, dev))
        self.cfg = raster._make_cfg(s, P, S, VS, M, dev, self.keep)
        for name in ("viewmatrix", "projmatrix", "campos", "patch_bbox"):
            src = {"viewmatrix": cam.world_view_transform, "projmatrix": cam.full_proj_transform,
                   "campos": cam.camera_center, "patch_bbox": cam.patch_bbox}[name]
            if getattr(self.cfg, name) != src.data_ptr():
                raise RuntimeError("FusedTrainStep: camera tensor `%s` must be contiguous, 16-byte aligned fp32 on the device "
                                   "(use pipeline.blocked_camera)" % name)
        cin = self.cin = _lib.RasterIn()
        cin.means3D, cin.opacities, cin.scales, cin.rotations = (pc.xyz.data_ptr(), pc.opacity.data_ptr(),
                                                                pc.scaling.data_ptr(), pc.rotation.data_ptr())
        cin.cov3D_precomp, cin.colors_precomp = None, None
        cin.shs = pc.shs.data_ptr()
        cin.features, cin.vfeatures = self.feats.data_ptr(), self.vfeats.data_ptr()
        cst = self.cst = _lib.RasterState()
        for k in ("rec", "cov3D", "clamped", "rect", "tiles_touched", "tile_count", "tile_cursor", "ranges", "big_tiles",
                  "num_rendered", "final_T", "final_D", "n_contrib", "vis_list", "vis_count"):
            setattr(cst, k, t[k].data_ptr())
        cst.sorted_keys = None
        cout = self.cout = _lib.RasterOut()
        for k in ("color", "normal", "depth", "opacity", "feature", "vfeature"):
            setattr(cout, k, self.img[k].data_ptr())
        cout.weights, cout.radii = self.weights.data_ptr(), self.radii.data_ptr()
        g = self.rgrads = _lib.RasterGrads()
        g.dL_dcolor, g.dL_dnormal, g.dL_ddepth = (self.gimg["color"].data_ptr(), self.gimg["normal"].data_ptr(),
                                                  self.gimg["depth"].data_ptr())
        g.dL_dopacity, g.dL_dfeature, g.dL_dvfeature = (self.gimg["opacity"].data_ptr(), self.gimg["feature"].data_ptr(),
                                                        self.gimg["vfeature"].data_ptr())
        g.geo_grad, g.dL_dfeatures, g.dL_dvfeatures = self.geo.data_ptr(), self.dfeat.data_ptr(), self.dvfeat.data_ptr()

        # shading: the packed layout of pipeline / shading._ShadePackedFn (training branch)
        Ns = int(pc.incident_dirs.shape[1])
        self.scfg_f = shading.ShadeCfg(P, Ns, He, We, shading.MODE_LEARNABLE, 0, shading.SHADE_VIEW_4X4, 0)
        self.scfg_b = shading.ShadeCfg(P, Ns, He, We, shading.MODE_LEARNABLE, 0,
                                       shading.SHADE_ENV_READY | shading.SHADE_ACCUMULATE | shading.SHADE_VIEW_4X4, 0)
        env3 = env_param if env_param.dim() == 3 else env_param[0]
        self.sin = shading.ShadeIn(
            pc.base_color.data_ptr(), pc.roughness.data_ptr(), None, pc.shading_normal.data_ptr(), None,
            pc.radiance.data_ptr(), pc.visibility.data_ptr(), pc.incident_dirs.data_ptr(), pc.incident_areas.data_ptr(),
            env3.data_ptr(), None, self.env_act.data_ptr(), cam.world_view_transform.data_ptr(), t["vis_list"].data_ptr(),
            t["vis_count"].data_ptr(), pc.xyz.data_ptr(), cam.camera_center.data_ptr(), t["num_rendered"][1:].data_ptr())
        # env taps of the (fixed) incident directions: no acos / atan2 per sample in the shading kernels
        self.env_taps = shading.refresh_env_taps(pc.incident_dirs, He, We)
        self.sin.env_taps = None if self.env_taps is None else self.env_taps.data_ptr()
        vp, fp = self.vfeats.data_ptr(), self.feats.data_ptr()
        self.sout = shading.ShadeOut(vp, vp + 4 * 40, None, None, None, fp, fp + 4, None, None, vp + 4 * 12,
                                     self.sums.data_ptr(), None, VS, S, S, 0)
        self._bind_grads()

        # image loss: L1 terms + the reference's surface term cos_loss(normal, depth2normal(depth)) (svgss.py:280-313)
        from .pipeline import camera_d2n_terms
        ifx, ify, cx, cy = camera_d2n_terms(cam)
        mode = losses.NORMAL_D2N if surface_term == "depth2normal" else losses.NORMAL_GEO
        self.image_mask = image_mask.to(**f32).contiguous() if image_mask is not None else None
        self.lcfg = losses.TrainLossCfg(W, H, 0, NV, 0, 6, float(lambda_pbr), float(lambda_normal), self.bg.data_ptr(),
                                        mode, ifx, ify, cx, cy, 0)
        self.lin = losses.TrainLossIn(self.img["color"].data_ptr(), self.img["normal"].data_ptr(), self.img["opacity"].data_ptr(),
                                      self.img["vfeature"].data_ptr(), self.gt.data_ptr(), self.img["depth"].data_ptr(),
                                      self.image_mask.data_ptr() if self.image_mask is not None else None)
        self.lgr = losses.TrainLossGrads(self.gimg["color"].data_ptr(), self.gimg["normal"].data_ptr(),
                                         self.gimg["depth"].data_ptr() if mode == losses.NORMAL_D2N else None,
                                         self.gimg["opacity"].data_ptr(), None, self.gimg["vfeature"].data_ptr())
        # radiance-consistency term (optional)
        self.rc = radiance_cache
        self.lambda_radiance = float(lambda_radiance)
        if self.rc is not None:
            from . import radiance as _rad
            _rad._L()
            rc = self.rc
            S2 = int(rc.incident_dirs.shape[1])
            hit = rc.hemi_index_buffers.reshape(P, S2)
            if hit.dtype != torch.int32 or not hit.is_contiguous():
                raise RuntimeError("FusedTrainStep: radiance_cache.hemi_index_buffers must be contiguous int32")
            for tns in (rc.incident_dirs, rc.incident_areas, rc.visibility_tracing, rc.uv_buffers, rc.radiances, rc.geo_normal):
                if tns.dtype != torch.float32 or not tns.is_contiguous() or tns.shape[0] != P:
                    raise RuntimeError("FusedTrainStep: radiance_cache tensors must be contiguous fp32 over the same surfels")
            self.rad = {"irr": torch.empty((P, 3), **f32), "saved": torch.empty((P, 8), **f32),
                        "scratch": torch.empty((_rad.SCRATCH_FLOATS,), **f32), "env_act": torch.empty((He, We, 3), **f32),
                        "env_scratch": torch.empty((_rad.ENV_COPIES, He, We, 4), **f32), "loss": torch.zeros(1, **f32),
                        "grad": torch.full((1,), self.lambda_radiance, **f32),
                        "ratio": rc.radiance_ratio.detach().reshape(1).to(**f32).contiguous(), "hit": hit}
            self.rcfg = _rad.RadianceLossCfg(P, S2, He, We, shading.MODE_LEARNABLE, _rad.RADIANCE_NORMALS_VERTEX_MAJOR,
                                             int(pc.roughness.shape[1]), 0)   # pc.shading_normal is [P,4,3]
            self.rin = _rad.RadianceLossIn(
                pc.xyz.data_ptr(), cam.camera_center.data_ptr(), rc.geo_normal.data_ptr(), rc.incident_dirs.data_ptr(),
                rc.incident_areas.data_ptr(), rc.visibility_tracing.data_ptr(), hit.data_ptr(), rc.uv_buffers.data_ptr(),
                rc.radiances.data_ptr(), self.rad["ratio"].data_ptr(), pc.shading_normal.data_ptr(), pc.base_color.data_ptr(),
                pc.roughness.data_ptr(), env3.data_ptr(), self.rad["env_act"].data_ptr(), t["num_rendered"][1:].data_ptr(), None)
            self.rad["taps"] = shading.refresh_env_taps(rc.incident_dirs, He, We)
            self.rin.env_taps = None if self.rad["taps"] is None else self.rad["taps"].data_ptr()
        self.cap = 0
        self._alloc_bins(capacity if capacity else raster._CAP_HINT.get((dev.index, P, W, H), 0))
        self.launches = 0
        self.result = StepResult(self, {
            "render": self.img["color"], "depth": self.img["depth"], "geo_normal": self.img["normal"],
            "opacity": self.img["opacity"], "raw_feature": self.img["feature"], "raw_vfeature": self.img["vfeature"],
            "radii": self.radii, "weights": self.weights, "viewspace_grad": self.dmeans2D,
            "diffuse_light": self.vfeats[:, 40:52], "loss_terms": self.loss_out,
            "loss_radiance": self.rad["loss"] if self.rc is not None else None})

    # ---------------------------------------------------------------------------------------------------------------
    def _bind_grads(self):
        """(Re)reads the parameter-gradient pointers from the bucket views."""
        b = self.bucket
        if not b.attached():
            b.attach()
        gv = [b.view(i) for i in range(len(self.params))]   # xyz, opacity, scaling, rotation, shs, base, rough, normal, env
        self.gviews = gv
        self.pg = _lib.ParamGrads(gv[0].data_ptr(), gv[1].data_ptr(), gv[2].data_ptr(), gv[3].data_ptr(), gv[4].data_ptr(),
                                  self.dmeans2D.data_ptr())
        vg, fg = self.dvfeat.data_ptr(), self.dfeat.data_ptr()
        VS, S = self.VS, self.S
        self.sgr = shading.ShadeGrads(
            vg, vg + 4 * 40, None, None, None, fg, fg + 4, None, None,
            gv[5].data_ptr(), gv[6].data_ptr(), None, gv[7].data_ptr(), None, None, None, gv[8].data_ptr(),
            vg + 4 * 12, self.sums.data_ptr(), None, self.env_scratch.data_ptr(), VS, S, S, 0, gv[0].data_ptr())

    def _alloc_bins(self, cap: int):
        cap = max(int(cap), 1)
        self.cap = cap
        self.t["keys"] = torch.empty((cap,), dtype=torch.int64, device=self.dev)
        self.t["point_list"] = torch.empty((cap,), dtype=torch.int32, device=self.dev)
        self.cst.keys, self.cst.point_list, self.cst.cap_R = self.t["keys"].data_ptr(), self.t["point_list"].data_ptr(), cap

    def refresh_taps(self):
        """Call after the incident directions were re-sampled in place (update_radiace): recomputes the cached env taps
        into the buffer the step's kernels already point at. A no-op when nothing changed."""
        if self.env_taps is not None:
            t = shading.refresh_env_taps(self.pc.incident_dirs, *self.env_hw)
            if t is None or t.data_ptr() != self.env_taps.data_ptr():
                raise RuntimeError("FusedTrainStep: the incident-direction buffer was replaced; rebuild the step")
        if self.rc is not None and self.rad.get("taps") is not None:
            shading.refresh_env_taps(self.rc.incident_dirs, *self.env_hw)

    def calibrate(self) -> int:
        """Sizes the binning buffers from one eager per-surfel preprocess of the current camera (one host sync)."""
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(self.L.svgir_raster_preprocess(C.byref(self.cfg), C.byref(self.cin), C.byref(self.cst), C.byref(self.cout),
                                                      C.c_void_p(cur.cuda_stream)), "raster_preprocess")
        R = int(self.t["num_rendered"][0].item())
        want = int(R * raster.ASYNC_SLACK) + raster.ASYNC_MARGIN
        if want > self.cap:
            self._alloc_bins(want)
        return R

    def grow(self, R: int):
        """After an overflow: bins for R instances (the owner re-captures: the buffers moved)."""
        self._alloc_bins(int(R * raster.ASYNC_SLACK) + raster.ASYNC_MARGIN)
        raster._CAP_HINT[(self.dev.index, self.P, self.W, self.H)] = self.cap

    def read_count(self):
        """(num_rendered, overflowed) of the last step; the caller has synchronised with the step."""
        R, ov = self.count_host.tolist()
        return int(R), bool(ov)

    # ---------------------------------------------------------------------------------------------------------------
    def enqueue(self):
        """Enqueues one whole step on the current stream (+ the side stream, forked from and joined to it). Nothing
        here blocks the host; capturable into a CUDA graph. Returns the loss (a view of the static loss buffer)."""
        L, dev = self.L, self.dev
        chk = _lib.check
        cur = torch.cuda.current_stream(dev)
        side = self.side
        cs, ss = C.c_void_p(cur.cuda_stream), C.c_void_p(side.cuda_stream)
        cfg, cin, cst, cout = C.byref(self.cfg), C.byref(self.cin), C.byref(self.cst), C.byref(self.cout)
        n0 = _lib.launch_count()
        with torch.cuda.device(dev):
            if not self.bucket.attached():
                self._bind_grads()
            if self.zero_grads:
                self.bucket.flat.zero_()
            self.arena.zero_()
            chk(L.svgir_raster_preprocess(cfg, cin, cst, cout, cs), "raster_preprocess")
            # binning needs the geometry only: side stream, under the shading forward
            fork = torch.cuda.Event()
            fork.record(cur)
            side.wait_event(fork)
            chk(L.svgir_raster_bin(cfg, cin, cst, cout, ss), "raster_bin")
            binned = torch.cuda.Event()
            binned.record(side)
            if self.rc is not None:
                # radiance-consistency term: needs the camera centre only. Its backward adds (atomically) into the same
                # albedo / roughness / env gradients the shading backward accumulates into, so it is ordered before it;
                # the overflow flag it tests is final once binning is (same stream).
                r = self.rad
                gv = self.gviews
                chk(L.svgir_radiance_loss_forward_backward(
                    C.byref(self.rcfg), C.byref(self.rin), r["grad"].data_ptr(), r["loss"].data_ptr(), r["irr"].data_ptr(), None,
                    r["saved"].data_ptr(), r["scratch"].data_ptr(), gv[5].data_ptr(), gv[6].data_ptr(), gv[8].data_ptr(),
                    r["env_scratch"].data_ptr(), ss), "radiance_loss_forward_backward")
                rad_done = torch.cuda.Event()
                rad_done.record(side)
            if self.fwd_reserve:   # leave SMs to the binning kernels running beside it (see FWD_RESERVE_SMS)
                L.svgir_shade_reserve_sms(self.fwd_reserve)
            chk(L.svgir_shade_forward(C.byref(self.scfg_f), C.byref(self.sin), C.byref(self.sout), cs), "shade_forward")
            if self.fwd_reserve:
                L.svgir_shade_reserve_sms(0)
            cur.wait_event(binned)
            chk(L.svgir_raster_composite(cfg, cin, cst, cout, cs), "raster_composite")
            self.count_host.copy_(self.t["num_rendered"], non_blocking=True)
            chk(L.svgir_train_loss_forward(C.byref(self.lcfg), C.byref(self.lin), self.loss_out.data_ptr(),
                                           self.loss_partials.data_ptr(), self.loss_counter.data_ptr(), cs), "train_loss_forward")
            chk(L.svgir_train_loss_backward(C.byref(self.lcfg), C.byref(self.lin), None, self.loss_out.data_ptr(),
                                            C.byref(self.lgr), cs), "train_loss_backward")
            chk(L.svgir_raster_backward_composite(cfg, cin, cst, C.byref(self.rgrads), cs), "raster_backward_composite")
            # parameter backward of the rasteriser (+ the all-reduce of its gradient segment) on the side stream, under the
            # shading backward
            fork2 = torch.cuda.Event()
            fork2.record(cur)
            side.wait_event(fork2)
            chk(L.svgir_raster_backward_params(cfg, cin, cst, self.geo.data_ptr(), C.byref(self.pg), ss), "raster_backward_params")
            seg = self.bucket.seg_bounds if self.reduce_in_step else []
            if self.reduce_in_step:
                # this rank's binning-overflow flag rides in the last segment: every rank sees "somebody overflowed"
                self.bucket.extra[0:1].copy_(self.t["num_rendered"][1:2])
                for k, (lo, hi) in enumerate(seg[:-1]):
                    lo4, hi4 = lo // 4 * 4, _al4(hi)
                    chk(L.svgir_peer_allreduce_range(self.peer.comm, lo4, hi4 - lo4, k, self.peer.BG_GRID, ss), "peer_allreduce")
                if len(seg) > 1:
                    L.svgir_shade_reserve_sms(self.peer.BG_GRID)
            done = torch.cuda.Event()
            done.record(side)
            if self.rc is not None:
                cur.wait_event(rad_done)
            chk(L.svgir_shade_backward(C.byref(self.scfg_b), C.byref(self.sin), C.byref(self.sgr), cs), "shade_backward")
            L.svgir_shade_reserve_sms(0)
            cur.wait_event(done)
            if self.reduce_in_step:
                lo, hi = seg[-1]
                lo4, hi4 = lo // 4 * 4, _al4(hi)
                chk(L.svgir_peer_allreduce_range(self.peer.comm, lo4, hi4 - lo4, len(seg) - 1, 0, cs), "peer_allreduce")
                self.flag_host.copy_(self.bucket.extra[0:1], non_blocking=True)
        self.launches = _lib.launch_count() - n0
        return self.loss_out[0]
