"""Fused tail of a stage-2 training iteration (G-buffer resolve + image loss), forward and backward.

`fused_train_loss` computes, from the rasteriser's RAW outputs, exactly what
`pipeline.image_loss(pipeline.render_view(...))` computes with ~150 torch kernels -- the
un-premultiply / opacity filter / rgb_to_srgb of gaussian_renderer/svgss.py:187-233 followed by the L1
terms of calculate_loss (svgss.py:280-294) and its 0.02-weighted surface term
cos_loss(rendered_normal, depth2normal(rendered_depth, image_mask, camera)) (svgss.py:300-313,
utils/image_utils.py:61-125, utils/loss_utils.py:117-119), whose gradient reaches the shading normal and the
rendered depth -- with one CUDA kernel per direction (csrc/resolve.cu). `fused_ssim`, `fused_edge_aware` and
`fused_tv` are the other image-space terms of calculate_loss (SSIM, svgss.py:282-293; edge-aware smoothness of
base colour / roughness, :366-378; TV of the env map, :386-390) as CUDA kernels (csrc/ssim.cu, csrc/loss_terms.cu).
The `*_torch` functions are the host-side torch mirror of the same reference code, used by `pipeline.image_loss`
(the un-fused tail) and as the checker of the kernels. No CPU / torch fallback inside the fused functions.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class TrainLossCfg(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("S", C.c_int32), ("NV", C.c_int32),
                ("pbr_ch", C.c_int32), ("normal_ch", C.c_int32),
                ("lambda_pbr", C.c_float), ("lambda_normal", C.c_float), ("bg", C.c_void_p),
                ("normal_mode", C.c_int32), ("inv_focal_x", C.c_float), ("inv_focal_y", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("reserved_", C.c_int32)]


class TrainLossIn(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("color", "geo_normal", "opacity", "vfeature", "gt", "depth", "mask")]


NORMAL_GEO = 0   # round-1 stand-in: mean(1 - <n_shade, rasteriser geo normal>)
NORMAL_D2N = 1   # the reference's term: cos_loss(n_shade, depth2normal(depth))


def d2n_camera_terms(H: int, W: int, tanfovx: float, tanfovy: float, prcppoint=(0.5, 0.5)):
    """(inv_focal_x, inv_focal_y, cx, cy) of depth2normal (utils/image_utils.py:73-81): the x coordinate is divided by
    fov2focal(FoVy, image_height) and y by fov2focal(FoVx, image_width) (the reference's own pairing; identical for
    square images); the principal point is prcppoint * (W, H)."""
    focal_x = H / (2.0 * float(tanfovy))
    focal_y = W / (2.0 * float(tanfovx))
    return 1.0 / focal_x, 1.0 / focal_y, float(prcppoint[0]) * W, float(prcppoint[1]) * H


class TrainLossGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("color", "geo_normal", "depth", "opacity", "feature", "vfeature")]


_bound = False


def _bind():
    global _bound
    L = _lib.lib()
    if not _bound:
        L.svgir_train_loss_blocks.argtypes = [C.c_int, C.c_int]
        L.svgir_train_loss_blocks.restype = C.c_int
        L.svgir_train_loss_forward.argtypes = [C.POINTER(TrainLossCfg), C.POINTER(TrainLossIn), C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p]
        L.svgir_train_loss_forward.restype = C.c_int
        L.svgir_train_loss_backward.argtypes = [C.POINTER(TrainLossCfg), C.POINTER(TrainLossIn), C.c_void_p, C.c_void_p,
                                                C.POINTER(TrainLossGrads), C.c_void_p]
        L.svgir_train_loss_backward.restype = C.c_int
        _bound = True
    return L


_SCRATCH: dict = {}


def _scratch(dev, nblocks):
    """(partials, counter) per device; the kernel leaves the counter at zero."""
    key = (dev.index, nblocks)
    s = _SCRATCH.get(key)
    if s is None:
        s = _SCRATCH[key] = (torch.empty(4 * nblocks, dtype=torch.float32, device=dev),
                             torch.zeros(1, dtype=torch.int32, device=dev))
    return s


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise RuntimeError("svgir_b200.losses: expected CUDA float32 tensors (no CPU fallback)")
    return t.contiguous()


class _FusedTrainLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, geo_normal, opacity, vfeature, gt, bg, depth, mask, lambda_pbr, lambda_normal, pbr_ch, normal_ch,
                normal_mode, cam_terms):
        L = _bind()
        color, geo_normal, opacity, vfeature, gt, bg = map(_f32c, (color, geo_normal, opacity, vfeature, gt, bg))
        depth = _f32c(depth) if depth is not None else None
        mask = _f32c(mask) if mask is not None else None
        if normal_mode != NORMAL_GEO and depth is None:
            raise ValueError("fused_train_loss: the depth2normal surface term needs the rendered depth")
        H, W = int(color.shape[-2]), int(color.shape[-1])
        NV = int(vfeature.shape[0])
        dev = color.device
        ifx, ify, cx, cy = cam_terms
        cfg = TrainLossCfg(W, H, 0, NV, int(pbr_ch), int(normal_ch), float(lambda_pbr), float(lambda_normal), bg.data_ptr(),
                           int(normal_mode), ifx, ify, cx, cy, 0)
        cin = TrainLossIn(color.data_ptr(), geo_normal.data_ptr(), opacity.data_ptr(), vfeature.data_ptr(), gt.data_ptr(),
                          depth.data_ptr() if depth is not None else None, mask.data_ptr() if mask is not None else None)
        out = torch.empty(8, dtype=torch.float32, device=dev)
        partials, counter = _scratch(dev, L.svgir_train_loss_blocks(W, H))
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(L.svgir_train_loss_forward(C.byref(cfg), C.byref(cin), out.data_ptr(), partials.data_ptr(),
                                                  counter.data_ptr(), stream), "train_loss_forward")
        ctx.save_for_backward(*[t for t in (color, geo_normal, opacity, vfeature, gt, bg, out, depth, mask) if t is not None])
        ctx.has = (depth is not None, mask is not None)
        ctx.params = (W, H, NV, int(pbr_ch), int(normal_ch), float(lambda_pbr), float(lambda_normal), int(normal_mode), cam_terms)
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, grad_loss, _grad_terms):
        L = _bind()
        saved = list(ctx.saved_tensors)
        color, geo_normal, opacity, vfeature, gt, bg, out = saved[:7]
        rest = saved[7:]
        depth = rest.pop(0) if ctx.has[0] else None
        mask = rest.pop(0) if ctx.has[1] else None
        W, H, NV, pbr_ch, normal_ch, lp, ln, mode, (ifx, ify, cx, cy) = ctx.params
        dev = color.device
        cfg = TrainLossCfg(W, H, 0, NV, pbr_ch, normal_ch, lp, ln, bg.data_ptr(), mode, ifx, ify, cx, cy, 0)
        cin = TrainLossIn(color.data_ptr(), geo_normal.data_ptr(), opacity.data_ptr(), vfeature.data_ptr(), gt.data_ptr(),
                          depth.data_ptr() if depth is not None else None, mask.data_ptr() if mask is not None else None)
        g_color, g_normal = torch.empty_like(color), torch.empty_like(geo_normal)
        g_opacity, g_vfeature = torch.empty_like(opacity), torch.empty_like(vfeature)
        g_depth = torch.empty_like(depth) if (depth is not None and mode != NORMAL_GEO) else None
        g = TrainLossGrads(g_color.data_ptr(), g_normal.data_ptr(), g_depth.data_ptr() if g_depth is not None else None,
                           g_opacity.data_ptr(), None, g_vfeature.data_ptr())
        grad_loss = _f32c(grad_loss.reshape(1))
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(L.svgir_train_loss_backward(C.byref(cfg), C.byref(cin), grad_loss.data_ptr(), out.data_ptr(), C.byref(g),
                                                   stream), "train_loss_backward")
        return g_color, g_normal, g_opacity, g_vfeature, None, None, g_depth, None, None, None, None, None, None, None


def fused_train_loss(color, geo_normal, opacity, vfeature, gt_image, bg, lambda_pbr=1.0, lambda_normal=0.02,
                     pbr_ch=0, normal_ch=6, depth=None, mask=None, cam_terms=None):
    """Returns (loss, terms[8] = total, l1, l1_pbr, surface term, pixels in the surface term's mean, -, -, -) from the
    rasteriser's raw outputs (`rendered_image, rendered_normal, rendered_opacity, rendered_vfeature` of svgss.py:171-184).
    With `depth` (the rasteriser's depth image) and `cam_terms` = d2n_camera_terms(...) the surface term is the
    reference's cos_loss(normal, depth2normal(depth, mask, camera)); without them the round-1 stand-in
    mean(1 - <normal, geo_normal>) (kept for A/B)."""
    mode = NORMAL_D2N if (depth is not None and cam_terms is not None) else NORMAL_GEO
    return _FusedTrainLoss.apply(color, geo_normal, opacity, vfeature, gt_image, bg, depth if mode == NORMAL_D2N else None,
                                 mask, lambda_pbr, lambda_normal, pbr_ch, normal_ch, mode,
                                 cam_terms if cam_terms is not None else (0.0, 0.0, 0.0, 0.0))


# ---- torch mirror of the reference's loss-tail functions (host side; also the checker of the kernels) ----------------
def depth2normal_torch(depth, mask, H, W, cam_terms):
    """utils/image_utils.py:61-125 on [1,H,W] depth / mask (mask None = ones); returns [3,H,W]."""
    ifx, ify, cx, cy = cam_terms
    dev, dt = depth.device, depth.dtype
    ys, xs = torch.meshgrid(torch.arange(H, device=dev, dtype=dt), torch.arange(W, device=dev, dtype=dt), indexing="ij")
    d = depth[0]
    P = torch.stack([(xs - cx) * d * ifx, (ys - cy) * d * ify, d], -1)            # [H,W,3]
    m = torch.ones((H, W, 1), device=dev, dtype=torch.bool) if mask is None else (mask[0] != 0)[..., None]
    Pp = torch.nn.functional.pad(P.permute(2, 0, 1)[None], [1, 1, 1, 1], mode="replicate")[0].permute(1, 2, 0)
    mp = torch.nn.functional.pad(m.permute(2, 0, 1)[None].to(dt), [1, 1, 1, 1], mode="replicate")[0].permute(1, 2, 0) != 0
    pc = Pp[1:-1, 1:-1] * mp[1:-1, 1:-1]
    pu = (Pp[:-2, 1:-1] - pc) * mp[:-2, 1:-1]
    pl = (Pp[1:-1, :-2] - pc) * mp[1:-1, :-2]
    pb = (Pp[2:, 1:-1] - pc) * mp[2:, 1:-1]
    pr = (Pp[1:-1, 2:] - pc) * mp[1:-1, 2:]
    cr = lambda a, b: torch.linalg.cross(a, b, dim=-1)
    n = cr(pu, pl) + cr(pr, pu) + cr(pb, pr) + cr(pl, pb)
    n = torch.nn.functional.normalize(n, dim=-1)
    return (n * mp[1:-1, 1:-1]).permute(2, 0, 1)


def cos_loss_torch(output, gt):
    """utils/loss_utils.py:117-119 with thrsh = 0, weight = 1."""
    cos = (output * gt).sum(0)
    return (1 - cos[cos < 1.0]).mean()


def spatial_gradient_torch(x):
    """kornia 0.6.12 `spatial_gradient(x[None], mode='sobel', order=1, normalized=True)[0]` for x [C,H,W] -> [C,2,H,W]."""
    Cn, H, W = x.shape
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], dtype=x.dtype, device=x.device) / 8.0
    k = torch.stack([kx, kx.t()])[:, None]
    xp = torch.nn.functional.pad(x.reshape(Cn, 1, H, W), [1, 1, 1, 1], mode="replicate")
    return torch.nn.functional.conv2d(xp, k)


def edge_aware_torch(data, img):
    """utils/loss_utils.py:103-104."""
    return (spatial_gradient_torch(data).abs() * torch.exp(-spatial_gradient_torch(img).abs())).sum(1).mean()


def tv_torch(x):
    """utils/loss_utils.py:112-116."""
    return torch.square(x[..., 1:, :] - x[..., :-1, :]).mean() + torch.square(x[..., :, 1:] - x[..., :, :-1]).mean()


class ResolveEvalOut(C.Structure):
    """svgir_resolve_eval_out (include/svgir_b200.h)."""
    _fields_ = [(n, C.c_void_p) for n in ("pbr", "normal", "base_color", "roughness", "lights", "local_lights",
                                           "visibility", "direct", "indirect")]


def resolve_eval(opacity: torch.Tensor, feature: torch.Tensor, vfeature: torch.Tensor, bg: torch.Tensor) -> dict:
    """The torch tail of render_view's eval branch (gaussian_renderer/svgss.py:187-262, is_training=False) as ONE
    kernel: raw (opacity-premultiplied) feature [7,H,W] / vfeature [16,H,W] -> the nine [3,H,W] result images.
    Forward only (the eval drivers run under torch.no_grad())."""
    L = _lib.lib()
    if feature.shape[0] != 7 or vfeature.shape[0] != 16:
        raise ValueError("resolve_eval expects the eval G-buffer: feature [7,H,W], vfeature [16,H,W]")
    H, W = int(opacity.shape[-2]), int(opacity.shape[-1])
    f32 = dict(dtype=torch.float32, device=opacity.device)
    names = [n for n, _ in ResolveEvalOut._fields_]
    store = torch.empty((len(names), 3, H, W), **f32)   # one allocation for the nine images
    out = ResolveEvalOut(*[store[i].data_ptr() for i in range(len(names))])
    opacity, feature, vfeature = opacity.contiguous(), feature.contiguous(), vfeature.contiguous()
    bg = bg.to(**f32).contiguous()
    L.svgir_resolve_eval.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(ResolveEvalOut), C.c_void_p]
    L.svgir_resolve_eval.restype = C.c_int
    _lib.check(L.svgir_resolve_eval(W, H, bg.data_ptr(), opacity.data_ptr(), feature.data_ptr(), vfeature.data_ptr(),
                                    C.byref(out), torch.cuda.current_stream(opacity.device).cuda_stream), "resolve_eval")
    return {n: store[i] for i, n in enumerate(names)}


# ---- SSIM ----------------------------------------------------------------------------------------
_SSIM_SCRATCH: dict = {}


def _ssim_bind():
    L = _lib.lib()
    if not getattr(_ssim_bind, "done", False):
        L.svgir_ssim_blocks.argtypes = [C.c_int, C.c_int, C.c_int]
        L.svgir_ssim_blocks.restype = C.c_int
        L.svgir_ssim_forward.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        L.svgir_ssim_forward.restype = C.c_int
        L.svgir_ssim_backward.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        L.svgir_ssim_backward.restype = C.c_int
        _ssim_bind.done = True
    return L


class _FusedSSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2):
        L = _ssim_bind()
        if img1.shape != img2.shape or img1.dim() != 3:
            raise ValueError("fused_ssim expects two [C,H,W] images of the same shape")
        if not img1.is_cuda:
            raise RuntimeError("fused_ssim: CUDA tensors only (there is no CPU fallback)")
        x, y = img1.detach().float().contiguous(), img2.detach().float().contiguous()
        Cn, H, W = (int(v) for v in x.shape)
        dev = x.device
        nb = int(L.svgir_ssim_blocks(Cn, H, W))
        key = (dev.index, nb)
        sc = _SSIM_SCRATCH.get(key)
        if sc is None:
            sc = _SSIM_SCRATCH[key] = (torch.empty(nb, dtype=torch.float32, device=dev),
                                       torch.zeros(1, dtype=torch.int32, device=dev))
        out = torch.empty(1, dtype=torch.float32, device=dev)
        gmaps = torch.empty((3, Cn, H, W), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.svgir_ssim_forward(Cn, H, W, x.data_ptr(), y.data_ptr(), out.data_ptr(),
                                        gmaps.data_ptr() if gmaps is not None else None, sc[0].data_ptr(), sc[1].data_ptr(),
                                        stream), "ssim_forward")
        ctx.save_for_backward(x, y, gmaps)
        return out[0]

    @staticmethod
    def backward(ctx, grad):
        L = _ssim_bind()
        x, y, gmaps = ctx.saved_tensors
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("fused_ssim: no gradient with respect to the second (ground-truth) image")
        if gmaps is None:
            return None, None
        Cn, H, W = (int(v) for v in x.shape)
        g = grad.detach().float().reshape(1).contiguous()
        d = torch.empty_like(x)
        _lib.check(L.svgir_ssim_backward(Cn, H, W, x.data_ptr(), y.data_ptr(), gmaps.data_ptr(), g.data_ptr(), d.data_ptr(),
                                         torch.cuda.current_stream(x.device).cuda_stream), "ssim_backward")
        return d, None


def fused_ssim(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    """`ssim(img1, img2)` of utils/loss_utils.py:32-62 ([C,H,W] images, 11x11 Gaussian window, mean) as one CUDA kernel
    per direction (csrc/ssim.cu). Differentiable with respect to img1; img2 is the ground truth."""
    return _FusedSSIM.apply(img1, img2)


# ---- edge-aware smoothness and TV ------------------------------------------------------------------
_EA_SCRATCH: dict = {}


def _terms_bind():
    L = _lib.lib()
    if not getattr(_terms_bind, "done", False):
        L.svgir_edge_aware_blocks.argtypes = [C.c_int, C.c_int, C.c_int]
        L.svgir_edge_aware_blocks.restype = C.c_int
        L.svgir_edge_aware_forward.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p]
        L.svgir_edge_aware_forward.restype = C.c_int
        L.svgir_edge_aware_backward.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p]
        L.svgir_edge_aware_backward.restype = C.c_int
        L.svgir_tv_loss.argtypes = [C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.svgir_tv_loss.restype = C.c_int
        _terms_bind.done = True
    return L


class _FusedEdgeAware(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, img, mask):
        L = _terms_bind()
        if data.shape != img.shape or data.dim() != 3:
            raise ValueError("fused_edge_aware expects data and img of the same [C,H,W] shape")
        if not data.is_cuda:
            raise RuntimeError("fused_edge_aware: CUDA tensors only (there is no CPU fallback)")
        x, y = data.detach().float().contiguous(), img.detach().float().contiguous()
        m = mask.detach().float().contiguous() if mask is not None else None
        Cn, H, W = (int(v) for v in x.shape)
        if m is not None and m.numel() != H * W:
            raise ValueError("fused_edge_aware: mask must be [1,H,W]")
        dev = x.device
        nb = int(L.svgir_edge_aware_blocks(Cn, H, W))
        key = (dev.index, nb)
        sc = _EA_SCRATCH.get(key)
        if sc is None:
            sc = _EA_SCRATCH[key] = (torch.empty(nb, dtype=torch.float32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
        out = torch.empty(1, dtype=torch.float32, device=dev)
        _lib.check(L.svgir_edge_aware_forward(Cn, H, W, x.data_ptr(), y.data_ptr(), m.data_ptr() if m is not None else None,
                                              out.data_ptr(), sc[0].data_ptr(), sc[1].data_ptr(),
                                              torch.cuda.current_stream(dev).cuda_stream), "edge_aware_forward")
        ctx.save_for_backward(*[t for t in (x, y, m) if t is not None])
        ctx.has_mask = m is not None
        return out[0]

    @staticmethod
    def backward(ctx, grad):
        L = _terms_bind()
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("fused_edge_aware: no gradient with respect to the guide image")
        sv = list(ctx.saved_tensors)
        x, y = sv[0], sv[1]
        m = sv[2] if ctx.has_mask else None
        Cn, H, W = (int(v) for v in x.shape)
        g = grad.detach().float().reshape(1).contiguous()
        d = torch.empty_like(x)
        _lib.check(L.svgir_edge_aware_backward(Cn, H, W, x.data_ptr(), y.data_ptr(), m.data_ptr() if m is not None else None,
                                               g.data_ptr(), d.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream),
                   "edge_aware_backward")
        return d, None, None


def fused_edge_aware(data: torch.Tensor, img: torch.Tensor, mask: torch.Tensor = None) -> torch.Tensor:
    """`first_order_edge_aware_loss(data * mask, img * mask)` of utils/loss_utils.py:103-104 as the application calls it
    (svgss.py:366-378: rendered base colour / roughness against the ground-truth image), one CUDA kernel per direction.
    Differentiable with respect to `data`."""
    return _FusedEdgeAware.apply(data, img, mask)


class _FusedTV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        L = _terms_bind()
        if x.dim() != 3 or not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("fused_tv expects a CUDA float32 [C,H,W] tensor (any strides; there is no CPU fallback)")
        Cn, H, W = (int(v) for v in x.shape)
        xd = x.detach()
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        grad = torch.empty_strided(tuple(xd.shape), tuple(xd.stride()), dtype=torch.float32, device=x.device)
        sc, sh, sw = (int(v) for v in xd.stride())
        _lib.check(L.svgir_tv_loss(Cn, H, W, sc, sh, sw, xd.data_ptr(), None, out.data_ptr(), grad.data_ptr(),
                                   torch.cuda.current_stream(x.device).cuda_stream), "tv_loss")
        ctx.save_for_backward(grad)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g


def fused_tv(x: torch.Tensor) -> torch.Tensor:
    """`tv_loss(x)` of utils/loss_utils.py:112-116 for x [C,H,W] with arbitrary strides -- the application passes
    `env[0].permute(2, 0, 1)` (svgss.py:388), which is read in place. One launch computes the loss and its gradient."""
    return _FusedTV.apply(x)
