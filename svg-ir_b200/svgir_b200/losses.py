"""Fused tail of a stage-2 training iteration (G-buffer resolve + image loss), forward and backward.

`fused_train_loss` computes, from the rasteriser's RAW outputs, exactly what
`pipeline.image_loss(pipeline.render_view(...))` computes with ~120 torch kernels -- the
un-premultiply / opacity filter / rgb_to_srgb of gaussian_renderer/svgss.py:187-233 followed by the L1
terms of calculate_loss (svgss.py:280-294) and the 0.02-weighted normal-consistency term (svgss.py:313)
-- with one CUDA kernel per direction (csrc/resolve.cu). No CPU / torch fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class TrainLossCfg(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("S", C.c_int32), ("NV", C.c_int32),
                ("pbr_ch", C.c_int32), ("normal_ch", C.c_int32),
                ("lambda_pbr", C.c_float), ("lambda_normal", C.c_float), ("bg", C.c_void_p)]


class TrainLossIn(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("color", "geo_normal", "opacity", "vfeature", "gt")]


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
        L.svgir_train_loss_backward.argtypes = [C.POINTER(TrainLossCfg), C.POINTER(TrainLossIn), C.c_void_p,
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
        s = _SCRATCH[key] = (torch.empty(3 * nblocks, dtype=torch.float32, device=dev),
                             torch.zeros(1, dtype=torch.int32, device=dev))
    return s


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise RuntimeError("svgir_b200.losses: expected CUDA float32 tensors (no CPU fallback)")
    return t.contiguous()


class _FusedTrainLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, geo_normal, opacity, vfeature, gt, bg, lambda_pbr, lambda_normal, pbr_ch, normal_ch):
        L = _bind()
        color, geo_normal, opacity, vfeature, gt, bg = map(_f32c, (color, geo_normal, opacity, vfeature, gt, bg))
        H, W = int(color.shape[-2]), int(color.shape[-1])
        NV = int(vfeature.shape[0])
        dev = color.device
        cfg = TrainLossCfg(W, H, 0, NV, int(pbr_ch), int(normal_ch), float(lambda_pbr), float(lambda_normal), bg.data_ptr())
        cin = TrainLossIn(color.data_ptr(), geo_normal.data_ptr(), opacity.data_ptr(), vfeature.data_ptr(), gt.data_ptr())
        out = torch.empty(4, dtype=torch.float32, device=dev)
        partials, counter = _scratch(dev, L.svgir_train_loss_blocks(W, H))
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(L.svgir_train_loss_forward(C.byref(cfg), C.byref(cin), out.data_ptr(), partials.data_ptr(),
                                                  counter.data_ptr(), stream), "train_loss_forward")
        ctx.save_for_backward(color, geo_normal, opacity, vfeature, gt, bg)
        ctx.params = (W, H, NV, int(pbr_ch), int(normal_ch), float(lambda_pbr), float(lambda_normal))
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, grad_loss, _grad_terms):
        L = _bind()
        color, geo_normal, opacity, vfeature, gt, bg = ctx.saved_tensors
        W, H, NV, pbr_ch, normal_ch, lp, ln = ctx.params
        dev = color.device
        cfg = TrainLossCfg(W, H, 0, NV, pbr_ch, normal_ch, lp, ln, bg.data_ptr())
        cin = TrainLossIn(color.data_ptr(), geo_normal.data_ptr(), opacity.data_ptr(), vfeature.data_ptr(), gt.data_ptr())
        g_color, g_normal = torch.empty_like(color), torch.empty_like(geo_normal)
        g_opacity, g_vfeature = torch.empty_like(opacity), torch.empty_like(vfeature)
        g = TrainLossGrads(g_color.data_ptr(), g_normal.data_ptr(), None, g_opacity.data_ptr(), None, g_vfeature.data_ptr())
        grad_loss = _f32c(grad_loss.reshape(1))
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(L.svgir_train_loss_backward(C.byref(cfg), C.byref(cin), grad_loss.data_ptr(), C.byref(g), stream),
                       "train_loss_backward")
        return g_color, g_normal, g_opacity, g_vfeature, None, None, None, None, None, None


def fused_train_loss(color, geo_normal, opacity, vfeature, gt_image, bg, lambda_pbr=1.0, lambda_normal=0.02,
                     pbr_ch=0, normal_ch=6):
    """Returns (loss, terms[4] = total, l1, l1_pbr, normal) from the rasteriser's raw outputs
    (`rendered_image, rendered_normal, rendered_opacity, rendered_vfeature` of svgss.py:171-184)."""
    return _FusedTrainLoss.apply(color, geo_normal, opacity, vfeature, gt_image, bg, lambda_pbr, lambda_normal,
                                 pbr_ch, normal_ch)


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
