"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64 arithmetic on float32 inputs) of the reference's loss-tail
functions that the CUDA kernels csrc/resolve.cu (surface term) and csrc/loss_terms.cu (edge-aware, TV) implement:

  depth2normal                 /root/reference/utils/image_utils.py:61-125
  cos_loss                     /root/reference/utils/loss_utils.py:117-119
  first_order_edge_aware_loss  /root/reference/utils/loss_utils.py:103-104, with kornia 0.6.12's
                               spatial_gradient(mode='sobel', order=1, normalized=True) restated (kornia is a
                               third-party dependency absent from /root/reference and from this image; pinned by
                               /root/reference/readme.md:34): 3x3 Sobel kernels / 8, replicate padding, cross-correlation
  tv_loss                      /root/reference/utils/loss_utils.py:112-116

Pinned by tests/golden/ref_losses.npz, produced by the reference's own functions (tests/golden/make_golden_losses.py).
Only tests/ may import this module."""
import numpy as np


def _pad_replicate(a):
    """[..., H, W] -> [..., H+2, W+2]"""
    return np.pad(a, [(0, 0)] * (a.ndim - 2) + [(1, 1), (1, 1)], mode="edge")


def depth2normal(depth, mask, H, W, fovx, fovy, prcppoint=(0.5, 0.5)):
    """depth [1,H,W], mask [1,H,W] -> [3,H,W]. K00 = fov2focal(FoVy, H) scales x, K11 = fov2focal(FoVx, W) scales y
    (image_utils.py:77-81, the reference's own pairing)."""
    d = np.asarray(depth, np.float64)[0]
    m = np.asarray(mask, np.float64)[0] != 0 if mask is not None else np.ones((H, W), bool)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    k00 = H / (2.0 * np.tan(fovy / 2.0))
    k11 = W / (2.0 * np.tan(fovx / 2.0))
    P = np.stack([(xs - prcppoint[0] * W) * d / k00, (ys - prcppoint[1] * H) * d / k11, d], 0)     # [3,H,W]
    Pp = _pad_replicate(P)
    mp = _pad_replicate(m.astype(np.float64)) != 0
    c = (slice(1, -1), slice(1, -1))
    pc = Pp[(slice(None),) + c] * mp[c]
    pu = (Pp[:, :-2, 1:-1] - pc) * mp[:-2, 1:-1]
    pl = (Pp[:, 1:-1, :-2] - pc) * mp[1:-1, :-2]
    pb = (Pp[:, 2:, 1:-1] - pc) * mp[2:, 1:-1]
    pr = (Pp[:, 1:-1, 2:] - pc) * mp[1:-1, 2:]
    cr = lambda a, b: np.cross(a, b, axis=0)
    n = cr(pu, pl) + cr(pr, pu) + cr(pb, pr) + cr(pl, pb)
    n = n / np.maximum(np.sqrt((n * n).sum(0, keepdims=True)), 1e-12)
    return n * mp[c]


def cos_loss(output, gt):
    cos = (np.asarray(output, np.float64) * np.asarray(gt, np.float64)).sum(0)
    sel = cos < 1.0
    return float((1.0 - cos[sel]).mean())


def spatial_gradient(x):
    """[C,H,W] -> [C,2,H,W] (d/dx, d/dy)."""
    xp = _pad_replicate(np.asarray(x, np.float64))
    tl, tc, tr = xp[:, :-2, :-2], xp[:, :-2, 1:-1], xp[:, :-2, 2:]
    ml, mr = xp[:, 1:-1, :-2], xp[:, 1:-1, 2:]
    bl, bc, br = xp[:, 2:, :-2], xp[:, 2:, 1:-1], xp[:, 2:, 2:]
    gx = ((tr - tl) + 2.0 * (mr - ml) + (br - bl)) / 8.0
    gy = ((bl - tl) + 2.0 * (bc - tc) + (br - tr)) / 8.0
    return np.stack([gx, gy], 1)


def first_order_edge_aware_loss(data, img):
    return float((np.abs(spatial_gradient(data)) * np.exp(-np.abs(spatial_gradient(img)))).sum(1).mean())


def tv_loss(x):
    x = np.asarray(x, np.float64)
    return float(np.square(x[..., 1:, :] - x[..., :-1, :]).mean() + np.square(x[..., :, 1:] - x[..., :, :-1]).mean())
