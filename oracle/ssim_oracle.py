"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's SSIM and of its gradient.

Follows /root/reference/utils/loss_utils.py:21-62 (`gaussian`, `create_window`, `ssim`, `_ssim`): 11-tap Gaussian
(sigma 1.5) window applied as a zero-padded depth-wise 2-D correlation to img1, img2, img1^2, img2^2, img1*img2;
C1 = 0.01^2, C2 = 0.03^2; ssim = mean of ((2 mu1 mu2 + C1)(2 s12 + C2)) / ((mu1^2 + mu2^2 + C1)(s1 + s2 + C2)).
The window is separable (outer product of the 1-D weights), which is how the CUDA kernels (csrc/ssim.cu) evaluate it.
The gradient with respect to img1 is written out analytically (the reference gets it from autograd):
    d ssim / d img1 = (1/N) [ w * g_mu1 + 2 img1 (w * g_e11) + img2 (w * g_e12) ]        (w * . = the same correlation)
with g_e11 = dm/ds1, g_e12 = 2 dm/dB, g_mu1 = 2 mu2 dm/dA + 2 mu1 dm/dC - 2 mu1 g_e11 - mu2 g_e12.
Pinned by tests/golden/ref_ssim.npz (the reference's own function + torch autograd, make_golden_ssim.py).
Only tests/ may import this module."""
import math

import numpy as np

WINDOW, SIGMA = 11, 1.5
C1, C2 = 0.01 ** 2, 0.03 ** 2


def gaussian_window(window_size: int = WINDOW, sigma: float = SIGMA) -> np.ndarray:
    """loss_utils.py:21-23: exp in double, stored as float32, normalised in float32."""
    g = np.array([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)], dtype=np.float32)
    return (g / g.sum(dtype=np.float32)).astype(np.float32)


def correlate(x: np.ndarray, g: np.ndarray) -> np.ndarray:
    """Zero-padded separable correlation of every channel of x [C,H,W] with g (x) g (conv2d, padding=window//2, groups=C)."""
    C, H, W = x.shape
    r = len(g) // 2
    xp = np.zeros((C, H, W + 2 * r), np.float64)
    xp[:, :, r:r + W] = x
    t = sum(float(g[k]) * xp[:, :, k:k + W] for k in range(len(g)))
    tp = np.zeros((C, H + 2 * r, W), np.float64)
    tp[:, r:r + H, :] = t
    return sum(float(g[k]) * tp[:, k:k + H, :] for k in range(len(g)))


def ssim_maps(img1: np.ndarray, img2: np.ndarray):
    g = gaussian_window()
    x, y = img1.astype(np.float64), img2.astype(np.float64)
    mu1, mu2 = correlate(x, g), correlate(y, g)
    s1 = correlate(x * x, g) - mu1 * mu1
    s2 = correlate(y * y, g) - mu2 * mu2
    s12 = correlate(x * y, g) - mu1 * mu2
    A, B = 2 * mu1 * mu2 + C1, 2 * s12 + C2
    Cc, D = mu1 * mu1 + mu2 * mu2 + C1, s1 + s2 + C2
    m = A * B / (Cc * D)
    g_e11 = -m / D
    g_e12 = 2 * A / (Cc * D)
    g_mu1 = 2 * mu2 * B / (Cc * D) - 2 * mu1 * m / Cc - 2 * mu1 * g_e11 - mu2 * g_e12
    return m, g_mu1, g_e11, g_e12


def ssim(img1: np.ndarray, img2: np.ndarray) -> float:
    return float(ssim_maps(img1, img2)[0].mean())


def ssim_grad(img1: np.ndarray, img2: np.ndarray) -> np.ndarray:
    """d ssim(img1, img2) / d img1, [C,H,W]."""
    m, g_mu1, g_e11, g_e12 = ssim_maps(img1, img2)
    g = gaussian_window()
    x, y = img1.astype(np.float64), img2.astype(np.float64)
    return ((correlate(g_mu1, g) + 2 * x * correlate(g_e11, g) + y * correlate(g_e12, g)) / m.size).astype(np.float32)
