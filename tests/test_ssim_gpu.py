"""Fused SSIM kernels (csrc/ssim.cu via svgir_b200.losses.fused_ssim) against goldens from the reference's own `ssim`
(utils/loss_utils.py:32-62 + torch autograd, tests/golden/ref_ssim.npz) and against a torch restatement at 800x800."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _torch_ssim(img1, img2, ws=11):
    """Restatement of loss_utils.py:21-62 on the tensors' device."""
    C = img1.shape[0]
    g = torch.tensor([math.exp(-(x - ws // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(ws)], dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)
    window = g.mm(g.t()).unsqueeze(0).unsqueeze(0).expand(C, 1, ws, ws).contiguous().to(img1)
    a, b = img1[None], img2[None]
    mu1, mu2 = F.conv2d(a, window, padding=ws // 2, groups=C), F.conv2d(b, window, padding=ws // 2, groups=C)
    s1 = F.conv2d(a * a, window, padding=ws // 2, groups=C) - mu1 * mu1
    s2 = F.conv2d(b * b, window, padding=ws // 2, groups=C) - mu2 * mu2
    s12 = F.conv2d(a * b, window, padding=ws // 2, groups=C) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))).mean()


def test_fused_ssim_matches_reference_golden():
    from svgir_b200 import losses
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(HERE, "golden", "ref_ssim.npz"))
    for tag in ("a", "b", "c"):
        x = torch.from_numpy(g[f"{tag}_img1"]).to(dev).requires_grad_(True)
        y = torch.from_numpy(g[f"{tag}_img2"]).to(dev)
        v = losses.fused_ssim(x, y)
        v.backward()
        assert abs(float(v) - float(g[f"{tag}_ssim"])) < 5e-6, tag          # fp32, tolerance written here
        ref = g[f"{tag}_grad"]
        err = np.abs(x.grad.cpu().numpy() - ref).max()
        assert err < 1e-5 * np.abs(ref).max() + 1e-9, (tag, err)


def test_fused_ssim_full_size_and_upstream_gradient():
    from svgir_b200 import losses
    dev = torch.device("cuda:0")
    gen = torch.Generator(dev).manual_seed(9)
    x = torch.rand(3, 800, 800, device=dev, generator=gen)
    y = (x + 0.1 * torch.randn(3, 800, 800, device=dev, generator=gen)).clamp(0, 1)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    la = 0.2 * (1.0 - losses.fused_ssim(xa, y))       # the way calculate_loss uses it (svgss.py:286)
    lb = 0.2 * (1.0 - _torch_ssim(xb, y))
    la.backward()
    lb.backward()
    assert abs(float(la) - float(lb)) < 5e-6
    rel = float((xa.grad - xb.grad).norm() / xb.grad.norm())
    assert rel < 1e-4, rel
    # deterministic: two launches give the same bits
    assert float(losses.fused_ssim(x, y)) == float(losses.fused_ssim(x, y))
