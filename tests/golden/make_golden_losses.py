"""Generates tests/golden/ref_losses.npz by running the REFERENCE's own loss-tail functions on CPU with seeded inputs:

  * depth2normal          /root/reference/utils/image_utils.py:61-125
  * cos_loss, tv_loss, first_order_edge_aware_loss   /root/reference/utils/loss_utils.py:91-119
  * the way calculate_loss combines them            /root/reference/gaussian_renderer/svgss.py:298-313, 366-397

with torch autograd supplying the gradients. Run in the build container only:

    python tests/golden/make_golden_losses.py

Absent third-party modules are replaced by stubs (matplotlib, cv2, torchvision, tqdm: imported at module level by
image_utils.py, unused here). `kornia` (pinned by the reference's readme.md:34 to 0.6.12) is absent too and IS used by
first_order_edge_aware_loss: its `kornia.filters.spatial_gradient(x, mode='sobel', order=1, normalized=True)` is
restated below from the published 0.6.12 algorithm (Sobel kernels divided by the sum of their absolute values = 8,
replicate padding, cross-correlation, output [B,C,2,H,W] = d/dx, d/dy) and injected as the stub's function.
"""
import math
import os
import sys
import types
from unittest import mock

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def kornia_spatial_gradient(x, mode="sobel", order=1, normalized=True):
    assert mode == "sobel" and order == 1 and normalized
    b, c, h, w = x.shape
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], dtype=x.dtype) / 8.0
    k = torch.stack([kx, kx.t()])[:, None]                       # [2,1,3,3]
    xp = F.pad(x.reshape(b * c, 1, h, w), [1, 1, 1, 1], mode="replicate")
    return F.conv2d(xp, k).view(b, c, 2, h, w)


class Cam:
    def __init__(self, H, W, fovx, fovy, prcp=(0.5, 0.5)):
        self.image_height, self.image_width, self.FoVx, self.FoVy = H, W, fovx, fovy
        self.prcppoint = torch.tensor(prcp, dtype=torch.float32)


def main():
    for name in ("matplotlib", "matplotlib.colors", "matplotlib.pyplot", "cv2", "torchvision", "torchvision.utils", "tqdm",
                 "glob"):
        if name not in sys.modules:
            m = mock.MagicMock(name=name)
            m.__path__ = []
            m.__spec__ = None
            sys.modules[name] = m
    kornia = types.ModuleType("kornia")
    kornia.filters = types.ModuleType("kornia.filters")
    kornia.filters.spatial_gradient = kornia_spatial_gradient
    kornia.filters.laplacian = mock.MagicMock()
    sys.modules["kornia"], sys.modules["kornia.filters"] = kornia, kornia.filters
    sys.path.insert(0, REF)
    from utils.image_utils import depth2normal          # the reference's functions, unmodified
    from utils.loss_utils import cos_loss, first_order_edge_aware_loss, tv_loss

    out = {}
    g = torch.Generator().manual_seed(20261018)
    cases = {"a": (48, 64, 0.6911112, True), "b": (37, 29, 0.9, False), "c": (64, 64, 0.6911112, False)}
    for tag, (H, W, fovx, with_mask) in cases.items():
        fovy = 2 * math.atan(math.tan(fovx / 2) * H / W)
        cam = Cam(H, W, fovx, fovy, (0.5, 0.5) if tag != "b" else (0.47, 0.52))
        yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        depth = (3.5 + 0.4 * torch.sin(xx / 7.0) * torch.cos(yy / 5.0) + 0.05 * torch.rand(H, W, generator=g))[None]
        if with_mask:
            mask = (((xx - W / 2) ** 2 + (yy - H / 2) ** 2) < (0.42 * min(H, W)) ** 2).float()[None]
        else:
            mask = torch.ones(1, H, W)
        normal = F.normalize(torch.randn(3, H, W, generator=g) + torch.tensor([0.0, 0.0, -2.0])[:, None, None], dim=0)
        normal = normal * (0.6 + 0.4 * torch.rand(1, H, W, generator=g))      # the rendered normal is not unit length
        depth.requires_grad_(True)
        normal.requires_grad_(True)
        d2n = depth2normal(depth, mask, cam)
        loss = cos_loss(normal, d2n)
        loss.backward()
        out[f"{tag}_depth"], out[f"{tag}_mask"], out[f"{tag}_normal"] = depth.detach().numpy(), mask.numpy(), normal.detach().numpy()
        out[f"{tag}_cam"] = np.array([H, W, fovx, fovy, float(cam.prcppoint[0]), float(cam.prcppoint[1])], np.float64)
        out[f"{tag}_d2n"], out[f"{tag}_cos_loss"] = d2n.detach().numpy(), np.float32(loss.item())
        out[f"{tag}_g_depth"], out[f"{tag}_g_normal"] = depth.grad.numpy(), normal.grad.numpy()
        # edge-aware smoothness of a 3-channel image against the (masked) ground truth, svgss.py:366-378
        data = torch.rand(3, H, W, generator=g).requires_grad_(True)
        img = torch.rand(3, H, W, generator=g)
        le = first_order_edge_aware_loss(data * mask, img * mask)
        le.backward()
        out[f"{tag}_ea_data"], out[f"{tag}_ea_img"] = data.detach().numpy(), img.numpy()
        out[f"{tag}_ea_loss"], out[f"{tag}_ea_grad"] = np.float32(le.item()), data.grad.numpy()
    # tv_loss of the env map, svgss.py:386-390: env [1,He,We,3] -> env[0].permute(2,0,1)
    env = (3.0 * torch.rand(1, 32, 64, 3, generator=g)).requires_grad_(True)
    lt = tv_loss(env[0].permute(2, 0, 1))
    lt.backward()
    out["tv_env"], out["tv_loss"], out["tv_grad"] = env.detach().numpy(), np.float32(lt.item()), env.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "ref_losses.npz"), **out)
    print({k: (v.shape if getattr(v, "shape", ()) else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
