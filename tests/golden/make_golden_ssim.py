"""Generates tests/golden/ref_ssim.npz by running the REFERENCE's own `ssim` (/root/reference/utils/loss_utils.py:21-62)
on CPU with seeded inputs, with torch autograd supplying d ssim / d img1. Run in the build container only:

    python tests/golden/make_golden_ssim.py

`kornia` (imported at module level by loss_utils.py, not used by ssim) is replaced by a stub.
"""
import os
import sys
from unittest import mock

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    for name in ("kornia", "kornia.filters"):
        if name not in sys.modules:
            m = mock.MagicMock(name=name)
            m.__path__ = []
            m.__spec__ = None
            sys.modules[name] = m
    sys.path.insert(0, REF)
    from utils.loss_utils import ssim  # the reference's function, unmodified
    out = {}
    g = torch.Generator().manual_seed(20261017)
    for tag, (C, H, W) in {"a": (3, 37, 53), "b": (3, 64, 48), "c": (1, 11, 9)}.items():
        x = torch.rand(C, H, W, generator=g)
        y = (x + 0.2 * torch.randn(C, H, W, generator=g)).clamp(0, 1) if tag != "b" else torch.rand(C, H, W, generator=g)
        x.requires_grad_(True)
        v = ssim(x, y)   # the application calls it on [3,H,W] images (svgss.py:282)
        v.backward()
        out[f"{tag}_img1"], out[f"{tag}_img2"] = x.detach().numpy(), y.numpy()
        out[f"{tag}_ssim"], out[f"{tag}_grad"] = np.float32(v.item()), x.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "ref_ssim.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
