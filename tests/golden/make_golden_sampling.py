"""Generates tests/golden/ref_sampling.npz by running the REFERENCE's own fibonacci_sphere_sampling /
rotation_between_z (utils/graphics_utils.py:9-37, utils/sh_utils.py:36-68) imported from /root/reference.
Those functions hard-code device='cuda'; the torch factory functions are wrapped so that request lands
on the CPU (values are device-independent fp32 elementwise math). Run in the build container:
    python tests/golden/make_golden_sampling.py
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _cpu(fn):
    def w(*a, **k):
        if k.get("device") == "cuda":
            k["device"] = "cpu"
        return fn(*a, **k)
    return w


def main():
    for name in ("arange", "rand", "zeros", "eye"):
        setattr(torch, name, _cpu(getattr(torch, name)))
    sys.path.insert(0, REF)
    from utils.graphics_utils import fibonacci_sphere_sampling
    g = torch.Generator().manual_seed(77)
    n = torch.nn.functional.normalize(torch.randn(257, 3, generator=g), dim=-1)
    n[0] = torch.tensor([0.0, 0.0, -1.0])   # the -I branch of rotation_between_z
    n[1] = torch.tensor([0.0, 0.0, 1.0])
    n[2] = torch.nn.functional.normalize(torch.tensor([1e-4, -2e-4, -1.0]), dim=0)
    out = {"normals": n.numpy()}
    for ns in (24, 64, 100):
        d, a = fibonacci_sphere_sampling(n, ns, random_rotate=False)
        out[f"dirs_fixed_{ns}"] = d.numpy()
        out[f"areas_fixed_{ns}"] = a.numpy()
    # random_rotate=True: make torch.rand deterministic and record what it returned
    u = torch.rand(257, 1, generator=g)
    real_rand = torch.rand
    torch.rand = lambda *a, **k: u.clone()
    d, a = fibonacci_sphere_sampling(n, 24, random_rotate=True)
    torch.rand = real_rand
    out["rand_u"] = u.numpy()
    out["dirs_random_24"] = d.numpy()
    # a batched [H,W,3] call exercises the pre_shape reshape
    d2, a2 = fibonacci_sphere_sampling(n[:256].reshape(16, 16, 3), 8, random_rotate=False)
    out["dirs_grid_8"] = d2.numpy()
    np.savez_compressed(os.path.join(HERE, "ref_sampling.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
