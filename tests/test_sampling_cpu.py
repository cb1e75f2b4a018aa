"""CPU: the sampler restatement in oracle/ against the golden produced by the reference's own
fibonacci_sphere_sampling (tests/golden/make_golden_sampling.py)."""
import os

import numpy as np
import torch

from oracle import render_equation_sh_oracle as RO

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_sampling.npz"))


def test_oracle_sampler_equals_reference_golden():
    n = torch.from_numpy(G["normals"])
    for ns in (24, 64, 100):
        d, a = RO.fibonacci_sphere_sampling(n, ns)
        assert np.abs(d.numpy() - G[f"dirs_fixed_{ns}"]).max() <= 1e-7
        assert (a.numpy() == G[f"areas_fixed_{ns}"]).all()
    d, _ = RO.fibonacci_sphere_sampling(n, 24, rand_u=torch.from_numpy(G["rand_u"]))
    assert np.abs(d.numpy() - G["dirs_random_24"]).max() <= 1e-7
    d, _ = RO.fibonacci_sphere_sampling(n[:256].reshape(16, 16, 3), 8)
    assert d.shape == (16, 16, 8, 3) and np.abs(d.numpy() - G["dirs_grid_8"]).max() <= 1e-7


def test_sampler_properties():
    torch.manual_seed(7)
    n = torch.nn.functional.normalize(torch.randn(500, 3), dim=-1)
    # rotation_between_z (utils/sh_utils.py:36-68) divides by 1+n.z: ill-conditioned at n = -z, as in the reference
    n = n[n[:, 2] > -0.999][:400]
    d, a = RO.fibonacci_sphere_sampling(n, 32)
    assert torch.allclose(d.norm(dim=-1), torch.ones(n.shape[0], 32), atol=1e-5)
    # hemisphere: z >= sin(10 deg) before the rotation, so n.d >= sin(10 deg)
    assert float((d * n[:, None]).sum(-1).min()) >= np.sin(np.pi / 18) - 1e-4
    assert torch.allclose(a, torch.full_like(a, 2 * np.pi))
