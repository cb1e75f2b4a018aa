"""GPU: svgir_b200.sampling (csrc/sampling.cu) against the reference golden and the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_sampling.npz"))
TOL = 2e-6  # fp32 directions; the reference's 3x3 product is a cuBLAS/loop bmm whose accumulation order is unspecified


def test_matches_reference_golden():
    from svgir_b200 import sampling
    n = torch.from_numpy(G["normals"]).cuda()
    for ns in (24, 64, 100):
        d, a = sampling.fibonacci_sphere_sampling(n, ns, random_rotate=False)
        assert d.shape == (257, ns, 3) and a.shape == (257, ns, 1)
        assert np.abs(d.cpu().numpy() - G[f"dirs_fixed_{ns}"]).max() <= TOL
        assert (a.cpu().numpy() == G[f"areas_fixed_{ns}"]).all()
    d, _ = sampling.fibonacci_sphere_sampling(n, 24, random_rotate=True, rand_u=torch.from_numpy(G["rand_u"]).cuda())
    # the azimuth offset goes through sinf/cosf of arguments up to ~60 rad: allow a few ulp of the argument
    assert np.abs(d.cpu().numpy() - G["dirs_random_24"]).max() <= 2e-5
    d, a = sampling.fibonacci_sphere_sampling(n[:256].reshape(16, 16, 3), 8, random_rotate=False)
    assert d.shape == (16, 16, 8, 3) and a.shape == (16, 16, 8, 1)
    assert np.abs(d.cpu().numpy() - G["dirs_grid_8"]).max() <= TOL


def test_large_and_drop_in_names():
    from oracle import render_equation_sh_oracle as RO
    from svgir_b200 import sampling
    torch.manual_seed(3)
    n = torch.nn.functional.normalize(torch.randn(200_000, 3, device="cuda"), dim=-1)
    d, a = sampling.sample_incident_rays(n, is_training=False, sample_num=24)
    do, ao = RO.fibonacci_sphere_sampling(n, 24)
    assert float((d - do).abs().max()) <= TOL and bool((a == ao).all())
    # training: same torch.rand draw as the reference makes (graphics_utils.py:21)
    torch.manual_seed(11)
    d1, _ = sampling.sample_incident_rays(n, is_training=True, sample_num=24)
    torch.manual_seed(11)
    u = torch.rand(200_000, 1, device="cuda")
    d2, _ = RO.fibonacci_sphere_sampling(n, 24, rand_u=u)
    assert float((d1 - d2).abs().max()) <= 2e-5
    e, ea = sampling.fibonacci_sphere_sampling(torch.zeros(0, 3, device="cuda"), 24)
    assert e.shape == (0, 24, 3) and ea.shape == (0, 24, 1)
    with pytest.raises(RuntimeError):
        sampling.fibonacci_sphere_sampling(torch.zeros(4, 3), 24)
