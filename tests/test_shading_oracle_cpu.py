"""CPU: the shading restatement (oracle/shading_oracle.py) against golden vectors produced by the
reference's own Python code (tests/golden/make_golden_shading.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import shading_oracle as SO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(tag):
    return dict(np.load(os.path.join(GOLD, f"ref_shading_{tag}.npz")))


@pytest.mark.parametrize("tag", ["tiny", "train_small", "eval_small"])
def test_rendering_equation4_matches_reference(tag):
    g = load(tag)
    names = ("base_color", "roughness", "shading_normals", "viewdirs", "radiance", "env_param")
    t = {k[3:]: torch.tensor(v) for k, v in g.items() if k.startswith("in_")}
    for k in names:
        t[k].requires_grad_(True)
    pbr, extra = SO.rendering_equation4(
        t["base_color"], t["roughness"], t["shading_normals"], t["viewdirs"], t["radiance"],
        lambda d: SO.direct_light_learnable(t["env_param"], d), t["visibility"], t["incident_dirs"],
        t["incident_areas"])
    outs = dict(pbr=pbr, diffuse_light=extra["diffuse_light"], specular=extra["specular"],
                direct=extra["direct"], indirect=extra["indirect"])
    for k, v in outs.items():
        np.testing.assert_allclose(v.detach().numpy(), g["out_" + k], rtol=2e-5, atol=2e-6, err_msg=k)
    np.testing.assert_allclose(extra["incident_lights"].detach().numpy(), g["out_incident_lights"], rtol=1e-5, atol=1e-6)
    loss = sum((outs[k] * torch.tensor(g["cot_" + k])).sum() for k in outs)
    loss.backward()
    for k in names:
        ref = g["grad_" + k]
        got = t[k].grad.numpy()
        err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
        assert err < 1e-4, (k, err)


def test_env_lookups_match_reference():
    g = dict(np.load(os.path.join(GOLD, "ref_envlight.npz")))
    d = torch.tensor(g["dirs"])
    out = SO.direct_light_learnable(torch.tensor(g["env_param"]), d).numpy()
    np.testing.assert_allclose(out, g["out_learnable"], rtol=1e-5, atol=1e-5)
    out = SO.direct_light_hdr(torch.tensor(g["hdr"]), d).numpy()
    np.testing.assert_allclose(out, g["out_hdr"], rtol=1e-5, atol=1e-5)
    out = SO.direct_light_hdr(torch.tensor(g["hdr"]), d, torch.tensor(g["transform"])).numpy()
    np.testing.assert_allclose(out, g["out_hdr_transformed"], rtol=1e-5, atol=1e-5)


def test_explicit_bilinear_equals_grid_sample():
    torch.manual_seed(0)
    env = torch.rand(3, 8, 16)
    gx = torch.rand(500) * 2.4 - 1.2
    gy = torch.rand(500) * 2.4 - 1.2
    ref = torch.nn.functional.grid_sample(env[None], torch.stack([gx, gy], -1)[None, None], align_corners=True)[0, :, 0].t()
    np.testing.assert_allclose(SO.grid_sample_bilinear_zeros(env, gx, gy).numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)
