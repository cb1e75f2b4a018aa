"""CUDA loss-tail kernels through the C ABI against golden vectors from the reference's own functions
(tests/golden/ref_losses.npz): the surface term cos_loss(normal, depth2normal(depth)) inside the fused loss
(csrc/resolve.cu), the edge-aware smoothness and the TV loss (csrc/loss_terms.cu); and at 800x800 against the torch
mirror. fp32; tolerances written below."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = dict(np.load(os.path.join(HERE, "golden", "ref_losses.npz")))


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_surface_term_matches_reference_golden(tag):
    """fused_train_loss with lambda_pbr = 0 and a colour image equal to the ground truth leaves exactly
    lambda_normal * cos_loss(normal, depth2normal(depth, mask, camera)); opacity = 1 makes the un-premultiply the identity."""
    from svgir_b200 import losses
    dev = torch.device("cuda:0")
    H, W, fovx, fovy, px, py = G[f"{tag}_cam"]
    H, W = int(H), int(W)
    terms = losses.d2n_camera_terms(H, W, math.tan(fovx / 2), math.tan(fovy / 2), (px, py))
    depth = torch.from_numpy(G[f"{tag}_depth"]).to(dev).requires_grad_(True)
    normal = torch.from_numpy(G[f"{tag}_normal"]).to(dev)
    mask = torch.from_numpy(G[f"{tag}_mask"]).to(dev)
    vf = torch.zeros(13, H, W, device=dev)
    vf[6:9] = normal
    vf.requires_grad_(True)
    gt = torch.rand(3, H, W, device=dev)
    loss, t = losses.fused_train_loss(gt.clone(), torch.zeros(3, H, W, device=dev), torch.ones(1, H, W, device=dev), vf, gt,
                                      torch.zeros(3, device=dev), lambda_pbr=0.0, lambda_normal=1.0, depth=depth, mask=mask,
                                      cam_terms=terms)
    loss.backward()
    assert abs(float(t[3]) - float(G[f"{tag}_cos_loss"])) < 3e-6, (float(t[3]), float(G[f"{tag}_cos_loss"]))
    assert abs(float(loss) - float(G[f"{tag}_cos_loss"])) < 3e-6
    assert _rel(vf.grad[6:9].cpu(), torch.from_numpy(G[f"{tag}_g_normal"])) < 1e-4
    assert _rel(depth.grad.cpu(), torch.from_numpy(G[f"{tag}_g_depth"])) < 1e-3     # north-star gradient tolerance


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_edge_aware_matches_reference_golden(tag):
    from svgir_b200 import losses
    dev = torch.device("cuda:0")
    data = torch.from_numpy(G[f"{tag}_ea_data"]).to(dev).requires_grad_(True)
    img = torch.from_numpy(G[f"{tag}_ea_img"]).to(dev)
    mask = torch.from_numpy(G[f"{tag}_mask"]).to(dev)
    v = losses.fused_edge_aware(data, img, mask)
    v.backward()
    assert abs(float(v) - float(G[f"{tag}_ea_loss"])) < 1e-6
    ref = torch.from_numpy(G[f"{tag}_ea_grad"])
    assert float((data.grad.cpu() - ref).abs().max()) <= 2e-6 * float(ref.abs().max()) + 1e-9


def test_tv_matches_reference_golden_in_place_on_hwc_env():
    from svgir_b200 import losses
    dev = torch.device("cuda:0")
    env = torch.from_numpy(G["tv_env"]).to(dev).requires_grad_(True)
    v = losses.fused_tv(env[0].permute(2, 0, 1))          # svgss.py:388, read through its strides
    (2.0 * v).backward()
    assert abs(float(v) - float(G["tv_loss"])) < 1e-5
    assert float((env.grad.cpu() - 2.0 * torch.from_numpy(G["tv_grad"])).abs().max()) < 1e-6


def test_loss_terms_full_size_vs_torch_mirror():
    """800x800 (the bench image size): kernels vs losses.*_torch on the same device."""
    from svgir_b200 import losses
    dev = torch.device("cuda:0")
    g = torch.Generator(dev).manual_seed(3)
    H = W = 800
    data = torch.rand(3, H, W, device=dev, generator=g)
    img = torch.rand(3, H, W, device=dev, generator=g)
    a, b = data.clone().requires_grad_(True), data.clone().requires_grad_(True)
    la, lb = losses.fused_edge_aware(a, img), losses.edge_aware_torch(b, img)
    la.backward(); lb.backward()
    assert abs(float(la) - float(lb)) < 2e-6 * abs(float(lb))
    assert _rel(a.grad, b.grad) < 1e-5
    assert float(losses.fused_edge_aware(data, img)) == float(losses.fused_edge_aware(data, img))   # deterministic forward
    # surface term at full size
    ys, xs = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32), torch.arange(W, device=dev, dtype=torch.float32), indexing="ij")
    depth = (3.5 + 0.4 * torch.sin(xs / 17.0) * torch.cos(ys / 13.0) + 0.02 * torch.rand(H, W, device=dev, generator=g))[None]
    shn = torch.nn.functional.normalize(torch.randn(3, H, W, device=dev, generator=g), dim=0) * 0.8
    terms = losses.d2n_camera_terms(H, W, 0.36, 0.36)
    da, db = depth.clone().requires_grad_(True), depth.clone().requires_grad_(True)
    vf = torch.zeros(13, H, W, device=dev)
    vf[6:9] = shn
    gt = torch.rand(3, H, W, device=dev, generator=g)
    lf, t = losses.fused_train_loss(gt.clone(), torch.zeros(3, H, W, device=dev), torch.ones(1, H, W, device=dev), vf, gt,
                                    torch.zeros(3, device=dev), lambda_pbr=0.0, lambda_normal=1.0, depth=da, cam_terms=terms)
    lt = losses.cos_loss_torch(shn, losses.depth2normal_torch(db, None, H, W, terms))
    lf.backward(); lt.backward()
    assert abs(float(lf) - float(lt)) < 5e-6 * abs(float(lt))
    assert _rel(da.grad, db.grad) < 1e-3
