"""The loss-tail oracle (oracle/loss_oracle.py, numpy) and the product's host-side torch mirror
(svgir_b200.losses.*_torch, used by pipeline.image_loss) against golden vectors produced by the reference's OWN
functions (tests/golden/make_golden_losses.py: depth2normal, cos_loss, first_order_edge_aware_loss, tv_loss + torch
autograd for the gradients). fp32 references; tolerances written below."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
G = dict(np.load(os.path.join(HERE, "golden", "ref_losses.npz")))
TAGS = ("a", "b", "c")


def _cam(tag):
    H, W, fovx, fovy, px, py = G[f"{tag}_cam"]
    return int(H), int(W), float(fovx), float(fovy), (float(px), float(py))


def test_oracle_depth2normal_and_cos_loss_match_reference():
    from oracle import loss_oracle as LO
    for tag in TAGS:
        H, W, fovx, fovy, pp = _cam(tag)
        n = LO.depth2normal(G[f"{tag}_depth"], G[f"{tag}_mask"], H, W, fovx, fovy, pp)
        # the normalised cross products amplify fp32 rounding where the depth differences nearly cancel: 2e-4 abs on unit vectors
        assert np.abs(n - G[f"{tag}_d2n"]).max() < 2e-4, tag
        assert abs(LO.cos_loss(G[f"{tag}_normal"], G[f"{tag}_d2n"]) - float(G[f"{tag}_cos_loss"])) < 2e-6, tag


def test_oracle_edge_aware_and_tv_match_reference():
    from oracle import loss_oracle as LO
    for tag in TAGS:
        m = G[f"{tag}_mask"]
        v = LO.first_order_edge_aware_loss(G[f"{tag}_ea_data"] * m, G[f"{tag}_ea_img"] * m)
        assert abs(v - float(G[f"{tag}_ea_loss"])) < 2e-7 + 2e-6 * abs(v), tag
    env = G["tv_env"]
    assert abs(LO.tv_loss(np.transpose(env[0], (2, 0, 1))) - float(G["tv_loss"])) < 1e-5


def test_torch_mirror_matches_reference_values_and_gradients():
    """svgir_b200.losses.depth2normal_torch / cos_loss_torch / edge_aware_torch / tv_torch: the host-side mirror the
    un-fused tail (pipeline.image_loss) runs and the GPU tests check the kernels against."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "svg-ir_b200"))
    from svgir_b200 import losses
    import math
    for tag in TAGS:
        H, W, fovx, fovy, pp = _cam(tag)
        terms = losses.d2n_camera_terms(H, W, math.tan(fovx / 2), math.tan(fovy / 2), pp)
        depth = torch.from_numpy(G[f"{tag}_depth"]).requires_grad_(True)
        normal = torch.from_numpy(G[f"{tag}_normal"]).requires_grad_(True)
        mask = torch.from_numpy(G[f"{tag}_mask"])
        d2n = losses.depth2normal_torch(depth, mask, H, W, terms)
        assert float((d2n.detach() - torch.from_numpy(G[f"{tag}_d2n"])).abs().max()) < 2e-4, tag
        loss = losses.cos_loss_torch(normal, d2n)
        loss.backward()
        assert abs(float(loss) - float(G[f"{tag}_cos_loss"])) < 2e-6, tag
        for got, key in ((depth.grad, "g_depth"), (normal.grad, "g_normal")):
            ref = torch.from_numpy(G[f"{tag}_{key}"])
            assert float((got - ref).norm() / ref.norm()) < 1e-4, (tag, key)
        data = torch.from_numpy(G[f"{tag}_ea_data"]).requires_grad_(True)
        le = losses.edge_aware_torch(data * mask, torch.from_numpy(G[f"{tag}_ea_img"]) * mask)
        le.backward()
        assert abs(float(le) - float(G[f"{tag}_ea_loss"])) < 1e-6, tag
        ref = torch.from_numpy(G[f"{tag}_ea_grad"])
        assert float((data.grad - ref).abs().max()) <= 1e-6 * float(ref.abs().max()) + 1e-9, tag
    env = torch.from_numpy(G["tv_env"]).requires_grad_(True)
    lt = losses.tv_torch(env[0].permute(2, 0, 1))
    lt.backward()
    assert abs(float(lt) - float(G["tv_loss"])) < 1e-5
    assert float((env.grad - torch.from_numpy(G["tv_grad"])).abs().max()) < 1e-7
