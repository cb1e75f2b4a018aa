"""Fused resolve + loss tail (csrc/resolve.cu via svgir_b200.losses) against the torch mirror of the reference's
tail (pipeline.render_view + pipeline.image_loss: gaussian_renderer/svgss.py:187-233, 280-294): same loss, same
pixel gradients, same parameter gradients after the full backward. fp32; tolerances written below."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _torch_tail(color, geo_normal, opacity, vfeature, gt, bg, lambda_pbr, lambda_normal):
    from svgir_b200 import pipeline
    inv_o = 1.0 / opacity.clamp_min(1e-5)
    vf = vfeature * inv_o
    pbr, shn = vf[0:3], vf[6:9]
    pbr_img = pipeline.rgb_to_srgb(pbr * opacity + (1 - opacity) * bg[:, None, None])
    l1 = (color - gt).abs().mean()
    l1p = (pbr_img - gt).abs().mean()
    nn = (1.0 - (shn * geo_normal).sum(0)).mean()
    return l1 + lambda_pbr * l1p + lambda_normal * nn, (l1, l1p, nn)


@pytest.mark.parametrize("W,H,bgv", [(160, 128, (0.0, 0.0, 0.0)), (333, 77, (0.1, 0.2, 0.3))])
def test_fused_tail_matches_torch(W, H, bgv):
    from svgir_b200 import losses
    dev = torch.device("cuda:0")
    g = torch.Generator(dev).manual_seed(5)
    r = lambda *s: torch.rand(*s, device=dev, generator=g)
    opacity = r(1, H, W)
    opacity[0, :4] = 0.0                      # empty pixels: the clamp_min(1e-5) branch
    opacity[0, 4:8] = 1e-6
    color = r(3, H, W) * 1.2 - 0.1
    geo_normal = (r(3, H, W) - 0.5) * opacity
    vfeature = r(13, H, W) * opacity * 1.3    # some pbr values above 1 -> the srgb clamp
    vfeature[0:3, 8:12] = 1e-4 * opacity[:, 8:12]   # linear branch of rgb_to_srgb
    gt = r(3, H, W)
    bg = torch.tensor(bgv, device=dev)
    ins_t = [t.clone().requires_grad_(True) for t in (color, geo_normal, opacity, vfeature)]
    ins_f = [t.clone().requires_grad_(True) for t in (color, geo_normal, opacity, vfeature)]
    loss_t, (l1, l1p, nn) = _torch_tail(*ins_t, gt, bg, 0.7, 0.02)
    loss_f, terms = losses.fused_train_loss(*ins_f, gt, bg, lambda_pbr=0.7, lambda_normal=0.02)
    (loss_t * 1.5).backward()
    (loss_f * 1.5).backward()
    assert abs(float(loss_f) - float(loss_t)) <= 2e-6 * abs(float(loss_t))
    for a, b in zip(terms[1:].tolist(), (float(l1), float(l1p), float(nn))):
        assert abs(a - b) <= 2e-6 * max(abs(b), 1e-3)
    for name, a, b in zip(("color", "geo_normal", "opacity", "vfeature"), ins_f, ins_t):
        err = (a.grad - b.grad).abs().max().item()
        scale = b.grad.abs().max().item()
        assert err <= 2e-5 * scale, (name, err, scale)   # fp32 re-association only (powf, 1/x)


def test_training_step_fused_matches_torch_tail():
    from svgir_b200 import pipeline, scene
    dev = torch.device("cuda:0")
    cloud = scene.make_surfels(6000, seed=11)
    mats = scene.make_materials(cloud, 16, seed=12, env_hw=(16, 32))
    cam = pipeline.camera_from_scene(scene.look_at_camera(160, 128, 1, 4), dev)
    gt = torch.rand(3, 128, 160, device=dev, generator=torch.Generator(dev).manual_seed(0))
    bg = torch.zeros(3, device=dev)
    out = []
    for fused in (False, True):
        pc = pipeline.model_from_scene(cloud, mats, dev)
        env = torch.from_numpy(mats["env_param"]).to(dev).requires_grad_(True)
        loss, res = pipeline.training_step(cam, pc, env, bg, gt, fused_loss=fused)
        out.append((float(loss), [t.grad.clone() for t in pc.trainable() + [env]], res))
    (l0, g0, r0), (l1, g1, r1) = out
    assert abs(l0 - l1) <= 2e-6 * abs(l0)
    assert torch.equal(r0["render"], r1["render"])
    for a, b in zip(g1, g0):
        rel = float((a - b).norm() / b.norm().clamp_min(1e-20))
        assert rel < 1e-3, rel    # north-star gradient tolerance (atomic order differs run to run)


@pytest.mark.parametrize("bgv", [(0.0, 0.0, 0.0), (0.1, 0.2, 0.3)])
def test_eval_frame_fused_resolve_matches_torch_tail(bgv):
    """render_view(is_training=False): the one-kernel eval resolve (losses.resolve_eval, csrc/resolve.cu) against the
    torch mirror of gaussian_renderer/svgss.py:187-262 on the same rasteriser outputs -- every result image within
    2e-6 absolute (powf vs torch.pow in rgb_to_srgb; everything else is the same fp32 arithmetic)."""
    import numpy as np
    from svgir_b200 import pipeline, scene, shading
    dev = torch.device("cuda:0")
    cloud = scene.make_surfels(5000, seed=21)
    mats = scene.make_materials(cloud, 24, seed=22, env_hw=(16, 32))
    pc = pipeline.model_from_scene(cloud, mats, dev, requires_grad=False)
    cam = pipeline.camera_from_scene(scene.look_at_camera(144, 112, 1, 4), dev)
    env = torch.from_numpy(np.random.default_rng(3).uniform(0, 4, (16, 32, 3)).astype(np.float32)).to(dev)
    bg = torch.tensor(bgv, device=dev)
    old = pipeline.FUSED_RESOLVE
    try:
        with torch.no_grad():
            pipeline.FUSED_RESOLVE = True
            got = pipeline.render_view(cam, pc, (env, shading.MODE_FIXED), bg, is_training=False)
            pipeline.FUSED_RESOLVE = False
            want = pipeline.render_view(cam, pc, (env, shading.MODE_FIXED), bg, is_training=False)
    finally:
        pipeline.FUSED_RESOLVE = old
    keys = ("pbr", "normal", "base_color", "roughness", "lights", "local_lights", "visibility", "direct", "indirect")
    for k in keys:
        assert got[k].shape == want[k].shape, (k, got[k].shape, want[k].shape)
        err = float((got[k] - want[k]).abs().max())
        assert err <= 2e-6 * max(1.0, float(want[k].abs().max())), (k, err)
    for k in ("render", "depth", "opacity", "geo_normal"):
        assert torch.equal(got[k], want[k]), k
    assert float(want["pbr"].max()) > 0.05 and float(want["opacity"].max()) > 0.5   # the frame is not empty
