"""Fused resolve + loss tail (csrc/resolve.cu via svgir_b200.losses) against the torch mirror of the reference's
tail (pipeline.render_view + pipeline.image_loss: gaussian_renderer/svgss.py:187-233, 280-294): same loss, same
pixel gradients, same parameter gradients after the full backward. fp32; tolerances written below."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _torch_tail(color, geo_normal, opacity, vfeature, gt, bg, lambda_pbr, lambda_normal, depth=None, mask=None, cam_terms=None):
    from svgir_b200 import losses, pipeline
    inv_o = 1.0 / opacity.clamp_min(1e-5)
    vf = vfeature * inv_o
    pbr, shn = vf[0:3], vf[6:9]
    pbr_img = pipeline.rgb_to_srgb(pbr * opacity + (1 - opacity) * bg[:, None, None])
    l1 = (color - gt).abs().mean()
    l1p = (pbr_img - gt).abs().mean()
    if depth is None:
        nn = (1.0 - (shn * geo_normal).sum(0)).mean()
    else:   # svgss.py:300-313: cos_loss(rendered_normal, depth2normal(rendered_depth, image_mask, camera))
        H, W = color.shape[-2:]
        nn = losses.cos_loss_torch(shn, losses.depth2normal_torch(depth, mask, H, W, cam_terms))
    return l1 + lambda_pbr * l1p + lambda_normal * nn, (l1, l1p, nn)


@pytest.mark.parametrize("W,H,bgv,mode", [(160, 128, (0.0, 0.0, 0.0), "geo"), (333, 77, (0.1, 0.2, 0.3), "geo"),
                                          (160, 128, (0.0, 0.0, 0.0), "d2n"), (333, 77, (0.1, 0.2, 0.3), "d2n_mask")])
def test_fused_tail_matches_torch(W, H, bgv, mode):
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
    vfeature[6:9] = (r(3, H, W) - 0.5) * opacity    # shading normals of both signs
    gt = r(3, H, W)
    bg = torch.tensor(bgv, device=dev)
    extra_t, extra_f, depth_t, depth_f = {}, {}, None, None
    if mode != "geo":
        ys, xs = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32), torch.arange(W, device=dev, dtype=torch.float32),
                                indexing="ij")
        depth = (3.5 + 0.4 * torch.sin(xs / 7.0) * torch.cos(ys / 5.0) + 0.05 * r(H, W))[None]
        mask = (((xs - W / 2) ** 2 + (ys - H / 2) ** 2) < (0.45 * min(H, W)) ** 2).float()[None] if mode == "d2n_mask" else None
        cam_terms = losses.d2n_camera_terms(H, W, 0.36, 0.36 * H / W, (0.5, 0.5) if mode == "d2n" else (0.47, 0.52))
        depth_t, depth_f = depth.clone().requires_grad_(True), depth.clone().requires_grad_(True)
        extra_t = dict(depth=depth_t, mask=mask, cam_terms=cam_terms)
        extra_f = dict(depth=depth_f, mask=mask, cam_terms=cam_terms)
    ins_t = [t.clone().requires_grad_(True) for t in (color, geo_normal, opacity, vfeature)]
    ins_f = [t.clone().requires_grad_(True) for t in (color, geo_normal, opacity, vfeature)]
    loss_t, (l1, l1p, nn) = _torch_tail(*ins_t, gt, bg, 0.7, 0.02, **extra_t)
    loss_f, terms = losses.fused_train_loss(*ins_f, gt, bg, lambda_pbr=0.7, lambda_normal=0.02, **extra_f)
    (loss_t * 1.5).backward()
    (loss_f * 1.5).backward()
    assert abs(float(loss_f) - float(loss_t)) <= 2e-6 * abs(float(loss_t))
    for a, b in zip(terms[1:4].tolist(), (float(l1), float(l1p), float(nn))):
        assert abs(a - b) <= 2e-6 * max(abs(b), 1e-3)
    names = ["color", "geo_normal", "opacity", "vfeature"]
    if mode != "geo":
        ins_f, ins_t, names = ins_f + [depth_f], ins_t + [depth_t], names + ["depth"]
    for name, a, b in zip(names, ins_f, ins_t):
        if mode != "geo" and name == "geo_normal":
            assert b.grad is None and float(a.grad.abs().max()) == 0.0   # the surface term does not read it
            continue
        err = (a.grad - b.grad).abs().max().item()
        scale = b.grad.abs().max().item()
        # fp32 re-association only (powf, 1/x; the depth gradient is a sum of five atomically added stencil terms)
        assert err <= (2e-5 if name != "depth" else 2e-4) * scale, (name, err, scale)


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
