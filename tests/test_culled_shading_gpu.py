"""Shading restricted to the rasteriser's visible-surfel work list (render_view(shade_culled=False)):
every image and every gradient must equal the shade-everything reference order (svgss.py:116-189),
because culled surfels are never composited and receive zero gradient. Also checks the work list itself."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(P=8000, W=176, H=120, Ns=32):
    from svgir_b200 import pipeline, scene
    dev = torch.device("cuda:0")
    cloud = scene.make_surfels(P, seed=31)
    mats = scene.make_materials(cloud, Ns, seed=32, env_hw=(16, 32))
    cam = pipeline.camera_from_scene(scene.look_at_camera(W, H, 1, 4), dev)
    gt = torch.rand(3, H, W, device=dev, generator=torch.Generator(dev).manual_seed(3))
    return pipeline, cloud, mats, cam, gt, dev


def _model(pipeline, cloud, mats, dev):
    pc = pipeline.model_from_scene(cloud, mats, dev)
    env = torch.from_numpy(mats["env_param"]).to(dev).requires_grad_(True)
    return pc, env


def test_visible_work_list_is_the_set_of_positive_radii():
    import svgss_rasterization as sv
    pipeline, cloud, mats, cam, gt, dev = _scene()
    pc, env = _model(pipeline, cloud, mats, dev)
    rs = sv.GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=torch.zeros(3, device=dev), scale_modifier=1.0, viewmatrix=cam.world_view_transform,
        projmatrix=cam.full_proj_transform, patch_bbox=cam.patch_bbox, prcppoint=cam.prcppoint, sh_degree=3,
        campos=cam.camera_center, prefiltered=False, debug=False, config=torch.ones(3, device=dev))
    st, (lst, cnt) = sv.preprocess_geometry(rs, pc.xyz, pc.opacity, pc.scaling, pc.rotation, None, pc.shs, None)
    n = int(cnt.item())
    radii = st.out["radii"]
    want = torch.nonzero(radii > 0).flatten().to(torch.int32)
    assert n == want.numel() and 0 < n < radii.numel()
    assert torch.equal(torch.sort(lst[:n]).values, want)


@pytest.mark.parametrize("is_training", [True, False])
def test_culled_shading_gives_identical_images_and_gradients(is_training):
    from svgir_b200 import shading
    pipeline, cloud, mats, cam, gt, dev = _scene()
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    outs = []
    for culled in (True, False):
        pc, env = _model(pipeline, cloud, mats, dev)
        if is_training:
            res = pipeline.render_view(cam, pc, (env, shading.MODE_LEARNABLE), bg, is_training=True, shade_culled=culled)
            loss = pipeline.image_loss(res, gt) + res["base_color"].mean() + res["roughness"].mean() + res["diffuse"].mean()
            loss.backward()
            grads = [t.grad.clone() for t in pc.trainable() + [env]]
        else:
            with torch.no_grad():
                res = pipeline.render_view(cam, pc, (env, shading.MODE_LEARNABLE), bg, is_training=False, shade_culled=culled)
            grads = []
        outs.append((res, grads))
    (ra, ga), (rb, gb) = outs
    for k in ("render", "pbr", "normal", "base_color", "roughness", "local_lights", "visibility", "opacity", "depth") + \
            (("diffuse",) if is_training else ("lights", "direct", "indirect")):
        assert torch.equal(ra[k], rb[k]), k
    assert int(ra["num_rendered"]) == int(rb["num_rendered"])
    for a, b in zip(ga, gb):
        assert float((a - b).norm()) <= 1e-3 * float(a.norm()) + 1e-12
    if is_training:
        vis = ra["radii"] > 0
        assert torch.equal(ra["diffuse_light"][vis], rb["diffuse_light"][vis])
        assert float(rb["diffuse_light"][~vis].abs().max()) == 0.0
        # per-surfel gradients of culled surfels are exactly zero in both orders
        assert float(ga[5][~vis].abs().max()) == 0.0 and float(gb[5][~vis].abs().max()) == 0.0
