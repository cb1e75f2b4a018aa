"""fused_step.FusedTrainStep (the step as a fixed sequence of C-ABI calls: no autograd between the kernels, gradients
ADDED into one flat buffer for the visible surfels only, binning / parameter backward on a side stream) must give what
`pipeline.training_step` -- the autograd mirror of the reference's render_view + loss + backward -- gives: same loss
and bit-identical images (both evaluate the view direction inside the shading kernel), every parameter gradient
within the atomic-order tolerance, also when several views accumulate and when a replay overflows its bins."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(P=6000, W=160, H=128, Ns=16, n_views=4):
    from svgir_b200 import pipeline, scene
    dev = torch.device("cuda:0")
    cloud = scene.make_surfels(P, seed=31)
    mats = scene.make_materials(cloud, Ns, seed=32, env_hw=(16, 32))
    cams = [pipeline.camera_from_scene(scene.look_at_camera(W, H, v, n_views), dev) for v in range(n_views)]
    gts = [torch.rand(3, H, W, device=dev, generator=torch.Generator(dev).manual_seed(40 + i)) for i in range(2)]
    return pipeline, cloud, mats, cams, gts, dev


def _model(pipeline, cloud, mats, dev):
    pc = pipeline.model_from_scene(cloud, mats, dev)
    env = torch.from_numpy(mats["env_param"]).to(dev).requires_grad_(True)
    return pc, env


NAMES = ("xyz", "opacity", "scaling", "rotation", "shs", "base_color", "roughness", "shading_normal", "env")


def _grads(pc, env):
    return [t.grad.detach().clone() for t in pc.trainable() + [env]]


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def _check_grads(got, want, tol=1e-3):
    for n, a, b in zip(NAMES, got, want):
        assert a.shape == b.shape, n
        assert _rel(a, b) < tol, (n, _rel(a, b))


def test_fused_step_eager_matches_autograd():
    """No CUDA graph: FusedTrainStep.enqueue() launch by launch on the current + side stream."""
    from svgir_b200 import fused_step
    pipeline, cloud, mats, cams, gts, dev = _setup()
    bg = torch.zeros(3, device=dev)
    pc_e, env_e = _model(pipeline, cloud, mats, dev)
    pc_f, env_f = _model(pipeline, cloud, mats, dev)
    cam = pipeline.blocked_camera(cams[0].image_height, cams[0].image_width, cams[0].tanfovx, cams[0].tanfovy,
                                  cams[0].world_view_transform, cams[0].full_proj_transform, cams[0].camera_center,
                                  cams[0].patch_bbox, cams[0].prcppoint, device=dev)
    gt = gts[0].clone()
    fs = fused_step.FusedTrainStep(pc_f, env_f, bg, cam, gt)
    fs.calibrate()
    for i in (0, 2, 1):
        cam.block.copy_(cams[i].block)
        gt.copy_(gts[i % 2])
        loss_f = fs.enqueue()
        torch.cuda.synchronize()
        R, overflow = fs.read_count()
        assert not overflow
        loss_e, res_e = pipeline.training_step(cams[i], pc_e, env_e, bg, gts[i % 2])
        assert R == int(res_e["num_rendered"])
        assert abs(float(loss_f) - float(loss_e)) <= 2e-6 * abs(float(loss_e))
        assert torch.equal(fs.result["render"], res_e["render"])
        assert torch.equal(fs.result["radii"], res_e["radii"])
        assert torch.equal(fs.result["raw_vfeature"], res_e["raw_vfeature"])
        torch.testing.assert_close(fs.result["weights"], res_e["weights"], rtol=1e-4, atol=1e-6)
        _check_grads(_grads(pc_f, env_f), _grads(pc_e, env_e))
        # the screen-space gradient of the densification statistic (means2D.grad in the reference)
        want2d = res_e["viewspace_points"].grad
        assert want2d is not None and _rel(fs.result["viewspace_grad"], want2d) < 1e-3
        # culled surfels: exactly zero everywhere
        culled = ~fs.result["visibility_filter"]
        for n, g in zip(NAMES[:-1], _grads(pc_f, env_f)[:-1]):
            assert float(g[culled].abs().sum()) == 0.0, n
    assert fs.launches >= 10


def test_fused_step_accumulates_views():
    """zero_grads=False: two views added into one bucket == the sum of the two autograd gradients."""
    from svgir_b200 import dist as svdist
    pipeline, cloud, mats, cams, gts, dev = _setup()
    bg = torch.zeros(3, device=dev)
    pc_e, env_e = _model(pipeline, cloud, mats, dev)
    pc_f, env_f = _model(pipeline, cloud, mats, dev)
    bucket = svdist.FlatGradBucket(pc_f.trainable() + [env_f])
    runner = pipeline.GraphedTrainingStep(pc_f, env_f, bg, cams[0], gts[0], bucket=bucket, zero_in_graph=False, fused=True)
    want = None
    for v in (1, 3):
        pipeline.training_step(cams[v], pc_e, env_e, bg, gts[v % 2])
        g = _grads(pc_e, env_e)
        want = g if want is None else [a + b for a, b in zip(want, g)]
    bucket.zero()
    for v in (1, 3):
        runner(cams[v], gts[v % 2])
    torch.cuda.synchronize()
    _check_grads(_grads(pc_f, env_f), want)
    assert runner.captures == 1


def test_fused_overflowed_replay_adds_nothing():
    """in accumulate mode a replay whose binning overflowed must contribute exactly zero, so that
    re-capture + re-run yields the complete sum. Forced here with a capacity of R + 16 and 1.6x larger surfels."""
    from svgir_b200 import dist as svdist, raster
    pipeline, cloud, mats, cams, gts, dev = _setup()
    bg = torch.zeros(3, device=dev)
    pc_e, env_e = _model(pipeline, cloud, mats, dev)
    pc_f, env_f = _model(pipeline, cloud, mats, dev)
    old = (raster.ASYNC_SLACK, raster.ASYNC_MARGIN)
    raster.ASYNC_SLACK, raster.ASYNC_MARGIN = 1.0, 16
    raster._CAP_HINT.clear()
    try:
        bucket = svdist.FlatGradBucket(pc_f.trainable() + [env_f])
        runner = pipeline.GraphedTrainingStep(pc_f, env_f, bg, cams[0], gts[0], bucket=bucket, zero_in_graph=False, fused=True)
        bucket.zero()
        runner(cams[0], gts[0])
        torch.cuda.synchronize()
        first = bucket.flat.clone()
        with torch.no_grad():
            pc_f.scaling.mul_(1.6)
        # the overflowed replay alone (no check): the bucket must be unchanged
        runner.load_inputs(cams[1], gts[1])
        runner.graph.replay()
        torch.cuda.synchronize()
        assert runner.fs.read_count()[1], "the test must overflow the bins"
        assert torch.equal(bucket.flat, first)
        # finish() re-captures and re-runs: first view + second view
        runner.finish()
        torch.cuda.synchronize()
        assert runner.captures == 2
        pipeline.training_step(cams[0], pc_e, env_e, bg, gts[0])
        want = _grads(pc_e, env_e)
        with torch.no_grad():
            pc_e.scaling.mul_(1.6)
        pipeline.training_step(cams[1], pc_e, env_e, bg, gts[1])
        want = [a + b for a, b in zip(want, _grads(pc_e, env_e))]
        _check_grads(_grads(pc_f, env_f), want)
    finally:
        raster.ASYNC_SLACK, raster.ASYNC_MARGIN = old
        raster._CAP_HINT.clear()


def test_prefetched_inputs_match_inline():
    """GraphedTrainingStep.prefetch / replay_prefetched (inputs uploaded on a copy stream one step ahead) == __call__."""
    pipeline, cloud, mats, cams, gts, dev = _setup()
    bg = torch.zeros(3, device=dev)
    pc_a, env_a = _model(pipeline, cloud, mats, dev)
    pc_b, env_b = _model(pipeline, cloud, mats, dev)
    ra = pipeline.GraphedTrainingStep(pc_a, env_a, bg, cams[0], gts[0])
    rb = pipeline.GraphedTrainingStep(pc_b, env_b, bg, cams[0], gts[0])
    host_cams = [pipeline.blocked_camera(c.image_height, c.image_width, c.tanfovx, c.tanfovy, c.world_view_transform.cpu(),
                                         c.full_proj_transform.cpu(), c.camera_center.cpu(), c.patch_bbox.cpu(),
                                         c.prcppoint.cpu(), pin=True) for c in cams]
    host_gts = [g.cpu().pin_memory() for g in gts]
    order = (0, 3, 1, 2)
    rb.prefetch(host_cams[order[0]], host_gts[0])
    for k, v in enumerate(order):
        la, _ = ra(cams[v], gts[k % 2])
        rb.replay_prefetched()
        if k + 1 < len(order):
            rb.prefetch(host_cams[order[k + 1]], host_gts[(k + 1) % 2])
        rb.finish()
        assert float(rb.loss) == float(la)
        assert torch.equal(rb.res["render"], ra.res["render"])


def _quat_R(q):
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q.unbind(1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)


def test_fused_step_carries_the_radiance_consistency_term():
    """FusedTrainStep(radiance_cache=...) = training_step + lambda_radiance * get_radiance_loss (svgss.py:319-320): the
    term's kernels run on the side stream and add into the same gradient buffers."""
    from svgir_b200 import fused_step, radiance
    pipeline, cloud, mats, cams, gts, dev = _setup(P=6000, Ns=16)
    bg = torch.zeros(3, device=dev)
    pc_e, env_e = _model(pipeline, cloud, mats, dev)
    pc_f, env_f = _model(pipeline, cloud, mats, dev)
    with torch.no_grad():
        R = _quat_R(pc_f.rotation)
        sc = pc_f.scaling.clone()
        sc[:, 2] = 1e-3
        Sinv = R @ torch.diag_embed(1.0 / sc ** 2) @ R.transpose(1, 2)
        ci = torch.stack([Sinv[:, 0, 0], Sinv[:, 0, 1], Sinv[:, 0, 2], Sinv[:, 1, 1], Sinv[:, 1, 2], Sinv[:, 2, 2]], 1).contiguous()
        gn = R[:, :, 2].contiguous()
        torch.manual_seed(5)
        rc = radiance.RadianceCache().update(pc_f.xyz, sc, pc_f.rotation, pc_f.opacity, gn, ci, pc_f.shs, sample_num=16)
        rc.radiances = torch.rand_like(rc.radiances)              # a target the irradiance cannot already match
    assert float((rc.hemi_index_buffers >= 0).float().mean()) > 0.01
    lam = 0.05
    cam = pipeline.blocked_camera(cams[0].image_height, cams[0].image_width, cams[0].tanfovx, cams[0].tanfovy,
                                  cams[0].world_view_transform, cams[0].full_proj_transform, cams[0].camera_center,
                                  cams[0].patch_bbox, cams[0].prcppoint, device=dev)
    gt = gts[0].clone()
    fs = fused_step.FusedTrainStep(pc_f, env_f, bg, cam, gt, radiance_cache=rc, lambda_radiance=lam)
    fs.calibrate()
    for i in (1, 3):
        cam.block.copy_(cams[i].block)
        loss_f = fs.enqueue()
        torch.cuda.synchronize()
        assert not fs.read_count()[1]
        loss_e, _ = pipeline.training_step(cams[i], pc_e, env_e, bg, gts[0])
        lr = rc.loss(cams[i].camera_center, (env_e, 0), pc_e.xyz, gn, pc_e.shading_normal, pc_e.base_color, pc_e.roughness)
        (lam * lr).backward()
        assert abs(float(loss_f) - float(loss_e)) <= 2e-6 * abs(float(loss_e))
        assert abs(float(fs.result["loss_radiance"]) - float(lr)) <= 1e-6 * abs(float(lr))
        _check_grads(_grads(pc_f, env_f), _grads(pc_e, env_e))
        # the term reaches surfels this view culls
        culled = ~fs.result["visibility_filter"]
        assert float(pc_f.base_color.grad[culled].abs().sum()) > 0.0
