"""Times the radiance-cache build and the radiance-consistency loss at the training shape (300k surfels x 64 samples).
usage: python tools/time_radiance.py [P] [S]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "svg-ir_b200"))
from svgir_b200 import bvh, radiance, sampling, scene  # noqa: E402


def quat_to_R(q):
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q.unbind(1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, out


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    dev = torch.device("cuda:0")
    cloud = scene.make_surfels(P, seed=1)
    t = lambda a: torch.from_numpy(a).to(dev)
    xyz, scaling, rot, opac, shs, gn = t(cloud.means3D), t(cloud.scales), t(cloud.rotations), t(cloud.opacity), t(cloud.shs), t(cloud.normals)
    scaling = scaling.clone()
    scaling[:, 2] = 1e-3
    R = quat_to_R(rot)
    Sinv = R @ torch.diag_embed(1.0 / scaling ** 2) @ R.transpose(1, 2)
    ci = torch.stack([Sinv[:, 0, 0], Sinv[:, 0, 1], Sinv[:, 0, 2], Sinv[:, 1, 1], Sinv[:, 1, 2], Sinv[:, 2, 2]], 1).contiguous()
    res = {"P": P, "S": S}
    ms, tracer = timed(lambda: bvh.RayTracer(xyz, scaling, rot), 3)
    res["tree_build_ms"] = ms
    ms, rec = timed(lambda: radiance.pack_surfels(xyz, scaling, rot, gn, opac, ci))
    res["pack_ms"] = ms
    u = torch.rand(P, 1, device=dev, generator=torch.Generator(dev).manual_seed(2))
    dirs, areas = sampling.fibonacci_sphere_sampling(gn, S, random_rotate=True, rand_u=u)
    ms, (rad, vis, hit, uv) = timed(lambda: radiance.render_radiance_with_sampling_SH(tracer.bvh, rec, shs, xyz, dirs, S), 3)
    res["cache_build_ms"] = ms
    res["rays_per_s"] = P * S / (ms * 1e-3)
    res["hit_fraction"] = float((hit >= 0).float().mean())
    res["occluded_fraction"] = float((vis == 0).float().mean())
    g = torch.Generator(dev).manual_seed(3)
    n12 = torch.randn(P, 12, device=dev, generator=g)
    alb = torch.rand(P, 12, device=dev, generator=g).requires_grad_(True)
    rough = (0.1 + 0.8 * torch.rand(P, 4, device=dev, generator=g)).requires_grad_(True)
    env = torch.randn(16, 32, 3, device=dev, generator=g).requires_grad_(True)
    cam = torch.tensor([0.0, 0.0, 4.0], device=dev)
    ratio = torch.tensor(1.0, device=dev)

    def fwd():
        return radiance.radiance_loss(cam, (env, 0), xyz, gn, dirs, areas, vis, hit, uv, rad, ratio, n12, alb, rough, return_aux=True)

    ms_f, (loss, irr, sel) = timed(fwd, 10)
    res["loss_forward_ms"] = ms_f

    def fb():
        l = radiance.radiance_loss(cam, (env, 0), xyz, gn, dirs, areas, vis, hit, uv, rad, ratio, n12, alb, rough)
        alb.grad = rough.grad = env.grad = None
        l.backward()
        return l

    ms_fb, _ = timed(fb, 10)
    res["loss_forward_backward_ms"] = ms_fb
    env.requires_grad_(False)
    ms_fb2, _ = timed(fb, 10)
    res["loss_forward_backward_no_env_grad_ms"] = ms_fb2
    env.requires_grad_(True)
    res["loss"] = float(loss)
    res["selected_hit_fraction"] = float((hit[torch.arange(P, device=dev), sel.long(), 0] >= 0).float().mean())
    # algorithmic bytes of the loss forward: per surfel dirs+vis of its own row (S*16 B) and, when the selected sample hit,
    # the hit surfel's dirs/hit/uv/areas rows (S*28 B) + 100 B of materials
    print(json.dumps(res))


if __name__ == "__main__":
    main()
