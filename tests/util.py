"""Shared helpers of the parity tests: seeded scenes, the three implementations side by side."""
from __future__ import annotations

import numpy as np
import torch

from svgir_b200 import scene


def make_case(P, W, H, S=4, VS=52, seed=1, view=0, use_sh=True):
    cl = scene.make_surfels(P, seed=seed)
    cam = scene.look_at_camera(W, H, view)
    rng = np.random.default_rng(seed + 100)
    feat = rng.uniform(0, 1, (P, S)).astype(np.float32)
    vfeat = rng.uniform(0, 1, (P, VS)).astype(np.float32)
    colors = None if use_sh else rng.uniform(0, 1, (P, 3)).astype(np.float32)
    return dict(cloud=cl, cam=cam, features=feat, vfeatures=vfeat, colors=colors, S=S, VS=VS,
                bg=np.array([0.1, 0.2, 0.3], np.float32), config=np.array([1, 1, 1], np.float32))


def pixel_grads(case, seed=7):
    rng = np.random.default_rng(seed)
    H, W = case["cam"].H, case["cam"].W
    n = H * W
    mk = lambda c: (rng.standard_normal((c, H, W)) / n).astype(np.float32)
    return dict(dL_dcolor=mk(3), dL_dnormal=mk(3), dL_ddepth=mk(1), dL_dopacity=mk(1),
                dL_dfeature=mk(case["S"]), dL_dvfeature=mk(case["VS"] // 4))


def run_oracle(case, backward=True, grads=None):
    from oracle import svgss as O
    cl, cam = case["cloud"], case["cam"]
    fw = O.forward(cam, cl.means3D, cl.opacity, cl.scales, cl.rotations, case["features"], case["vfeatures"],
                   shs=cl.shs if case["colors"] is None else None, colors_precomp=case["colors"],
                   bg=case["bg"], config=case["config"])
    bw = None
    if backward:
        g = grads or pixel_grads(case)
        bw = O.backward(fw, g["dL_dcolor"], g["dL_dnormal"], g["dL_ddepth"], g["dL_dopacity"],
                        g["dL_dfeature"], g["dL_dvfeature"])
    return fw, bw


def to_cuda(case):
    cl, cam = case["cloud"], case["cam"]
    d = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return dict(means3D=d(cl.means3D), opacity=d(cl.opacity), scales=d(cl.scales), rotations=d(cl.rotations),
                shs=d(cl.shs) if case["colors"] is None else None, colors=d(case["colors"]),
                features=d(case["features"]), vfeatures=d(case["vfeatures"]), bg=d(case["bg"]),
                viewmatrix=d(cam.viewmatrix), projmatrix=d(cam.projmatrix), campos=d(cam.campos),
                patch_bbox=d(cam.patch_bbox), prcppoint=d(cam.prcppoint), config=d(case["config"]))


def run_ours(case, backward=True, grads=None, want_sorted_keys=True, debug=False):
    """Calls the C ABI through svgir_b200.raster. Returns (out dict of torch tensors, state, bwd dict)."""
    from svgir_b200 import raster
    cam = case["cam"]
    t = to_cuda(case)
    s = raster.RasterSettings(image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                              bg=t["bg"], scale_modifier=1.0, viewmatrix=t["viewmatrix"],
                              projmatrix=t["projmatrix"], sh_degree=3, campos=t["campos"],
                              patch_bbox=t["patch_bbox"], config=t["config"], debug=debug)
    out, st = raster.forward(s, t["means3D"], t["opacity"], t["scales"], t["rotations"], None, t["shs"],
                             t["colors"], t["features"], t["vfeatures"], want_sorted_keys=want_sorted_keys)
    bw = None
    if backward:
        g = grads or pixel_grads(case)
        gt = {k: torch.from_numpy(v).cuda() for k, v in g.items()}
        bw = raster.backward(st, out["radii"], gt, want_debug=True)
    torch.cuda.synchronize()
    return out, st, bw


def run_ref(case, backward=True, grads=None):
    """Unmodified reference CUDA rasteriser (oracle/_ref). Returns (RefSvgss, out, bwd)."""
    from oracle import ref_cuda
    cam = case["cam"]
    t = to_cuda(case)
    r = ref_cuda.RefSvgss()
    P = t["means3D"].shape[0]
    z = lambda *s: torch.zeros(s, device="cuda")
    out = r.forward(bg=t["bg"], means3D=t["means3D"],
                    features=t["features"] if t["features"] is not None else z(P, 0),
                    vfeatures=t["vfeatures"] if t["vfeatures"] is not None else z(P, 0),
                    colors=t["colors"], opacity=t["opacity"], scales=t["scales"], rotations=t["rotations"],
                    scale_modifier=1.0, viewmatrix=t["viewmatrix"], projmatrix=t["projmatrix"],
                    prcppoint=t["prcppoint"], patchbbox=t["patch_bbox"], tanfovx=cam.tanfovx,
                    tanfovy=cam.tanfovy, H=cam.H, W=cam.W, sh=t["shs"], degree=3, campos=t["campos"],
                    config=t["config"])
    bw = None
    if backward:
        g = grads or pixel_grads(case)
        gt = {k: torch.from_numpy(v).cuda() for k, v in g.items()}
        bw = r.backward(gt["dL_dcolor"], gt["dL_dnormal"], gt["dL_ddepth"], gt["dL_dopacity"],
                        gt["dL_dfeature"], gt["dL_dvfeature"])
    return r, out, bw


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    d = np.linalg.norm(a - b)
    n = max(np.linalg.norm(b), 1e-30)
    return d / n


def make_bvh_case(P, n_rays, seed=11, surfel=True, size=0.08):
    """Random occluder cloud + random rays for the LBVH / visibility tests: centres uniform in the unit
    cube, random orientations, disc-like scales (tiny z when `surfel`), rays from random points."""
    rng = np.random.default_rng(seed)
    means = rng.uniform(-1, 1, (P, 3)).astype(np.float32)
    q = rng.standard_normal((P, 4)).astype(np.float32)  # un-normalised on purpose (build_rotation normalises)
    sxy = np.exp(rng.normal(np.log(size), 0.4, (P, 2)))
    sz = np.full((P, 1), 1e-4) if surfel else np.exp(rng.normal(np.log(0.03), 0.3, (P, 1)))
    scales = np.concatenate([sxy, sz], 1).astype(np.float32)
    opacity = (1 / (1 + np.exp(-rng.normal(0.5, 2.0, (P,))))).astype(np.float32)
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    r, x, y, z = qn.T
    normals = np.stack([2 * (x * z + r * y), 2 * (y * z - r * x), 1 - 2 * (x * x + y * y)], 1).astype(np.float32)
    ro = rng.uniform(-1, 1, (n_rays, 3)).astype(np.float32)
    rd = rng.standard_normal((n_rays, 3))
    rd = (rd / np.linalg.norm(rd, axis=1, keepdims=True)).astype(np.float32)
    return dict(means=means, scales=scales, rotations=q, opacity=opacity, normals=normals, rays_o=ro, rays_d=rd)
