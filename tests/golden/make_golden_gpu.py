"""Generates tests/golden/ref_svgss_*.npz by running the UNMODIFIED reference CUDA rasteriser
(oracle/_ref/libsvgss_ref.so, built by oracle/Makefile from /root/reference/svgss_rasterization)
on a B200:

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out'      (then copy the .npz files here)

Inputs are the seeded synthetic cases of tests/util.py (regenerated from the seed, not stored);
stored are the reference's outputs and internal state: radii, sorted 64-bit keys, point_list, tile
ranges, n_contrib, final_T, all forward images and every gradient tensor."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest  # noqa: F401,E402
import util  # noqa: E402

BVH_CASE = dict(P=3000, n_rays=6000, seed=41, size=0.02)
CASES = {"train": dict(P=2500, W=112, H=80, S=4, VS=52, seed=31), "eval": dict(P=1500, W=72, H=96, S=7, VS=64, seed=32)}


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for tag, kw in CASES.items():
        case = util.make_case(**kw)
        g = util.pixel_grads(case)
        r, out, bw = util.run_ref(case, grads=g)
        R, W, H = out["num_rendered"], kw["W"], kw["H"]
        T = ((W + 15) // 16) * ((H + 15) // 16)
        save = {"num_rendered": np.int64(R)}
        for k in ("color", "normal", "depth", "opacity", "feature", "vfeature", "weights", "radii"):
            save["out_" + k] = out[k].cpu().numpy()
        save["keys"] = r.state("keys", (R,), torch.int64).cpu().numpy().astype(np.uint64)
        save["point_list"] = r.state("point_list", (R,), torch.int32).cpu().numpy().astype(np.uint32)
        save["ranges"] = r.state("ranges", (T, 2), torch.int32).cpu().numpy().astype(np.uint32)
        save["n_contrib"] = r.state("n_contrib", (H * W,), torch.int32).cpu().numpy().astype(np.uint32)
        save["final_T"] = r.state("final_T", (H * W,), torch.float32).cpu().numpy()
        for k, v in bw.items():
            save["grad_" + k] = v.cpu().numpy()
        np.savez_compressed(os.path.join(outdir, f"ref_svgss_{tag}.npz"), **save)
        print(tag, "R", R, "visible", int((out["radii"] > 0).sum()))
    bvh_golden(outdir)


def bvh_golden(outdir):
    """ref_bvh_small.npz: the reference LBVH kernels (oracle/_ref/libbvh_ref.so = unmodified
    submodules/bvh/src/construct.cu + trace.cu) on the seeded case of util.make_bvh_case: node table,
    boxes, Morton codes, and the opacity trace (hit counts + visibility)."""
    from oracle import ref_cuda
    from test_bvh_oracle_cpu import torch_leaf_boxes
    c = util.make_bvh_case(**BVH_CASE)
    t = {k: torch.from_numpy(v).cuda() for k, v in c.items()}
    nodes, aabbs = torch_leaf_boxes(t["means"], t["scales"], t["rotations"])
    # The reference's bottom-up box merge races (no fence between box store and flag CAS,
    # construct.cu:232-263): now and then an internal box comes out as an incomplete merge. A run hit by
    # the race is recognisable without any other implementation -- some parent box is not the merge of its
    # children's boxes -- and is simply repeated, so the golden holds the reference's race-free output.
    for attempt in range(20):
        ref = ref_cuda.RefBvh(nodes, aabbs, t["means"], t["scales"], t["rotations"])
        n, a = ref.nodes.cpu().numpy(), ref.aabbs.cpu().numpy()
        P = c["means"].shape[0]
        l, r = n[:P - 1, 1], n[:P - 1, 2]
        if (a[:P - 1, :3] == np.minimum(a[l, :3], a[r, :3])).all() and (a[:P - 1, 3:] == np.maximum(a[l, 3:], a[r, 3:])).all():
            break
        print("bvh: reference build hit its merge race, repeating (attempt %d)" % attempt)
    else:
        raise RuntimeError("no race-free reference build in 20 attempts")
    # gaussian_model.py:379-382 get_inverse_covariance = L L^T with L = R diag(1/s) (general_utils.py:151-160)
    q = torch.nn.functional.normalize(t["rotations"], dim=-1)
    r, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y), 2 * (x * y + r * z),
                     1 - 2 * (x * x + z * z), 2 * (y * z - r * x), 2 * (x * z - r * y), 2 * (y * z + r * x),
                     1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    L = R * (1.0 / t["scales"])[:, None, :]
    M = L @ L.transpose(1, 2)
    ci = torch.stack([M[:, 0, 0], M[:, 0, 1], M[:, 0, 2], M[:, 1, 1], M[:, 1, 2], M[:, 2, 2]], -1).contiguous()
    cnt, vis = ref.trace_opacity(t["rays_o"], t["rays_d"], t["means"], ci, t["opacity"], t["normals"])
    np.savez_compressed(os.path.join(outdir, "ref_bvh_small.npz"), nodes=ref.nodes.cpu().numpy(),
                        aabbs=ref.aabbs.cpu().numpy(), morton=ref.morton.cpu().numpy().astype(np.uint64),
                        cov_inv=ci.cpu().numpy(), contributes=cnt.cpu().numpy(), visibility=vis.cpu().numpy())
    print("bvh: occluded rays", float((vis == 0).float().mean()), "mean hits", float(cnt.float().mean()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else HERE)
