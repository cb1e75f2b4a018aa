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


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else HERE)
