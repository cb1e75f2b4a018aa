#!/usr/bin/env python
"""Host-side cProfile of the eager C3-train step (diagnostic): where does the Python/driver time go?"""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "svg-ir_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch

import bench


def main():
    from svgir_b200 import pipeline
    pipeline.SHADE_CULLED = "--shade-all" in sys.argv
    dev = torch.device("cuda:0")
    cloud, mats, cams, gts = bench.build_host_workload()
    pc = pipeline.model_from_scene(cloud, mats, dev)
    env = torch.from_numpy(mats["env_param"]).to(dev).requires_grad_(True)
    bg = torch.zeros(3, device=dev)
    cam_dev = [pipeline.camera_from_scene(c, dev) for c in cams]
    gt = torch.from_numpy(gts[0]).to(dev)
    step = lambda i: pipeline.training_step(cam_dev[i % len(cam_dev)], pc, env, bg, gt)
    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for i in range(10):
        step(4 + i)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(35)
    st.sort_stats("tottime").print_stats(20)


if __name__ == "__main__":
    main()
