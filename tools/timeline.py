#!/usr/bin/env python
"""GPU timeline of the C3-train step (no nsys in the image): runs a few steps under torch.profiler
(CUPTI) and prints every kernel / memcpy of ONE step with its start offset, duration and the idle
gap before it, plus busy / idle totals. Diagnostic only -- never a bench number.

  python tools/timeline.py [--graph] [--steps 4] > gpurun_out/timeline.txt
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "svg-ir_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch
from torch.profiler import ProfilerActivity, profile

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--shade-all", action="store_true")
    args = ap.parse_args()
    from svgir_b200 import pipeline
    pipeline.SHADE_CULLED = bool(args.shade_all)
    dev = torch.device("cuda:0")
    cloud, mats, cams, gts = bench.build_host_workload()
    pc = pipeline.model_from_scene(cloud, mats, dev)
    env = torch.from_numpy(mats["env_param"]).to(dev).requires_grad_(True)
    bg = torch.zeros(3, device=dev)
    cam_dev = [pipeline.camera_from_scene(c, dev) for c in cams]
    gt = torch.from_numpy(gts[0]).to(dev)
    if args.graph:
        runner = pipeline.GraphedTrainingStep(pc, env, bg, cam_dev[0], gt)
        step = lambda i: runner(cam_dev[i % len(cam_dev)], gt)
    else:
        step = lambda i: pipeline.training_step(cam_dev[i % len(cam_dev)], pc, env, bg, gt)
    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    marks = []
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            torch.cuda.synchronize()
            step(4 + i)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # split into steps at gaps > 300 us following a synchronize (steps are separated by syncs)
    groups, cur, last_end = [], [], None
    for e in evs:
        s, t = e.time_range.start, e.time_range.end
        if last_end is not None and s - last_end > 400 and cur and len(cur) > 20:
            groups.append(cur)
            cur = []
        cur.append(e)
        last_end = max(last_end or t, t)
    if cur:
        groups.append(cur)
    g = groups[-1] if groups else []
    if not g:
        print("no CUDA events")
        return
    t0 = g[0].time_range.start
    busy, prev_end = 0.0, t0
    print("%9s %9s %8s  %s" % ("start_us", "dur_us", "gap_us", "name"))
    for e in g:
        s, t = e.time_range.start, e.time_range.end
        gap = s - prev_end
        print("%9.1f %9.1f %8.1f  %s" % (s - t0, t - s, gap, e.name[:110]))
        busy += t - s
        prev_end = max(prev_end, t)
    span = prev_end - t0
    print("# step span %.1f us, kernel busy %.1f us (%.1f%%), idle %.1f us, %d launches" %
          (span, busy, 100 * busy / span, span - busy, len(g)))
    # aggregate by name
    agg = {}
    for e in g:
        a = agg.setdefault(e.name[:80], [0, 0.0])
        a[0] += 1
        a[1] += e.time_range.end - e.time_range.start
    print("# by kernel:")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("#  %8.1f us  x%-3d %s" % (t, n, k))


if __name__ == "__main__":
    main()
