#!/usr/bin/env python
"""bench.py -- headline benchmark of the svgir_b200 hot path.

Metric (BASELINE.json): stage-2 `svgss + render_equation` forward+backward iterations per second at
800x800 on a TensoIR-shaped synthetic cloud (config C3-train of SURVEY.md 8(d): 300k spatially-varying
surfels, one 800x800 view per iteration, 64 light samples, S=4 / VS=52 G-buffer channels).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU; under
                                                           torchrun every rank renders its own view and
                                                           the per-surfel gradients are summed over NVLink
                                                           peer memory inside the step's CUDA graph)
  python bench.py --workload relight ...                   C3-eval relighting frame, ms/frame
  python bench.py --impl reference ...                     the reference path restated on the host CPU
                                                           (oracle/), bounded sample per step
  python bench.py --impl reference_cuda ...                (extra) reference CUDA rasteriser (oracle/_ref)
                                                           + the reference's torch shading graph on the GPU

One JSON line on stdout (rank 0). See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "svg-ir_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

P_SURFELS, WIDTH, HEIGHT, NS, S_FEAT, VS_FEAT, N_VIEWS = 300_000, 800, 800, 64, 4, 52, 8
WORKLOAD = ("C3-train: stage-2 svgss+render_equation fwd+bwd, %dk SV surfels, one %dx%d view/iter, Ns=%d, S=%d, VS=%d"
            % (P_SURFELS // 1000, WIDTH, HEIGHT, NS, S_FEAT, VS_FEAT))
METRIC, UNIT = "fwd+bwd iters/sec at 800x800", "it/s"
REF_TILE_STEP = 4        # --impl reference renders 1/16 of the tiles and shades 1/16 of the surfels per step


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def build_host_workload(seed=1234):
    from svgir_b200 import scene
    cloud = scene.make_surfels(P_SURFELS, seed=seed)
    mats = scene.make_materials(cloud, NS, seed=seed + 1)
    cams = [scene.look_at_camera(WIDTH, HEIGHT, v, N_VIEWS) for v in range(N_VIEWS)]
    rng = np.random.default_rng(seed + 2)
    gts = [rng.uniform(0, 1, (3, HEIGHT, WIDTH)).astype(np.float32) for _ in range(2)]
    return cloud, mats, cams, gts


def flat_grads(params):
    return torch.cat([p.grad.reshape(-1) for p in params])


def run_ours(args):
    import torch.distributed as dist
    from svgir_b200 import _lib, pipeline, shading
    from svgir_b200 import dist as svdist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the svgir_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()  # fail loudly if the extension is missing
    # default: shade only the surfels that survive the rasteriser's culling (identical images and gradients,
    # tests/test_culled_shading_gpu.py); --shade-all shades every surfel in the reference's order
    pipeline.SHADE_CULLED = bool(args.shade_all)
    # default: the resolve + loss tail is one fused kernel per direction (svgir_b200.losses); --torch-loss runs the torch
    # mirror of the reference's tail instead (same loss and gradients, tests/test_fused_loss_gpu.py)
    pipeline.FUSED_LOSS = not args.torch_loss

    cloud, mats, cams, gts = build_host_workload()
    pc = pipeline.model_from_scene(cloud, mats, dev)
    env = torch.from_numpy(mats["env_param"]).to(dev).requires_grad_(True)
    bg = torch.zeros(3, device=dev)
    cam_dev = [pipeline.camera_from_scene(c, dev) for c in cams]
    gt_dev = [torch.from_numpy(g).to(dev) for g in gts]
    params = pc.trainable() + [env]
    # N>1: .grad of every parameter is a view into one flat buffer, reduced with ONE NCCL all-reduce
    # per step (svgir_b200/dist.py); N=1 lets autograd hand over its gradient tensors.
    # The bucket is laid out in two segments in the order the backward pass finishes them (rasteriser-side
    # gradients, then shading-side ones); each segment's all-reduce is issued from inside the backward pass, so
    # the first one travels over NVLink while the shading backward kernel is still running, and both are
    # captured INSIDE the step's CUDA graph (--reduce post: one all-reduce after the graph instead).
    overlap = world > 1 and args.reduce == "overlap"
    # --reduce p2p: the bucket lives in peer-mapped (symmetric) memory and ONE svgir kernel per rank sums it over
    # NVLink at the end of the step's graph (csrc/peer_allreduce.cu); falls back to the NCCL all-reduce after the
    # graph when the box cannot provide peer-mapped memory
    peer, reduce_note = None, None
    if world > 1 and args.reduce in ("p2p", "p2p-overlap"):
        ok = torch.ones(1, device=dev)
        try:
            peer = svdist.PeerAllReduce(dev)
            if args.reduce == "p2p":
                bucket = svdist.FlatGradBucket(params, extra_floats=1, alloc=peer.allocate, reducer=peer.all_reduce)
            else:   # rasteriser-side segment summed on a side stream under the shading backward
                bucket = svdist.FlatGradBucket(params, segments=pipeline.reduce_segments(pc), extra_floats=1,
                                               alloc=peer.allocate, segment_peer=peer)
        except Exception as e:  # noqa: BLE001
            ok.zero_()
            reduce_note = "p2p unavailable (%s: %s); NCCL all-reduce after the step" % (type(e).__name__, str(e)[:120])
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0.0:
            peer, bucket = None, None
            reduce_note = reduce_note or "p2p unavailable on another rank; NCCL all-reduce after the step"
    in_graph = overlap or peer is not None
    if peer is None:
        bg_group = svdist.background_group(args.bg_ctas) if overlap and args.bg_ctas > 0 else None
        bucket = svdist.FlatGradBucket(params, segments=pipeline.reduce_segments(pc) if overlap else None,
                                       segment_groups=[bg_group, None] if overlap else None,
                                       extra_floats=1) if world > 1 else None
    # The step is captured once into a CUDA graph (pipeline.GraphedTrainingStep) and replayed: one
    # cudaGraphLaunch per iteration, camera + ground truth copied into static buffers, binning capacity
    # checked after every replay. --eager runs the same step launch by launch instead.
    runner = None if args.eager else pipeline.GraphedTrainingStep(pc, env, bg, cam_dev[0], gt_dev[0], bucket=bucket,
                                                                         reduce_in_graph=in_graph)

    def eager_step(i):
        v = (i * world + rank) % N_VIEWS
        if bucket is None:
            return pipeline.training_step(cam_dev[v], pc, env, bg, gt_dev[i % len(gt_dev)])
        bucket.zero()
        loss, res = pipeline.training_step(cam_dev[v], pc, env, bg, gt_dev[i % len(gt_dev)], zero_grad=False,
                                           overlap_bucket=bucket if in_graph else None)
        if not in_graph:
            bucket.all_reduce()  # per-surfel gradient exchange over NVLink
        return loss, res

    def step(i):
        if runner is None:
            return eager_step(i)
        v = (i * world + rank) % N_VIEWS
        loss, res = runner(cam_dev[v], gt_dev[i % len(gt_dev)])
        if bucket is not None and not in_graph:
            bucket.all_reduce()
        return loss, res

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) --------------------------------------------------
    for i in range(args.warmup):
        loss, res = step(i)
    sync_all()
    stats = {"R": int(res["num_rendered"]), "P_vis": int(res["visibility_filter"].sum())}
    _lib.launch_count(reset=True)
    _lib.timing_collect(reset=True)
    if runner is None:
        _lib.timing_enable(True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(args.steps):
        loss, res = step(args.warmup + i)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    launches = _lib.launch_count() if runner is None else runner.launches_per_step * args.steps
    if runner is not None:
        # per-kernel device times: the same kernels on the same inputs launched one by one, each bracketed
        # by CUDA events on the launching stream (events cannot be timed inside a replayed graph)
        _lib.timing_enable(True)
        for i in range(min(args.steps, 8)):
            eager_step(args.warmup + i)
        torch.cuda.synchronize()
    _lib.timing_enable(False)
    ktimes = {k: _lib.timing_collect(k) for k in ("composite_bwd", "composite_fwd", "shade_fwd", "shade_bwd", "preprocess",
                                                  "preprocess_bwd", "emit", "sort_small", "tile_scan", "train_loss_fwd", "train_loss_bwd",
                                                  "peer_allreduce")}
    _lib.timing_collect(reset=True)
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * args.steps / (ms_max / 1e3)

    # ---- end-to-end through the public API with host buffers (`e2e`) ----------------------------
    # A training iteration's INPUTS are the camera and the ground-truth image (the surfel parameters and the
    # per-surfel light buffers are optimiser state resident on the GPU, exactly as in the reference, whose
    # rasteriser API takes CUDA tensors): every step copies them from pinned host memory, runs the step
    # through the public call and reads the loss and num_rendered back. `e2e_cold` additionally re-uploads
    # ALL parameters and light buffers every step (the worst case: nothing resident).
    gt_host = [torch.from_numpy(g).pin_memory() for g in gts]
    cam_host = [pipeline.blocked_camera(HEIGHT, WIDTH, c.tanfovx, c.tanfovy, *[torch.from_numpy(getattr(c, k))
                for k in ("viewmatrix", "projmatrix", "campos", "patch_bbox", "prcppoint")], pin=True) for c in cams]
    h2d_bytes = gt_host[0].numel() * 4 + cam_host[0].block.numel() * 4

    def e2e_step(i):
        v = (i * world + rank) % N_VIEWS
        if runner is not None:
            loss, res = runner(cam_host[v], gt_host[i % len(gt_host)])   # H2D into the graph's static inputs
        else:
            c = cam_host[v]
            cam = pipeline.ViewCamera(HEIGHT, WIDTH, c.tanfovx, c.tanfovy, *[getattr(c, k).to(dev, non_blocking=True) for k in
                                      ("world_view_transform", "full_proj_transform", "camera_center", "patch_bbox", "prcppoint")])
            if bucket is not None:
                bucket.zero()
            loss, res = pipeline.training_step(cam, pc, env, bg, gt_host[i % len(gt_host)].to(dev, non_blocking=True),
                                               zero_grad=bucket is None, overlap_bucket=bucket if in_graph else None)
        if bucket is not None and not in_graph:
            bucket.all_reduce()
        return float(loss.item()), int(res["num_rendered"])  # D2H of the step's result

    host = {}
    for name, arr in (("xyz", cloud.means3D), ("opacity", cloud.opacity), ("scaling", cloud.scales),
                      ("rotation", cloud.rotations), ("shs", cloud.shs), ("base_color", mats["base_color"]),
                      ("roughness", mats["roughness"]), ("shading_normal", mats["shading_normals"]),
                      ("radiance", mats["radiance"]), ("visibility", mats["visibility"]),
                      ("incident_dirs", mats["incident_dirs"]), ("incident_areas", mats["incident_areas"]),
                      ("env", mats["env_param"]), ("gt", gts[0])):
        host[name] = torch.from_numpy(arr).pin_memory()
    cold_bytes = sum(t.numel() * 4 for t in host.values()) + h2d_bytes - gt_host[0].numel() * 4

    def cold_step(i):
        v = (i * world + rank) % N_VIEWS
        d = {k: t.to(dev, non_blocking=True) for k, t in host.items()}
        c = cam_host[v]
        cam = pipeline.ViewCamera(HEIGHT, WIDTH, c.tanfovx, c.tanfovy, *[getattr(c, k).to(dev, non_blocking=True) for k in
                                  ("world_view_transform", "full_proj_transform", "camera_center", "patch_bbox", "prcppoint")])
        m = pipeline.SurfelModel(d["xyz"], d["opacity"], d["scaling"], d["rotation"], d["shs"], d["base_color"],
                                 d["roughness"], d["shading_normal"], d["radiance"], d["visibility"],
                                 d["incident_dirs"], d["incident_areas"])
        for t in m.trainable():
            t.requires_grad_(True)
        envp = d["env"].requires_grad_(True)
        loss, res = pipeline.training_step(cam, m, envp, bg, d["gt"], zero_grad=False)
        if world > 1:
            dist.all_reduce(flat_grads(m.trainable() + [envp]))
        return float(loss.item()), int(res["num_rendered"])

    def timed(fn, n):
        for i in range(2):
            fn(i)
        sync_all()
        e0.record()
        for i in range(n):
            fn(2 + i)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * n / (float(t.item()) / 1e3)

    e2e_value = e2e_cold = None
    if not args.no_e2e:
        e2e_value = timed(e2e_step, args.steps)
        e2e_cold = timed(cold_step, min(args.steps, 5))

    if rank == 0:
        pk, pk_src = peaks()
        R, Pv = stats["R"], stats["P_vis"]
        b_rec = 104 + 4 * S_FEAT + 4 * VS_FEAT
        b_pix = 4 * (3 + 3 + 1 + 1 + S_FEAT + VS_FEAT // 4)
        n_sh = P_SURFELS if args.shade_all else Pv
        alg = {  # SURVEY.md 8(d) algorithmic bytes per launch, at this view's measured R / P_vis
            "composite_bwd": R * b_rec + WIDTH * HEIGHT * (b_pix + 12) + Pv * 4 * (15 + S_FEAT + VS_FEAT),
            "composite_fwd": R * b_rec + WIDTH * HEIGHT * (b_pix + 12),
            "shade_fwd": n_sh * (NS * 32 + 124) + n_sh * 4 * (12 * 5 + S_FEAT),
            "shade_bwd": n_sh * (NS * 32 + 124) + n_sh * 4 * (12 * 5 + S_FEAT) + n_sh * 4 * (12 + 4 + 12 + 3),
        }
        kt = {k: (v[0] / max(v[1], 1)) for k, v in ktimes.items()}  # avg ms per launch
        dom = max(("composite_bwd", "composite_fwd", "shade_fwd", "shade_bwd"), key=lambda k: kt[k])
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get(dom)
        except Exception:
            pass
        achieved = alg[dom] / (kt[dom] * 1e-3) / 1e9 if kt[dom] > 0 else 0.0
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "views_per_step": world, "parallelism": f"view-dp{world}",
                       "l2": "working set > L2 (per-sample light buffers 614 MB/iter)", "R": R, "P_vis": Pv,
                       "shading": "all %d surfels (reference order)" % P_SURFELS if args.shade_all else
                                  "the %d surfels that survive culling (preprocess runs first; images and gradients identical)" % Pv},
            "clocks": clk,
            "e2e": {"value": round(e2e_value, 3) if e2e_value else None, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": 12, "inputs": "camera matrices + ground-truth image from pinned host memory; loss + "
                    "num_rendered read back; parameters / light buffers resident (optimiser state)"},
            "e2e_cold": {"value": round(e2e_cold, 3) if e2e_cold else None, "unit": UNIT, "h2d_bytes_per_step": int(cold_bytes),
                         "note": "worst case: every parameter and light buffer re-uploaded each step (eager path)"},
            "loss_tail": "torch mirror of svgss.py:187-294 (~120 elementwise kernels)" if args.torch_loss else
                         "fused resolve+loss kernels (csrc/resolve.cu), one per direction",
            "launch_mode": "eager" if runner is None else "cuda-graph (1 capture, %d svgir kernels/step)" % runner.launches_per_step,
            "kernel_timing": "CUDA events around each launch on the launching stream" + ("" if runner is None else
                             ", separate eager pass of the same kernels/inputs right after the timed region"),
            "gpu_launches": int(launches),
            "grad_allreduce": None if bucket is None else {
                "bytes": bucket.nbytes, "mode":
                (("svgir_peer_allreduce over NVLink peer memory (%s), in the step's graph: " % (
                    "NVSwitch multicast ld_reduce/st" if peer.multicast else "128-bit peer loads/stores")) +
                 ("one kernel at the end of the step" if args.reduce == "p2p" else
                  "rasteriser-side segment on a side stream (%d CTAs) under the shading backward, shading-side segment after it"
                  % peer.BG_GRID)) if peer is not None else
                (("2 segments issued inside the backward pass, captured in the step's graph; the overlapped one on a "
                  "%d-CTA communicator" % args.bg_ctas if args.bg_ctas > 0 else
                  "2 segments issued inside the backward pass, captured in the step's graph") if overlap else
                 "one NCCL all-reduce after the step"), "note": reduce_note},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": pk["hbm_gbs"],
                         "peak_source": pk_src, "unit": "GB/s", "frac": round(achieved / pk["hbm_gbs"], 4),
                         "traffic": traffic, "algorithmic_bytes": int(alg[dom]), "avg_ms": round(kt[dom], 4)},
            "kernels_ms": {k: round(v, 4) for k, v in kt.items()},
            # the same per-launch event times grouped by stage: svgss rasteriser forward+backward alone (BASELINE.json
            # configs[1]'s shape), the render_equation shading, the resolve+loss tail
            "stage_ms": {
                "svgss_fwd_bwd": round(sum(kt[k] for k in ("preprocess", "tile_scan", "emit", "sort_small", "composite_fwd",
                                                           "composite_bwd", "preprocess_bwd")) + kt["tile_scan"], 4),
                "render_equation_fwd_bwd": round(kt["shade_fwd"] + kt["shade_bwd"], 4),
                "loss_tail": round(kt["train_loss_fwd"] + kt["train_loss_bwd"], 4)},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_sample(1, 0)
        print(json.dumps(line), flush=True)
    if world > 1:
        # drop the captured graph before the communicator goes away; with collectives captured INSIDE the graph
        # (--reduce overlap) NCCL's teardown was seen to block (gpurun_out/s2), so that mode leaves without it
        runner = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if in_graph:
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
def run_c4(args):
    """C4 (BASELINE.json configs[3]): multi-view data-parallel training -- 1M SV surfels, 8 views of 800x800 per step,
    sharded over the ranks by view (rank r renders views r, r+N, ...; strong scaling: the step's work is fixed). Every
    rank replays ONE captured step graph per local view, accumulating into its flat gradient bucket (zero_in_graph=
    False), then the buckets are summed once: by the svgir peer-memory kernel (default) or NCCL (--reduce post).
    value = view-iterations per second = 8 x steps / time. Not part of the driver's default run."""
    import torch.distributed as dist
    from svgir_b200 import _lib, pipeline, scene
    from svgir_b200 import dist as svdist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the svgir_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    pipeline.SHADE_CULLED = bool(args.shade_all)
    pipeline.FUSED_LOSS = not args.torch_loss
    P, V = 1_000_000, 8
    cloud = scene.make_surfels(P, seed=1238)
    mats = scene.make_materials(cloud, NS, seed=1239)
    pc = pipeline.model_from_scene(cloud, mats, dev)
    env = torch.from_numpy(mats["env_param"]).to(dev).requires_grad_(True)
    bg = torch.zeros(3, device=dev)
    cams = [pipeline.camera_from_scene(scene.look_at_camera(WIDTH, HEIGHT, v, V), dev) for v in range(V)]
    rng = np.random.default_rng(1240)
    gts = [torch.from_numpy(rng.uniform(0, 1, (3, HEIGHT, WIDTH)).astype(np.float32)).to(dev) for _ in range(2)]
    params = pc.trainable() + [env]
    mine = svdist.views_for_rank(V, rank, world)

    peer, note = None, None
    if world > 1 and args.reduce != "post":
        ok = torch.ones(1, device=dev)
        try:
            peer = svdist.PeerAllReduce(dev)
            bucket = svdist.FlatGradBucket(params, alloc=peer.allocate, reducer=peer.all_reduce)
        except Exception as e:  # noqa: BLE001
            ok.zero_()
            note = "p2p unavailable (%s: %s); NCCL all-reduce" % (type(e).__name__, str(e)[:120])
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0.0:
            peer = None
            note = note or "p2p unavailable on another rank; NCCL all-reduce"
    if peer is None:
        bucket = svdist.FlatGradBucket(params)
    runner = pipeline.GraphedTrainingStep(pc, env, bg, cams[0], gts[0], bucket=bucket, zero_in_graph=False)

    def step(i):
        bucket.zero()
        R = 0
        for v in mine:
            loss, res = runner(cams[v], gts[(i + v) % 2])
            R += int(res["num_rendered"])
        bucket.all_reduce()   # one exchange per step (no-op at N=1)
        return loss, R

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        loss, R = step(i)
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(args.steps):
        loss, R = step(args.warmup + i)
    e1.record()
    sync_all()
    clk = clocks.stop() if rank == 0 else None
    t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    # per-kernel times of one local view, launched eagerly
    _lib.timing_collect(reset=True)
    _lib.timing_enable(True)
    bucket.zero()
    pipeline.training_step(cams[mine[0]], pc, env, bg, gts[0], zero_grad=False)
    torch.cuda.synchronize()
    _lib.timing_enable(False)
    kt = {}
    for k in ("composite_bwd", "composite_fwd", "shade_fwd", "shade_bwd", "preprocess", "preprocess_bwd", "emit", "sort_small"):
        t, n = _lib.timing_collect(k)
        kt[k] = round(t / max(n, 1), 4)
    _lib.timing_collect(reset=True)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": round(V * args.steps / (ms / 1e3), 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C4: multi-view data-parallel stage-2 training, %dk SV surfels, %d views of %dx%d per step, Ns=%d, "
                                   "S=%d, VS=%d" % (P // 1000, V, WIDTH, HEIGHT, NS, S_FEAT, VS_FEAT),
                       "views_per_step": V, "views_per_rank": len(mine), "parallelism": f"view-dp{world}",
                       "l2": "working set > L2 (per-sample light buffers 2 GB/view)", "R_last_step_local": R},
            "clocks": clk, "gpu_launches": int(runner.launches_per_step * len(mine) * args.steps + (args.steps if peer is not None else 0)),
            "grad_allreduce": None if world == 1 else {"bytes": bucket.nbytes, "mode": "svgir_peer_allreduce after the local views"
                                                       if peer is not None else "one NCCL all-reduce after the local views", "note": note},
            "kernels_ms": kt}), flush=True)
    if world > 1:
        runner = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ---------------------------------------------------------------------------------------------
def run_relight(args):
    """C3-eval (BASELINE.json configs[2]): relighting frame = render_equation over all 300k surfels with
    Ns=384 samples under a fixed HDR env map (EnvLight semantics, scene/envmap.py:54-72) + svgss forward
    with the eval G-buffer (S=7, VS=64) at 800x800. Forward only; reports ms/frame. Under torchrun the
    view x envmap grid is sharded round-robin with no collective."""
    import torch.distributed as dist
    from svgir_b200 import _lib, pipeline, scene, shading
    from svgir_b200 import dist as svdist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the svgir_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    pipeline.SHADE_CULLED = bool(args.shade_all)
    # default: the resolve + loss tail is one fused kernel per direction (svgir_b200.losses); --torch-loss runs the torch
    # mirror of the reference's tail instead (same loss and gradients, tests/test_fused_loss_gpu.py)
    pipeline.FUSED_LOSS = not args.torch_loss
    ns, n_env = 384, 5
    cloud = scene.make_surfels(P_SURFELS, seed=1236)
    mats = scene.make_materials(cloud, ns, seed=1237)
    pc = pipeline.model_from_scene(cloud, mats, dev, requires_grad=False)
    cams = [pipeline.camera_from_scene(scene.look_at_camera(WIDTH, HEIGHT, v, N_VIEWS), dev) for v in range(N_VIEWS)]
    rng = np.random.default_rng(7)
    envs = [torch.from_numpy(rng.uniform(0, 4, (32, 64, 3)).astype(np.float32)).to(dev) for _ in range(n_env)]
    bg = torch.zeros(3, device=dev)
    grid = svdist.relight_grid_for_rank(N_VIEWS, n_env, rank, world)

    # The frame is captured once into a CUDA graph (pipeline.GraphedRelightFrame) and replayed per (view, env map):
    # camera block + env map copied into static buffers, binning capacity checked after every replay.
    # --eager issues the ~70 launches of a frame one by one instead (and blocks on num_rendered mid-frame).
    runner = None if args.eager else pipeline.GraphedRelightFrame(pc, envs[0], bg, cams[0])

    def eager_frame(i):
        e, v = grid[i % len(grid)]  # (env, view), env-major
        with torch.no_grad():
            return pipeline.render_view(cams[v], pc, (envs[e], shading.MODE_FIXED), bg, is_training=False)

    def frame(i):
        if runner is None:
            return eager_frame(i)
        e, v = grid[i % len(grid)]
        return runner(cams[v], envs[e])

    for i in range(args.warmup):
        res = frame(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    _lib.launch_count(reset=True)
    _lib.timing_collect(reset=True)
    if runner is None:
        _lib.timing_enable(True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        res = frame(args.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clk = clocks.stop() if rank == 0 else None
    launches = _lib.launch_count() if runner is None else runner.launches_per_frame * args.steps
    if runner is not None:   # per-kernel times: the same frames launched eagerly, each launch bracketed by events
        _lib.timing_enable(True)
        for i in range(min(args.steps, 5)):
            eager_frame(args.warmup + i)
        torch.cuda.synchronize()
    _lib.timing_enable(False)
    kt = {}
    for k in ("shade_fwd", "composite_fwd", "preprocess", "emit", "sort_small", "tile_scan"):
        t, n = _lib.timing_collect(k)
        kt[k] = t / max(n, 1)
    _lib.timing_collect(reset=True)
    t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_frame = float(t_ms.item()) / args.steps

    # end to end: camera block + env map from pinned host memory every frame, the relit image read back to pinned
    # host memory (what eval_relighting_tensoIR.py:331-340 saves)
    e2e_ms = None
    if not args.no_e2e:
        cam_np = [scene.look_at_camera(WIDTH, HEIGHT, v, N_VIEWS) for v in range(N_VIEWS)]
        cam_host = [pipeline.blocked_camera(HEIGHT, WIDTH, c.tanfovx, c.tanfovy, *[torch.from_numpy(getattr(c, k))
                    for k in ("viewmatrix", "projmatrix", "campos", "patch_bbox", "prcppoint")], pin=True) for c in cam_np]
        env_host = [e.cpu().pin_memory() for e in envs]
        img_host = torch.empty((3, HEIGHT, WIDTH), dtype=torch.float32).pin_memory()

        def e2e_frame(i):
            e, v = grid[i % len(grid)]
            if runner is not None:
                r = runner(cam_host[v], env_host[e])
            else:
                c = cam_host[v]
                cam = pipeline.blocked_camera(HEIGHT, WIDTH, c.tanfovx, c.tanfovy, c.world_view_transform, c.full_proj_transform,
                                              c.camera_center, c.patch_bbox, c.prcppoint, device=dev)
                with torch.no_grad():
                    r = pipeline.render_view(cam, pc, (env_host[e].to(dev, non_blocking=True), shading.MODE_FIXED), bg,
                                             is_training=False)
            img_host.copy_(r["pbr"], non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for i in range(2):
            e2e_frame(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        for i in range(args.steps):
            e2e_frame(2 + i)
        e1.record()
        torch.cuda.synchronize()
        t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2.item()) / args.steps
        e2e_h2d = cam_host[0].block.numel() * 4 + env_host[0].numel() * 4
        e2e_d2h = img_host.numel() * 4
    if rank == 0:
        pk, pk_src = peaks()
        # shading runs on the surfels that survive the rasteriser's culling unless --shade-all
        n_sh = P_SURFELS if args.shade_all else int(res["visibility_filter"].sum())
        alg = n_sh * (ns * 32 + 124) + n_sh * 4 * (12 * 5 + 7)
        ach = alg / (kt["shade_fwd"] * 1e-3) / 1e9 if kt["shade_fwd"] > 0 else 0.0
        print(json.dumps({
            "metric": "relight ms/frame", "value": round(ms_frame / world, 4), "unit": "ms/frame", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_frame, 4), "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3-eval: relight frame = render_equation (Ns=%d) over %dk surfels + svgss forward "
                                   "S=7/VS=64 at %dx%d, fixed HDR env map" % (ns, P_SURFELS // 1000, WIDTH, HEIGHT),
                       "grid": "%d views x %d env maps, round-robin over ranks" % (N_VIEWS, n_env),
                       "l2": "working set > L2 (light buffers 3.7 GB/frame)", "R": int(res["num_rendered"]),
                       "surfels_shaded": n_sh},
            "clocks": clk, "gpu_launches": int(launches),
            "launch_mode": "eager" if runner is None else "cuda-graph (1 capture, %d svgir kernels/frame)" % runner.launches_per_frame,
            "e2e": None if e2e_ms is None else {"value": round(e2e_ms / world, 4), "unit": "ms/frame", "h2d_bytes_per_step": int(e2e_h2d),
                                                "d2h_bytes_per_step": int(e2e_d2h),
                                                "inputs": "camera block + env map from pinned host memory; relit image read back"},
            "roofline": {"bound": "hbm", "kernel": "shade_fwd", "achieved": round(ach, 1), "peak": pk["hbm_gbs"],
                         "peak_source": pk_src, "unit": "GB/s", "frac": round(ach / pk["hbm_gbs"], 4), "traffic": None,
                         "algorithmic_bytes": int(alg), "avg_ms": round(kt["shade_fwd"], 4)},
            "kernels_ms": {k: round(v, 4) for k, v in kt.items()}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
_REF_CACHE = {}


def _ref_inputs():
    if "w" not in _REF_CACHE:
        _REF_CACHE["w"] = build_host_workload()
    return _REF_CACHE["w"]


def reference_cpu_step(view: int) -> float:
    """One BOUNDED SAMPLE of the workload on the host CPU with the oracle: full preprocess + binning,
    compositing fwd+bwd of every 4th tile row/column (1/16 of the tiles), shading fwd+bwd of every
    16th surfel. Returns seconds."""
    from oracle import svgss as O, shading_oracle as SO
    cloud, mats, cams, gts = _ref_inputs()
    cam = cams[view % N_VIEWS]
    t0 = time.perf_counter()
    sl = slice(view % 16, None, 16)
    tt = {k: torch.from_numpy(np.ascontiguousarray(mats[k][sl])) for k in
          ("base_color", "roughness", "shading_normals", "radiance", "visibility", "incident_dirs", "incident_areas")}
    envp = torch.from_numpy(mats["env_param"]).requires_grad_(True)
    for k in ("base_color", "roughness", "shading_normals"):
        tt[k].requires_grad_(True)
    vd = torch.from_numpy(cam.campos[None] - cloud.means3D[sl])
    vd = torch.nn.functional.normalize(vd, dim=-1)
    pbr, extra = SO.rendering_equation4(tt["base_color"], tt["roughness"], tt["shading_normals"], vd, tt["radiance"],
                                        lambda d: SO.direct_light_learnable(envp, d), tt["visibility"],
                                        tt["incident_dirs"], tt["incident_areas"])
    feats, vfeats = SO.pack_features(pbr, extra, tt["base_color"], tt["roughness"], tt["shading_normals"],
                                     torch.from_numpy(cam.viewmatrix[:3, :3].copy()), True)
    (vfeats.sum() + feats.sum()).backward()
    # rasteriser: features of the un-shaded surfels do not change its cost
    rng = np.random.default_rng(view)
    f = rng.uniform(0, 1, (P_SURFELS, S_FEAT)).astype(np.float32)
    vf = rng.uniform(0, 1, (P_SURFELS, VS_FEAT)).astype(np.float32)
    O.lib().oracle_set_tile_step(REF_TILE_STEP)
    try:
        fw = O.forward(cam, cloud.means3D, cloud.opacity, cloud.scales, cloud.rotations, f, vf, shs=cloud.shs)
        g = [np.full(fw[k].shape, 1.0 / (HEIGHT * WIDTH), np.float32) for k in
             ("color", "normal_img", "depth", "opacity", "feature", "vfeature")]
        O.backward(fw, *g)
    finally:
        O.lib().oracle_set_tile_step(1)
    return time.perf_counter() - t0


def cpu_reference_sample(steps: int, warmup: int) -> dict:
    torch.set_num_threads(os.cpu_count() or 1)
    for i in range(warmup):
        reference_cpu_step(i)
    ts = [reference_cpu_step(warmup + i) for i in range(steps)]
    frac = 1.0 / (REF_TILE_STEP * REF_TILE_STEP)
    t = float(np.mean(ts))
    return {"value": round(frac / t, 6), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": ("oracle (C + torch-CPU restatement of the reference path) on 1/%d of the step: all %dk surfels "
                       "preprocessed+binned, every %dth tile row/col composited fwd+bwd, every 16th surfel shaded fwd+bwd; "
                       "%.2f s per sample, value = sample fraction / time" % (REF_TILE_STEP ** 2, P_SURFELS // 1000,
                                                                               REF_TILE_STEP, t)),
            "sample_seconds": round(t, 3)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_sample(args.steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(cb["sample_seconds"] * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD}, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_reference_cuda(args):
    """Extra arm: the reference's own CUDA rasteriser (oracle/_ref, unmodified sources) plus the
    reference's torch shading graph, both on this GPU (BASELINE.md 'R-step')."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import ref_cuda, shading_oracle as SO
    if not (torch.cuda.is_available() and ref_cuda.available()):
        print(json.dumps({"impl": "reference_cuda", "unavailable": "needs a GPU and oracle/_ref/libsvgss_ref.so"}))
        return
    dev = torch.device("cuda:0")
    cloud, mats, cams, gts = build_host_workload()
    d = lambda a: torch.from_numpy(a).to(dev)
    t = {k: d(v) for k, v in mats.items()}
    geo = dict(means3D=d(cloud.means3D), opacity=d(cloud.opacity), scales=d(cloud.scales), rotations=d(cloud.rotations),
               shs=d(cloud.shs))
    for k in ("base_color", "roughness", "shading_normals", "env_param"):
        t[k].requires_grad_(True)
    bg = torch.zeros(3, device=dev)
    cfg3 = torch.ones(3, device=dev)
    r = ref_cuda.RefSvgss()
    n = HEIGHT * WIDTH
    gpix = [torch.full(s, 1.0 / n, device=dev) for s in ((3, HEIGHT, WIDTH), (3, HEIGHT, WIDTH), (1, HEIGHT, WIDTH),
                                                         (1, HEIGHT, WIDTH), (S_FEAT, HEIGHT, WIDTH), (VS_FEAT // 4, HEIGHT, WIDTH))]

    def step(i):
        cam = cams[i % N_VIEWS]
        for k in ("base_color", "roughness", "shading_normals", "env_param"):
            t[k].grad = None
        vd = torch.nn.functional.normalize(d(cam.campos)[None] - geo["means3D"], dim=-1)
        pbr, extra = SO.rendering_equation4(t["base_color"], t["roughness"], t["shading_normals"], vd, t["radiance"],
                                            lambda x: SO.direct_light_learnable(t["env_param"], x), t["visibility"],
                                            t["incident_dirs"], t["incident_areas"])
        feats, vfeats = SO.pack_features(pbr, extra, t["base_color"], t["roughness"], t["shading_normals"],
                                         d(cam.viewmatrix[:3, :3].copy()), True)
        r.forward(bg=bg, means3D=geo["means3D"], features=feats.detach().contiguous(), vfeatures=vfeats.detach().contiguous(),
                  colors=None, opacity=geo["opacity"], scales=geo["scales"], rotations=geo["rotations"], scale_modifier=1.0,
                  viewmatrix=d(cam.viewmatrix), projmatrix=d(cam.projmatrix), prcppoint=d(cam.prcppoint),
                  patchbbox=d(cam.patch_bbox), tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, H=HEIGHT, W=WIDTH, sh=geo["shs"],
                  degree=3, campos=d(cam.campos), config=cfg3)
        g = r.backward(*gpix)
        torch.autograd.backward([vfeats], [g["dL_dvfeatures"]])

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"impl": "reference_cuda", "metric": METRIC, "value": round(args.steps / (ms / 1e3), 3), "unit": UNIT,
                      "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
                      "config": {"workload": WORKLOAD}, "note": "reference CUDA rasteriser (sm_100 build of the unmodified "
                      "sources) + reference torch shading graph on the same B200"}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference_cuda"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--shade-all", action="store_true", help="shade culled surfels too (reference order: shading before the rasteriser)")
    ap.add_argument("--torch-loss", action="store_true", help="resolve + loss tail in torch (the reference's ~120 kernels) instead of the fused kernels")
    ap.add_argument("--reduce", default="p2p-overlap", choices=["overlap", "post", "p2p", "p2p-overlap"],
                    help="N>1 gradient exchange. p2p-overlap (default): svgir kernels over NVLink peer memory inside the step's graph, "
                         "the rasteriser-side segment on a side stream under the shading backward, the shading-side segment after it "
                         "(B200 x8: 3021 it/s = 96.5%% of 8 x N=1); p2p: one such kernel at the end of the step (2891 it/s); both fall "
                         "back to `post` if the box has no peer-mapped memory. post: one NCCL all-reduce after the graph (2739 it/s). "
                         "overlap: NCCL, segment-wise inside the backward pass (slower: the NCCL kernel takes SMs from the shading backward)")
    ap.add_argument("--bg-ctas", type=int, default=4, help="--reduce overlap: CTA limit of the communicator that carries the "
                    "segment overlapped with the shading backward (0 = default communicator for both segments)")
    ap.add_argument("--eager", action="store_true", help="launch the step kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--workload", default="train", choices=["train", "relight", "c4"],
                    help="train = C3-train fwd+bwd it/s (headline); relight = C3-eval forward ms/frame (Ns=384, S=7, VS=64); "
                         "c4 = 1M surfels, 8 views per step sharded over the ranks (strong scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference_cuda":
        run_reference_cuda(args)
    elif args.workload == "relight":
        run_relight(args)
    elif args.workload == "c4":
        run_c4(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
