#!/usr/bin/env python
"""bench.py -- headline benchmark of the svgir_b200 hot path.

Metric (BASELINE.json): stage-2 `svgss + render_equation` forward+backward iterations per second at
800x800 on a TensoIR-shaped synthetic cloud (config C3-train of SURVEY.md 8(d): 300k spatially-varying
surfels, one 800x800 view per iteration, 64 light samples, S=4 / VS=52 G-buffer channels).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU; under
                                                           torchrun every rank renders its own view and
                                                           the per-surfel gradients are summed over NVLink
                                                           peer memory inside the step's CUDA graph).
                                                           The default line also carries, measured after the
                                                           timed region: `c4` (BASELINE configs[3]: 1M surfels,
                                                           8 views/step, strong scaling; every N), and at N=1
                                                           `relight` (C3-eval ms/frame), `visibility` (LBVH
                                                           build + trace vs the reference kernels),
                                                           `reference_cuda` (the reference CUDA extension +
                                                           torch shading on this GPU: the denominator of the
                                                           >=8x target) and `cpu_baseline`; at N>1
                                                           `grad_allreduce.max_rel_err` (in-graph exchange vs
                                                           NCCL on the same gradients). --no-extras skips them.
  python bench.py --workload relight|c4|visibility ...     one of those workloads on its own
  python bench.py --impl reference ...                     the reference path restated on the host CPU
                                                           (oracle/), FULL steps, all host threads
  python bench.py --impl reference_cuda ...                (extra) reference CUDA rasteriser (oracle/_ref)
                                                           + the torch restatement of the reference's shading
                                                           graph (oracle/shading_oracle.py) on the GPU

One JSON line on stdout (rank 0). See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import gc
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
C5_SHAPE = ("C5", 2_000_000, 1920, 1080, 200)   # BASELINE.json configs[4]
WORKLOAD = ("C3-train: stage-2 svgss+render_equation fwd+bwd, %dk SV surfels, one %dx%d view/iter, Ns=%d, S=%d, VS=%d"
            % (P_SURFELS // 1000, WIDTH, HEIGHT, NS, S_FEAT, VS_FEAT))
METRIC, UNIT = "fwd+bwd iters/sec at 800x800", "it/s"
L2_NOTE = "working set > L2 (per-sample light buffers 614 MB/iter)"
EXTRAS_DEADLINE_S = 420   # watchdog: if an extra workload hangs, the headline line is printed without it


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def train_config(world: int) -> dict:
    """`config` of the headline line -- identical in the repo arm and the reference arm."""
    return {"workload": WORKLOAD, "views_per_step": world, "parallelism": f"view-dp{world}", "l2": L2_NOTE}


class ClockSampler:
    """Samples SM clock / power / throttle reasons of one GPU from a thread of this process (NVML, ~3 ms period;
    `nvidia-smi -lms` as the fallback) from start() on. mark() brackets the timed region: stop() reports the samples
    inside it and, for context, over the whole loaded window (warm-up + timed region + what follows)."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc, self.thread = gpu_index, [], None, None
        self.stop_flag = False
        self.t0 = self.t1 = None
        self.source = None

    def _nvml_loop(self, R, h):
        def const(new, old, default):
            return getattr(R, new, getattr(R, old, default))
        bits = [const("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                const("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                const("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                const("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
        get_reasons = getattr(R, "nvmlDeviceGetCurrentClocksEventReasons", None) or R.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = float(R.nvmlDeviceGetMaxClockInfo(h, R.NVML_CLOCK_SM))
        while not self.stop_flag:
            try:
                sm = float(R.nvmlDeviceGetClockInfo(h, R.NVML_CLOCK_SM))
                pw = R.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = int(get_reasons(h))
                self.rows.append((time.perf_counter(), sm, mx, pw, [bool(rs & b) for b in bits]))
            except Exception:
                pass
            time.sleep(0.003)

    def _smi_loop(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]), float(r[3]),
                                  [v.lower().startswith("active") for v in r[4:8]]))
            except Exception:
                pass

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber the devices: resolve through the PCI bus id torch reports
            try:
                pr = torch.cuda.get_device_properties(self.idx)
                h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)   # raises here, not in the thread, if NVML cannot answer
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(pynvml, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.source = "nvidia-smi -lms 20"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self, begin: bool):
        if begin:
            self.t0 = time.perf_counter()
        else:
            self.t1 = time.perf_counter()

    def stop(self) -> dict:
        self.stop_flag = True
        if self.proc is not None:
            time.sleep(0.1)
            self.proc.terminate()
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        self.thread.join(timeout=1.0)
        rows = list(self.rows)
        inside = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        use = inside if inside else rows

        def summarise(rr):
            sm = [r[1] for r in rr]
            reasons = sorted({n for r in rr for n, v in zip(self.NAMES, r[4]) if v})
            return (float(np.median(sm)) if sm else None, reasons, max((r[3] for r in rr), default=None))

        sm_med, reasons, pw = summarise(use)
        sm_all, reasons_all, _ = summarise(rows)
        return {"sm_mhz": sm_med, "sm_max_mhz": max((r[2] for r in rows), default=None), "reasons": reasons,
                "samples": len(inside), "power_w_max": pw, "source": self.source,
                "window": "timed region" if inside else "whole run (no sample fell inside the timed region)",
                "whole_run": {"samples": len(rows), "sm_mhz": sm_all, "reasons": reasons_all}}


class Ctx:
    """Per-process launch context (torchrun contract: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py: no CUDA device; the svgir_b200 path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        from svgir_b200 import _lib
        _lib.lib()  # fail loudly if the extension is missing

    def sync_all(self):
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def max_ms(self, ms: float) -> float:
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def make_peer_bucket(ctx: Ctx, make_bucket):
    """(peer, bucket, note): the flat gradient bucket in peer-mapped memory, or (None, None, why). The ranks first AGREE
    (MIN all-reduce) that each of them can construct a PeerAllReduce; only then do they enter allocate(), which runs
    collectives itself (symmetric-memory rendezvous + barrier) -- a rank failing inside a collective cannot be caught
    by the others, so nothing that may fail rank-locally is left between the agreement and the rendezvous."""
    import torch.distributed as dist
    from svgir_b200 import dist as svdist
    ok = torch.ones(1, device=ctx.dev)
    peer, note = None, None
    try:
        import importlib
        importlib.import_module("torch.distributed._symmetric_memory")   # import failures are rank-local
        peer = svdist.PeerAllReduce(ctx.dev)
    except Exception as e:  # noqa: BLE001
        ok.zero_()
        note = "p2p unavailable (%s: %s); NCCL all-reduce" % (type(e).__name__, str(e)[:120])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok.item()) == 0.0:
        return None, None, note or "p2p unavailable on another rank; NCCL all-reduce"
    return peer, make_bucket(peer), None


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


def measure_train(args, ctx: Ctx, clocks: ClockSampler) -> dict:
    """The headline: C3-train, weak scaling (one view per rank per step). Returns the JSON line (a dict; on every rank)."""
    import torch.distributed as dist
    from svgir_b200 import _lib, pipeline
    from svgir_b200 import dist as svdist
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
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
    # N>1: .grad of every parameter is a view into one flat buffer (svgir_b200/dist.py), laid out in two
    # segments in the order the backward pass finishes them (rasteriser-side gradients, then shading-side ones).
    #   p2p-overlap (default): the buffer lives in peer-mapped memory and svgir kernels sum it over NVLink INSIDE the
    #       step's CUDA graph, the first segment on a side stream under the shading backward (csrc/peer_allreduce.cu)
    #   p2p: one such kernel at the end of the step's graph;  post: one NCCL all-reduce after the graph;
    #   overlap: NCCL, segment-wise inside the backward pass
    overlap = world > 1 and args.reduce == "overlap"
    peer, bucket, reduce_note = None, None, None
    if world > 1 and args.reduce in ("p2p", "p2p-overlap"):
        if args.reduce == "p2p":
            mk = lambda pr: svdist.FlatGradBucket(params, extra_floats=1, alloc=pr.allocate, reducer=pr.all_reduce)
        else:
            mk = lambda pr: svdist.FlatGradBucket(params, segments=pipeline.reduce_segments(pc), extra_floats=1,
                                                  alloc=pr.allocate, segment_peer=pr)
        peer, bucket, reduce_note = make_peer_bucket(ctx, mk)
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

    def eager_step(i, reduce=True):
        v = (i * world + rank) % N_VIEWS
        if bucket is None:
            return pipeline.training_step(cam_dev[v], pc, env, bg, gt_dev[i % len(gt_dev)])
        bucket.zero()
        loss, res = pipeline.training_step(cam_dev[v], pc, env, bg, gt_dev[i % len(gt_dev)], zero_grad=False,
                                           overlap_bucket=bucket if (in_graph and reduce) else None)
        if not in_graph and reduce:
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

    # ---- device-resident throughput (`value`) --------------------------------------------------
    for i in range(args.warmup):
        loss, res = step(i)
    ctx.sync_all()
    stats = {"R": int(res["num_rendered"]), "P_vis": int(res["visibility_filter"].sum())}
    _lib.launch_count(reset=True)
    _lib.timing_collect(reset=True)
    if runner is None:
        _lib.timing_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync_all()
    clocks.mark(True)
    e0.record()
    for i in range(args.steps):
        loss, res = step(args.warmup + i)
    e1.record()
    ctx.sync_all()
    clocks.mark(False)
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() if runner is None else runner.launches_per_step * args.steps
    if runner is not None:
        # per-kernel device times: the same kernels on the same inputs launched one by one, each bracketed
        # by CUDA events on the launching stream (events cannot be timed inside a replayed graph)
        _lib.timing_enable(True)
        for i in range(min(args.steps, 8)):
            if runner.fused:   # the step's own kernel sequence, launched outside the graph
                runner.load_inputs(cam_dev[((args.warmup + i) * world + rank) % N_VIEWS], gt_dev[i % len(gt_dev)])
                runner.fs.enqueue()
            else:
                eager_step(args.warmup + i)
        torch.cuda.synchronize()
    _lib.timing_enable(False)
    ktimes = {k: _lib.timing_collect(k) for k in ("composite_bwd", "composite_fwd", "shade_fwd", "shade_bwd", "preprocess",
                                                  "preprocess_bwd", "emit", "sort_small", "tile_scan", "train_loss_fwd", "train_loss_bwd",
                                                  "peer_allreduce")}
    _lib.timing_collect(reset=True)
    ms_max = ctx.max_ms(ms)
    value = world * args.steps / (ms_max / 1e3)

    # ---- correctness of the in-graph gradient exchange (N>1) -----------------------------------------
    # One step's LOCAL gradients (no exchange) are cloned and summed by NCCL; the same local gradients are then summed
    # in place by the path the step graph uses (svgir peer-memory kernels segment by segment, or NCCL). Both sums see
    # identical inputs, so the difference is association order only.
    ar_check = None
    if world > 1 and bucket is not None:
        try:
            eager_step(args.warmup, reduce=False)
            ctx.sync_all()
            want = bucket.flat.clone()
            dist.all_reduce(want)
            if bucket.segment_peer is not None:
                bucket.segment_peer.begin_segments()
                nseg = len(bucket.seg_bounds)
                for k, (lo, hi) in enumerate(bucket.seg_bounds):
                    bucket.segment_peer.reduce_segment(k, lo, hi, last=(k == nseg - 1))
                bucket.segment_peer.end_segments()
            else:
                bucket.all_reduce()
            ctx.sync_all()
            got = bucket.flat
            scale = float(want.abs().max())
            err = float((got - want).abs().max())
            big = want.abs() > 1e-3 * scale
            elem = float(((got - want).abs()[big] / want.abs()[big]).max()) if bool(big.any()) else 0.0
            t = torch.tensor([err / max(scale, 1e-30), elem], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ar_check = {"max_rel_err": float(t[0]), "max_elem_rel_err": float(t[1]),
                        "check": "one step's local gradients summed by the step's own exchange path vs an NCCL all-reduce of a "
                                 "clone; max |diff| / max |sum| (and, for elements above 1e-3 of the max, the largest element-wise "
                                 "relative difference), max over ranks"}
            del want
        except Exception as e:  # noqa: BLE001
            ar_check = {"max_rel_err": None, "check_error": "%s: %s" % (type(e).__name__, str(e)[:200])}

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

    def e2e_pipelined(n):
        """The same per-step copies and read-backs, but step k+1's inputs are uploaded on a copy stream into the
        graph's second input slot while step k computes (GraphedTrainingStep.prefetch): what a data loader does."""
        runner.prefetch(cam_host[rank % N_VIEWS], gt_host[0])
        for i in range(n):
            runner.replay_prefetched()
            if i + 1 < n:
                runner.prefetch(cam_host[((i + 1) * world + rank) % N_VIEWS], gt_host[(i + 1) % len(gt_host)])
            if bucket is not None and not in_graph:
                bucket.all_reduce()
            R = runner.finish()
            float(runner.loss.item()), int(R)

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
        ctx.sync_all()
        e0.record()
        for i in range(n):
            fn(2 + i)
        e1.record()
        ctx.sync_all()
        return world * n / (ctx.max_ms(e0.elapsed_time(e1)) / 1e3)

    e2e_value = e2e_cold = e2e_pipe = None
    if not args.no_e2e:
        e2e_value = timed(e2e_step, args.steps)
        if runner is not None and hasattr(runner, "prefetch"):
            e2e_pipelined(2)
            ctx.sync_all()
            e0.record()
            e2e_pipelined(args.steps)
            e1.record()
            ctx.sync_all()
            e2e_pipe = world * args.steps / (ctx.max_ms(e0.elapsed_time(e1)) / 1e3)
        e2e_cold = timed(cold_step, min(args.steps, 5))

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
    traffic, limiter = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        traffic, limiter = tj.get(dom), tj.get(dom + "_limiter")
    except Exception:
        pass
    achieved = alg[dom] / (kt[dom] * 1e-3) / 1e9 if kt[dom] > 0 else 0.0
    # the pipelined host-input leg is the end-to-end number when it ran (same bytes copied and read back per step)
    e2e_best = max([v for v in (e2e_value, e2e_pipe) if v] or [0.0]) or None
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": train_config(world),
        "workload_stats": {"R": R, "P_vis": Pv,
                           "shading": "all %d surfels (reference order)" % P_SURFELS if args.shade_all else
                           "the %d surfels that survive culling (preprocess runs first; images and gradients identical)" % Pv},
        "clocks": None,
        "e2e": {"value": round(e2e_best, 3) if e2e_best else None, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": 12, "inputs": "camera matrices + ground-truth image from pinned host memory; loss + "
                "num_rendered read back; parameters / light buffers resident (optimiser state)",
                "in_line": round(e2e_value, 3) if e2e_value else None,
                "prefetched": round(e2e_pipe, 3) if e2e_pipe else None,
                "note": "in_line: the step's inputs are copied right before its replay; prefetched: step k+1's inputs are "
                        "uploaded on a copy stream while step k computes (same bytes, same per-step read-back); value = the better"},
        "e2e_cold": {"value": round(e2e_cold, 3) if e2e_cold else None, "unit": UNIT, "h2d_bytes_per_step": int(cold_bytes),
                     "note": "worst case: every parameter and light buffer re-uploaded each step (eager path)"},
        "loss_tail": "torch mirror of svgss.py:187-294 (~120 elementwise kernels)" if args.torch_loss else
                     "fused resolve+loss kernels (csrc/resolve.cu), one per direction",
        "launch_mode": "eager" if runner is None else "cuda-graph (1 capture, %d svgir kernels/step)" % runner.launches_per_step,
        "kernel_timing": "CUDA events around each launch on the launching stream" + ("" if runner is None else
                         ", separate eager pass of the same kernels/inputs right after the timed region"),
        "gpu_launches": int(launches),
        "grad_allreduce": None if bucket is None else dict({
            "bytes": bucket.nbytes, "mode":
            (("svgir_peer_allreduce over NVLink peer memory (%s), in the step's graph: " % (
                "NVSwitch multicast ld_reduce/st" if peer.multicast else "128-bit peer loads/stores")) +
             ("one kernel at the end of the step" if args.reduce == "p2p" else
              "rasteriser-side segment on a side stream (%d CTAs) under the shading backward, shading-side segment after it"
              % peer.BG_GRID)) if peer is not None else
            (("2 segments issued inside the backward pass, captured in the step's graph; the overlapped one on a "
              "%d-CTA communicator" % args.bg_ctas if args.bg_ctas > 0 else
              "2 segments issued inside the backward pass, captured in the step's graph") if overlap else
             "one NCCL all-reduce after the step"), "note": reduce_note}, **(ar_check or {})),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": pk["hbm_gbs"],
                     "peak_source": pk_src, "unit": "GB/s", "frac": round(achieved / pk["hbm_gbs"], 4),
                     "traffic": traffic, "algorithmic_bytes": int(alg[dom]), "avg_ms": round(kt[dom], 4),
                     # the compositors are not HBM-bound: what the committed ncu capture of this
                     # kernel shows as its limiter (profiles/ncu_traffic.json <- profiles/r02/g9_ncu_step_full.txt)
                     "limiter": limiter},
        "kernels_ms": {k: round(v, 4) for k, v in kt.items()},
        # the same per-launch event times grouped by stage: svgss rasteriser forward+backward alone (BASELINE.json
        # configs[1]'s shape), the render_equation shading, the resolve+loss tail
        "stage_ms": {
            "svgss_fwd_bwd": round(sum(kt[k] for k in ("preprocess", "tile_scan", "emit", "sort_small", "composite_fwd",
                                                       "composite_bwd", "preprocess_bwd")) + kt["tile_scan"], 4),
            "render_equation_fwd_bwd": round(kt["shade_fwd"] + kt["shade_bwd"], 4),
            "loss_tail": round(kt["train_loss_fwd"] + kt["train_loss_bwd"], 4)},
    }
    # release the workload before the extra configurations run
    runner = None
    del pc, env, cam_dev, gt_dev, params, bucket, peer, host, gt_host
    gc.collect()
    torch.cuda.empty_cache()
    return line


# ---------------------------------------------------------------------------------------------
def measure_c4(args, ctx: Ctx, steps: int, warmup: int) -> dict:
    """C4 (BASELINE.json configs[3]): multi-view data-parallel training -- 1M SV surfels, 8 views of 800x800 per step,
    sharded over the ranks by view (rank r renders views r, r+N, ...; STRONG scaling: the step's work is fixed). Every
    rank replays ONE captured step graph per local view, accumulating into its flat gradient bucket (zero_in_graph=
    False), then the buckets are summed once: by the svgir peer-memory kernel (default) or NCCL (--reduce post).
    value = view-iterations per second = 8 x steps / time."""
    from svgir_b200 import _lib, pipeline, scene
    from svgir_b200 import dist as svdist
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    pipeline.SHADE_CULLED = bool(args.shade_all)
    pipeline.FUSED_LOSS = not args.torch_loss
    P, V = 1_000_000, 8
    cloud = scene.make_surfels(P, seed=1238)
    mats = scene.make_materials_torch(cloud, NS, 1239, dev)   # 2 GB of per-sample buffers: generated on the device
    pc = pipeline.model_from_scene(cloud, mats, dev)
    env = mats["env_param"].clone().requires_grad_(True)
    bg = torch.zeros(3, device=dev)
    cams = [pipeline.camera_from_scene(scene.look_at_camera(WIDTH, HEIGHT, v, V), dev) for v in range(V)]
    g = torch.Generator(device=dev)
    g.manual_seed(1240)
    gts = [torch.rand((3, HEIGHT, WIDTH), generator=g, device=dev) for _ in range(2)]
    params = pc.trainable() + [env]
    mine = svdist.views_for_rank(V, rank, world)

    # Exchange (N > 1). p2p-overlap (default): the LAST local view's graph carries the all-reduce -- the rasteriser-side
    # gradient segment on a side stream under that view's shading backward, the shading-side segment after it (the same
    # in-step exchange as the headline); the views before it replay an accumulate-only graph. p2p: one peer-memory
    # kernel after the local views; post: one NCCL all-reduce after them.
    peer, bucket, note = None, None, None
    in_step = world > 1 and args.reduce == "p2p-overlap" and len(mine) > 0
    if world > 1 and args.reduce != "post":
        if in_step:
            mk = lambda pr: svdist.FlatGradBucket(params, segments=pipeline.reduce_segments(pc), extra_floats=1,
                                                  alloc=pr.allocate, segment_peer=pr)
        else:
            mk = lambda pr: svdist.FlatGradBucket(params, alloc=pr.allocate, reducer=pr.all_reduce)
        peer, bucket, note = make_peer_bucket(ctx, mk)
    in_step = in_step and peer is not None
    if bucket is None:
        bucket = svdist.FlatGradBucket(params)
    runner = pipeline.GraphedTrainingStep(pc, env, bg, cams[0], gts[0], bucket=bucket, zero_in_graph=False)
    runner_last = pipeline.GraphedTrainingStep(pc, env, bg, cams[0], gts[0], bucket=bucket, zero_in_graph=False,
                                               reduce_in_graph=True, on_overflow="raise") if in_step else None
    redone = [0]

    def step(i):
        for _ in range(4):
            bucket.zero()
            R = 0
            loss = None
            try:
                for k, v in enumerate(mine):
                    r = runner_last if (in_step and k == len(mine) - 1) else runner
                    loss, res = r(cams[v], gts[(i + v) % 2])
                    R += int(res["num_rendered"])
            except pipeline.BinOverflow:   # raised on every rank together (the flag travels with the gradients)
                redone[0] += 1
                continue
            if not in_step:
                bucket.all_reduce()   # one exchange per step (no-op at N=1)
            return loss, R
        raise RuntimeError("C4: binning capacity did not converge")

    for i in range(warmup):
        loss, R = step(i)
    ctx.sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss, R = step(warmup + i)
    e1.record()
    ctx.sync_all()
    ms = ctx.max_ms(e0.elapsed_time(e1))
    # the exchange alone (all ranks enter together): bucket bytes / time
    ar_ms = None
    if world > 1:
        ctx.sync_all()
        e0.record()
        for _ in range(5):
            bucket.all_reduce()
        e1.record()
        ctx.sync_all()
        ar_ms = ctx.max_ms(e0.elapsed_time(e1)) / 5
    # per-kernel times of one local view, launched eagerly
    kt = {}
    if mine:
        _lib.timing_collect(reset=True)
        _lib.timing_enable(True)
        bucket.zero()
        pipeline.training_step(cams[mine[0]], pc, env, bg, gts[0], zero_grad=False)
        torch.cuda.synchronize()
        _lib.timing_enable(False)
        for k in ("composite_bwd", "composite_fwd", "shade_fwd", "shade_bwd", "preprocess", "preprocess_bwd", "emit", "sort_small"):
            t, n = _lib.timing_collect(k)
            kt[k] = round(t / max(n, 1), 4)
        _lib.timing_collect(reset=True)
    out = {
        "metric": METRIC, "value": round(V * steps / (ms / 1e3), 3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": round(ms / steps, 4), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4: multi-view data-parallel stage-2 training, %dk SV surfels, %d views of %dx%d per step, Ns=%d, "
                               "S=%d, VS=%d" % (P // 1000, V, WIDTH, HEIGHT, NS, S_FEAT, VS_FEAT),
                   "views_per_step": V, "views_per_rank": len(mine), "parallelism": f"view-dp{world}",
                   "l2": "working set > L2 (per-sample light buffers 2 GB/view)"},
        "R_last_step_local": R,
        "gpu_launches": int((runner.launches_per_step * (len(mine) - 1) + runner_last.launches_per_step) * steps) if in_step
        else int(runner.launches_per_step * len(mine) * steps + (steps if peer is not None else 0)),
        "steps_redone_after_overflow": redone[0],
        "grad_allreduce": None if world == 1 else {"bytes": bucket.nbytes, "ms": None if ar_ms is None else round(ar_ms, 4),
                                                   "mode": ("svgir_peer_allreduce inside the last local view's graph: rasteriser-side "
                                                            "segment under its shading backward, shading-side segment after it"
                                                            if in_step else "svgir_peer_allreduce after the local views")
                                                   if peer is not None else "one NCCL all-reduce after the local views", "note": note,
                                                   "ms_note": "ms = the whole bucket exchanged alone, all ranks entering together"},
        "kernels_ms": kt}
    runner = runner_last = None
    del pc, env, params, bucket, peer, mats
    from svgir_b200 import shading as _shading
    _shading.clear_env_tap_cache()   # its entries pin the model's direction buffers
    gc.collect()
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------
def measure_relight(args, ctx: Ctx, steps: int, warmup: int, clocks: ClockSampler = None, shape=None) -> dict:
    """C3-eval (BASELINE.json configs[2]): relighting frame = render_equation over all 300k surfels with
    Ns=384 samples under a fixed HDR env map (EnvLight semantics, scene/envmap.py:54-72) + svgss forward
    with the eval G-buffer (S=7, VS=64) at 800x800. Forward only; reports ms/frame. Under torchrun the
    view x envmap grid is sharded round-robin with no collective.
    shape = (label, P, W, H, n_views): C5 (BASELINE.json configs[4]) is the same frame at 2 M surfels, 1920x1080, with a
    200-view x 5-env-map sweep; a bounded number of its frames is timed and the sweep time follows from ms/frame."""
    label, P_SURFELS, WIDTH, HEIGHT, N_VIEWS = shape if shape is not None else ("C3-eval", 300_000, 800, 800, 8)
    from svgir_b200 import _lib, pipeline, scene, shading
    from svgir_b200 import dist as svdist
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    pipeline.SHADE_CULLED = bool(args.shade_all)
    pipeline.FUSED_LOSS = not args.torch_loss
    ns, n_env = 384, 5
    cloud = scene.make_surfels(P_SURFELS, seed=1236)
    mats = scene.make_materials_torch(cloud, ns, 1237, dev)   # 3.7 GB of per-sample buffers: generated on the device
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

    for i in range(warmup):
        res = frame(i)
    ctx.sync_all()
    _lib.launch_count(reset=True)
    _lib.timing_collect(reset=True)
    if runner is None:
        _lib.timing_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if clocks is not None:
        clocks.mark(True)
    e0.record()
    for i in range(steps):
        res = frame(warmup + i)
    e1.record()
    ctx.sync_all()
    if clocks is not None:
        clocks.mark(False)
    launches = _lib.launch_count() if runner is None else runner.launches_per_frame * steps
    if runner is not None:   # per-kernel times: the same frames launched eagerly, each launch bracketed by events
        _lib.timing_enable(True)
        for i in range(min(steps, 5)):
            eager_frame(warmup + i)
        torch.cuda.synchronize()
    _lib.timing_enable(False)
    kt = {}
    for k in ("shade_fwd", "composite_fwd", "preprocess", "emit", "sort_small", "tile_scan"):
        t, n = _lib.timing_collect(k)
        kt[k] = t / max(n, 1)
    _lib.timing_collect(reset=True)
    ms_frame = ctx.max_ms(e0.elapsed_time(e1)) / steps

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
        ctx.sync_all()
        e0.record()
        for i in range(steps):
            e2e_frame(2 + i)
        e1.record()
        ctx.sync_all()
        e2e_ms = ctx.max_ms(e0.elapsed_time(e1)) / steps
        e2e_h2d = cam_host[0].block.numel() * 4 + env_host[0].numel() * 4
        e2e_d2h = img_host.numel() * 4
    pk, pk_src = peaks()
    # shading runs on the surfels that survive the rasteriser's culling unless --shade-all
    n_sh = P_SURFELS if args.shade_all else int(res["visibility_filter"].sum())
    alg = n_sh * (ns * 32 + 124) + n_sh * 4 * (12 * 5 + 7)
    ach = alg / (kt["shade_fwd"] * 1e-3) / 1e9 if kt["shade_fwd"] > 0 else 0.0
    out = {
        "metric": "relight ms/frame", "value": round(ms_frame / world, 4), "unit": "ms/frame", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms_frame, 4), "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: relight frame = render_equation (Ns=%d) over %dk surfels + svgss forward "
                               "S=7/VS=64 at %dx%d, fixed HDR env map" % (label, ns, P_SURFELS // 1000, WIDTH, HEIGHT),
                   "grid": "%d views x %d env maps, round-robin over ranks" % (N_VIEWS, n_env),
                   "l2": "working set > L2 (light buffers %.1f GB resident)" % (P_SURFELS * ns * 32 / 1e9)},
        "sweep": {"frames": N_VIEWS * n_env, "seconds_at_this_rate": round(N_VIEWS * n_env * ms_frame / world / 1e3, 3),
                  "frames_per_s_all_ranks": round(world * 1e3 / ms_frame, 1)},
        "workload_stats": {"R": int(res["num_rendered"]), "surfels_shaded": n_sh},
        "gpu_launches": int(launches),
        "launch_mode": "eager" if runner is None else "cuda-graph (1 capture, %d svgir kernels/frame)" % runner.launches_per_frame,
        "e2e": None if e2e_ms is None else {"value": round(e2e_ms / world, 4), "unit": "ms/frame", "h2d_bytes_per_step": int(e2e_h2d),
                                            "d2h_bytes_per_step": int(e2e_d2h),
                                            "inputs": "camera block + env map from pinned host memory; relit image read back"},
        "roofline": {"bound": "hbm", "kernel": "shade_fwd", "achieved": round(ach, 1), "peak": pk["hbm_gbs"],
                     "peak_source": pk_src, "unit": "GB/s", "frac": round(ach / pk["hbm_gbs"], 4), "traffic": None,
                     "algorithmic_bytes": int(alg), "avg_ms": round(kt["shade_fwd"], 4)},
        "kernels_ms": {k: round(v, 4) for k, v in kt.items()}}
    runner = None
    del pc, mats
    from svgir_b200 import shading as _shading
    _shading.clear_env_tap_cache()   # its entries pin the model's direction buffers
    gc.collect()
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------
def measure_visibility(args, ctx: Ctx, reps: int = 5) -> dict:
    """Visibility precompute (BASELINE.md 'R-bvh'): LBVH build over 300k surfels + one opacity ray per
    (surfel, incident sample) for Ns = 64 (300k surfels) and Ns = 384 (100k surfels, one eval chunk), our kernels
    (csrc/bvh.cu) beside the reference's (submodules/bvh/src/{construct,trace}.cu compiled into
    oracle/_ref/libbvh_ref.so) on the same inputs. Only the reference leg touches oracle/."""
    from svgir_b200 import bvh as svbvh, scene
    dev = ctx.dev
    cloud = scene.make_surfels(P_SURFELS, seed=1234)
    d = lambda a: torch.from_numpy(a).to(dev)
    means, scales, rots, opac = d(cloud.means3D), d(cloud.scales), d(cloud.rotations), d(cloud.opacity)
    normals = d(cloud.normals)
    # The rasteriser scene is 1e-6 thin along the normal: Sigma^-1 would carry 1e12 entries and the trace's
    # fp32 quadratic forms would be rounding noise (Appendix C 17). The reference's tracer sees the LEARNED third scale,
    # which is initialised equal to the other two (gaussian_model.py:707) and never updated by the surface rasteriser
    # (quirk 1), so the tracing benches use scales.z = scales.x. Both legs get the same tensors.
    scales = scales.clone()
    scales[:, 2] = scales[:, 0]
    # Sigma^-1 = (R diag(1/s)) (R diag(1/s))^T, upper triangle (gaussian_model.py:379-382)
    q = rots / rots.norm(dim=-1, keepdim=True)
    r, x, y, z = q.unbind(-1)
    Rm = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                      2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                      2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    Lm = Rm * (1.0 / scales)[:, None, :]
    Mi = Lm @ Lm.transpose(1, 2)
    symm_inv = torch.stack([Mi[:, 0, 0], Mi[:, 0, 1], Mi[:, 0, 2], Mi[:, 1, 1], Mi[:, 1, 2], Mi[:, 2, 2]], -1).contiguous()
    RefBvh = None
    try:
        from oracle import ref_cuda
        if ref_cuda.available("bvh"):
            RefBvh = ref_cuda.RefBvh
    except Exception:
        RefBvh = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn, n):
        for _ in range(3):   # allocator / lazy module loading warm-up: the build is a dozen short launches
            r = fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, r

    out = {"surfels": P_SURFELS, "note": "CUDA events, resident inputs, mean of %d repetitions after one warm-up; build = leaf "
                                          "boxes + construct_bvh (submodules/bvh/__init__.py:29-60), trace = trace_bvh_opacity with "
                                          "the application's 0.05 origin offset (:62-71)" % reps}
    nodes0, aabbs0 = svbvh.leaf_aabbs(means, scales, rots)

    def build_ours():
        n_, a_ = svbvh.leaf_aabbs(means, scales, rots)
        return svbvh.Bvh(n_, a_)

    build_ms, tree = timeit(build_ours, reps)
    out["build_ms"] = round(build_ms, 4)
    rtree = None
    if RefBvh is not None:
        rb, rtree = timeit(lambda: RefBvh(nodes0, aabbs0, means, scales, rots), reps)
        out["reference_build_ms"] = round(rb, 4)
        out["build_speedup"] = round(rb / build_ms, 2)
        # the reference's bottom-up box merge has no memory fence between a thread's box store and its atomicCAS on the
        # parent (construct.cu:240-258): a few dozen internal boxes per build come out as a stale partial merge (different
        # ones every run), and rays through them lose hits. Node links / Morton codes are bit-equal. The traces below are
        # compared over the SAME race-free boxes (tests/test_bvh_gpu.py does the same); the stale count is reported.
        out["reference_stale_boxes"] = int((tree.aabbs != rtree.aabbs).any(-1).sum())
        out["nodes_bit_equal"] = bool(torch.equal(tree.nodes, rtree.nodes))
        rtree.aabbs.copy_(tree.aabbs)
    for ns in (64, 384):
        n_s = 100_000 if ns == 384 else P_SURFELS
        dirs_np, _ = scene.fibonacci_hemisphere_dirs(cloud.normals[:n_s], ns)
        dirs = torch.from_numpy(dirs_np).to(dev)
        orig = (means[:n_s, None, :] + 0.05 * dirs).contiguous()
        t_ms, (_, vis) = timeit(lambda: tree.trace_opacity(orig, dirs, means, symm_inv, opac, normals), reps)
        key = "trace_ns%d" % ns
        nr = n_s * ns
        out[key] = {"rays": nr, "ms": round(t_ms, 4), "mrays_per_s": round(nr / t_ms / 1e3, 1)}
        if rtree is not None:
            r_ms, (_, rvis) = timeit(lambda: rtree.trace_opacity(orig, dirs, means, symm_inv, opac, normals), max(reps // 2, 1))
            out[key]["reference_ms"] = round(r_ms, 4)
            out[key]["speedup"] = round(r_ms / t_ms, 2)
            # the trace stops a ray (visibility 0) once its transmittance drops to 0.9: a ray whose product lands within an ulp
            # of the threshold flips between "T" and 0, so the difference is reported as the fraction of rays that disagree
            # (tests/test_bvh_gpu.py holds the survivors to 1e-7) and the largest difference among the rays both keep
            a, b = vis.reshape(-1), rvis.reshape(-1)
            both = (a > 0) & (b > 0)
            out[key]["rays_flipped_at_threshold"] = int(((a > 0) != (b > 0)).sum())
            out[key]["mismatch_fraction"] = float(((a - b).abs() > 1e-6).float().mean())
            out[key]["max_abs_diff_surviving_rays"] = float((a[both] - b[both]).abs().max()) if bool(both.any()) else 0.0
        del dirs, orig
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------
_REF_CACHE = {}


def measure_radiance(args, ctx: Ctx, steps: int = 10, warmup: int = 3) -> dict:
    """the radiance cache (update_radiace: tree + 64 hit-to-hit rays per surfel) and the training step with
    the radiance-consistency term in it (svgss.py:319-320, lambda_radiance = 0.05), on the headline scene."""
    from svgir_b200 import _lib, pipeline, radiance
    dev = ctx.dev
    cloud, mats, cams, gts = build_host_workload()
    pc = pipeline.model_from_scene(cloud, mats, dev)
    env = torch.from_numpy(mats["env_param"]).to(dev).requires_grad_(True)
    bg = torch.zeros(3, device=dev)
    cam_dev = [pipeline.camera_from_scene(c, dev) for c in cams]
    gt_dev = [torch.from_numpy(g).to(dev) for g in gts]
    with torch.no_grad():
        q = pc.rotation / pc.rotation.norm(dim=1, keepdim=True)
        r, x, y, z = q.unbind(1)
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                         2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                         2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
        sc = pc.scaling.detach().clone()
        sc[:, 2] = sc[:, 0]     # as in measure_visibility: the tracer sees the learned third scale (= the others at init)
        Sinv = R @ torch.diag_embed(1.0 / sc ** 2) @ R.transpose(1, 2)
        ci = torch.stack([Sinv[:, 0, 0], Sinv[:, 0, 1], Sinv[:, 0, 2], Sinv[:, 1, 1], Sinv[:, 1, 2], Sinv[:, 2, 2]], 1).contiguous()
        gn = R[:, :, 2].contiguous()
    rc = radiance.RadianceCache()
    torch.manual_seed(11)
    rc.update(pc.xyz, sc, pc.rotation, pc.opacity, gn, ci, pc.shs, sample_num=NS)   # warm-up (allocations)
    torch.cuda.synchronize()
    _lib.timing_collect(reset=True)
    _lib.timing_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc.update(pc.xyz, sc, pc.rotation, pc.opacity, gn, ci, pc.shs, sample_num=NS)
    e1.record()
    torch.cuda.synchronize()
    _lib.timing_enable(False)
    trace_ms = _lib.timing_collect("radiance_cache")[0]
    out = {"cache": {"update_ms": round(e0.elapsed_time(e1), 2), "trace_kernel_ms": round(trace_ms, 2),
                     "rays": P_SURFELS * NS, "rays_per_s": round(P_SURFELS * NS / (trace_ms * 1e-3)) if trace_ms else None,
                     "first_hit_fraction": round(float((rc.hemi_index_buffers >= 0).float().mean()), 4),
                     "occluded_fraction": round(float((rc.visibility_tracing == 0).float().mean()), 4)}}
    runner = pipeline.GraphedTrainingStep(pc, env, bg, cam_dev[0], gt_dev[0], radiance_cache=rc, lambda_radiance=0.05)
    for i in range(warmup):
        runner(cam_dev[i % N_VIEWS], gt_dev[i % len(gt_dev)])
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        runner(cam_dev[(warmup + i) % N_VIEWS], gt_dev[i % len(gt_dev)])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    _lib.timing_collect(reset=True)
    _lib.timing_enable(True)
    for i in range(3):
        runner.fs.enqueue()
    torch.cuda.synchronize()
    _lib.timing_enable(False)
    k = {}
    for name in ("radiance_loss_fused",):
        t, n = _lib.timing_collect(name)
        k[name] = round(t / n, 4) if n else None
    out["step_with_term"] = {"value": round(1000.0 / ms, 2), "unit": "it/s", "ms_per_step": round(ms, 4), "steps": steps,
                             "warmup": warmup, "lambda_radiance": 0.05, "kernels_ms": k,
                             "loss_radiance": float(runner.fs.result["loss_radiance"]),
                             "note": "the headline step plus get_radiance_loss; select+forward and backward kernels run on "
                                     "the side stream under the rasteriser"}
    out["config"] = {"workload": "C3-train + radiance term: %dk surfels x %d samples, env %dx%d" % (
        P_SURFELS // 1000, NS, int(env.shape[-3]), int(env.shape[-2])), "parity": "unpinned (Slang "
                     "kernels cannot run here; numpy restatement oracle/radiance_oracle.py)"}
    return out


def _ref_inputs():
    if "w" not in _REF_CACHE:
        _REF_CACHE["w"] = build_host_workload()
    return _REF_CACHE["w"]


def reference_cpu_step(view: int) -> float:
    """One FULL step of the workload on the host CPU with the oracle: shading fwd+bwd of every surfel (the reference's
    order: all surfels are shaded), preprocess + binning + compositing fwd+bwd of every tile. Returns seconds."""
    from oracle import svgss as O, shading_oracle as SO
    cloud, mats, cams, gts = _ref_inputs()
    cam = cams[view % N_VIEWS]
    t0 = time.perf_counter()
    tt = {k: torch.from_numpy(mats[k]) for k in
          ("base_color", "roughness", "shading_normals", "radiance", "visibility", "incident_dirs", "incident_areas")}
    envp = torch.from_numpy(mats["env_param"]).requires_grad_(True)
    for k in ("base_color", "roughness", "shading_normals"):
        tt[k] = tt[k].clone().requires_grad_(True)
    vd = torch.from_numpy(cam.campos[None] - cloud.means3D)
    vd = torch.nn.functional.normalize(vd, dim=-1)
    pbr, extra = SO.rendering_equation4(tt["base_color"], tt["roughness"], tt["shading_normals"], vd, tt["radiance"],
                                        lambda d: SO.direct_light_learnable(envp, d), tt["visibility"],
                                        tt["incident_dirs"], tt["incident_areas"])
    feats, vfeats = SO.pack_features(pbr, extra, tt["base_color"], tt["roughness"], tt["shading_normals"],
                                     torch.from_numpy(cam.viewmatrix[:3, :3].copy()), True)
    f = np.ascontiguousarray(feats.detach().numpy())
    vf = np.ascontiguousarray(vfeats.detach().numpy())
    O.lib().oracle_set_tile_step(1)
    fw = O.forward(cam, cloud.means3D, cloud.opacity, cloud.scales, cloud.rotations, f, vf, shs=cloud.shs)
    g = [np.full(fw[k].shape, 1.0 / (HEIGHT * WIDTH), np.float32) for k in
         ("color", "normal_img", "depth", "opacity", "feature", "vfeature")]
    bw = O.backward(fw, *g)
    # `features` (mean visibility / mean cached radiance) carry no trainable input: the gradient flows through vfeatures
    torch.autograd.backward([vfeats],
                            [torch.from_numpy(np.ascontiguousarray(bw["dL_dvfeatures"], dtype=np.float32).reshape(vfeats.shape))])
    return time.perf_counter() - t0


def cpu_reference_run(steps: int, warmup: int) -> dict:
    torch.set_num_threads(os.cpu_count() or 1)
    for i in range(warmup):
        reference_cpu_step(i)
    ts = [reference_cpu_step(warmup + i) for i in range(steps)]
    t = float(np.mean(ts))
    return {"value": round(1.0 / t, 6), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": ("oracle (C + torch-CPU restatement of the reference path): %d FULL step(s) of the workload -- all %dk surfels "
                       "shaded fwd+bwd (torch CPU, all threads), preprocessed and binned, every tile composited fwd+bwd (C, OpenMP); "
                       "%.2f s per step" % (steps, P_SURFELS // 1000, t)),
            "step_seconds": round(t, 3)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cb = cpu_reference_run(args.steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(cb["step_seconds"] * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": train_config(world), "cpu_baseline": cb,
            "note": "rank 0 only: one view per step on the host CPU, every step a full step; warm-up capped at 1 step",
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def measure_reference_cuda(steps: int, warmup: int) -> dict:
    """The reference's own CUDA rasteriser (oracle/_ref, unmodified sources, sm_100 build) plus the torch shading graph of
    the reference (its restatement oracle/shading_oracle.py, pinned to goldens from the reference's rendering_equation4:
    /root/reference is not on the GPU box), both on this GPU (BASELINE.md 'R-step'): the denominator of the >=8x target."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import ref_cuda, shading_oracle as SO
    if not (torch.cuda.is_available() and ref_cuda.available()):
        return {"impl": "reference_cuda", "unavailable": "needs a GPU and oracle/_ref/libsvgss_ref.so"}
    dev = torch.device("cuda", torch.cuda.current_device())
    cloud, mats, cams, gts = _ref_inputs()
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

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out = {"impl": "reference_cuda", "metric": METRIC, "value": round(steps / (ms / 1e3), 3), "unit": UNIT,
           "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": round(ms / steps, 3),
           "config": {"workload": WORKLOAD}, "note": "reference CUDA rasteriser (sm_100 build of the unmodified sources, "
           "oracle/_ref/libsvgss_ref.so) + the torch restatement of the reference's shading graph (oracle/shading_oracle.py; "
           "the reference Python cannot travel to the GPU box) on the same B200, same inputs as the headline"}
    del t, geo, r
    from svgir_b200 import shading as _shading
    _shading.clear_env_tap_cache()   # its entries pin the model's direction buffers
    gc.collect()
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------
def finish_process(ctx: Ctx):
    """Leave without tearing NCCL down: with collectives / peer kernels captured in CUDA graphs the communicator's
    destructor was seen to block (gpurun_out/s2)."""
    sys.stdout.flush()
    sys.stderr.flush()
    if ctx.world > 1:
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
        os._exit(0)


def run_ours(args):
    ctx = Ctx()
    clocks = ClockSampler(ctx.local)
    if ctx.rank == 0:
        clocks.start()   # before the warm-up, so that samples exist by the time the (short) timed region runs
    line = measure_train(args, ctx, clocks)
    if ctx.rank == 0:
        line["clocks"] = clocks.stop()

    printed = threading.Event()

    def emit():
        if not printed.is_set():
            printed.set()
            if ctx.rank == 0:
                print(json.dumps(line), flush=True)

    def watchdog():
        line["extras_error"] = "an extra workload did not finish within %d s; the headline is unaffected" % EXTRAS_DEADLINE_S
        emit()
        os._exit(0)

    if not args.no_extras:
        timer = threading.Timer(EXTRAS_DEADLINE_S, watchdog)
        timer.daemon = True
        timer.start()

        def extra(name, fn):
            t0 = time.perf_counter()
            try:
                r = fn()
            except Exception as e:  # noqa: BLE001
                r = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            if isinstance(r, dict):
                r["wall_s"] = round(time.perf_counter() - t0, 1)
            line[name] = r

        # every N: C4 (1M surfels, 8 views/step, strong scaling) so that its efficiency can be read from the scaling run
        extra("c4", lambda: {k: v for k, v in measure_c4(args, ctx, steps=5, warmup=3).items()
                             if k in ("value", "unit", "ms_per_step", "steps", "warmup", "scaling", "config", "grad_allreduce",
                                      "kernels_ms", "gpu_launches", "R_last_step_local")})
        # every N: C5 (2 M surfels, 1920x1080, 200 views x 5 env maps sharded view x env map over the ranks, no collective)
        extra("c5", lambda: {k: v for k, v in measure_relight(args, ctx, steps=8, warmup=2, shape=C5_SHAPE).items()
                             if k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "workload_stats", "sweep",
                                      "e2e", "kernels_ms", "launch_mode")})
        if ctx.world == 1:
            extra("relight", lambda: {k: v for k, v in measure_relight(args, ctx, steps=10, warmup=3).items()
                                      if k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "workload_stats", "e2e",
                                               "roofline", "kernels_ms", "launch_mode")})
            extra("visibility", lambda: measure_visibility(args, ctx))
            extra("radiance", lambda: measure_radiance(args, ctx))

            def refcuda():
                r = measure_reference_cuda(steps=5, warmup=2)
                if r.get("value"):
                    r["ratio"] = round(line["value"] / r["value"], 2)
                    if line["e2e"]["value"]:
                        r["e2e_ratio"] = round(line["e2e"]["value"] / r["value"], 2)
                return r
            extra("reference_cuda", refcuda)
        timer.cancel()
    if ctx.world == 1 and not args.no_cpu_baseline and ctx.rank == 0:
        line["cpu_baseline"] = cpu_reference_run(2, 0)
    emit()
    finish_process(ctx)


def run_workload(args):
    ctx = Ctx()
    clocks = ClockSampler(ctx.local)
    if ctx.rank == 0:
        clocks.start()
    if args.workload == "relight":
        out = measure_relight(args, ctx, args.steps, args.warmup, clocks)
    elif args.workload == "c5":
        out = measure_relight(args, ctx, args.steps, args.warmup, clocks, shape=C5_SHAPE)
    elif args.workload == "c4":
        clocks.mark(True)
        out = measure_c4(args, ctx, args.steps, args.warmup)
        clocks.mark(False)
    elif args.workload == "radiance":
        out = {"metric": "radiance cache + training step with the radiance-consistency term", "n_gpus": 1,
               "radiance": measure_radiance(args, ctx, args.steps, args.warmup)}
    else:
        out = {"metric": "visibility precompute (LBVH build + opacity trace)", "n_gpus": 1, "visibility": measure_visibility(args, ctx)}
    if ctx.rank == 0:
        out["clocks"] = clocks.stop()
        print(json.dumps(out), flush=True)
    finish_process(ctx)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference_cuda"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra keys of the default line (c4, relight, visibility, "
                    "reference_cuda)")
    ap.add_argument("--shade-all", action="store_true", help="shade culled surfels too (reference order: shading before the rasteriser)")
    ap.add_argument("--torch-loss", action="store_true", help="resolve + loss tail in torch (the reference's ~120 kernels) instead of the fused kernels")
    ap.add_argument("--reduce", default="p2p-overlap", choices=["overlap", "post", "p2p", "p2p-overlap"],
                    help="N>1 gradient exchange. p2p-overlap (default): svgir kernels over NVLink peer memory inside the step's graph, "
                         "the rasteriser-side segment on a side stream under the shading backward, the shading-side segment after it; "
                         "p2p: one such kernel at the end of the step; both fall back to `post` if the box has no peer-mapped "
                         "memory. post: one NCCL all-reduce after the graph. overlap: NCCL, segment-wise inside the backward pass "
                         "(slower: the NCCL kernel takes SMs from the shading backward)")
    ap.add_argument("--bg-ctas", type=int, default=4, help="--reduce overlap: CTA limit of the communicator that carries the "
                    "segment overlapped with the shading backward (0 = default communicator for both segments)")
    ap.add_argument("--eager", action="store_true", help="launch the step kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--workload", default="train", choices=["train", "relight", "c4", "c5", "visibility", "radiance"],
                    help="train = C3-train fwd+bwd it/s (headline, plus the extra keys); relight = C3-eval forward ms/frame (Ns=384, "
                         "S=7, VS=64); c4 = 1M surfels, 8 views per step sharded over the ranks (strong scaling); visibility = LBVH "
                         "build + opacity trace vs the reference kernels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference_cuda":
        print(json.dumps(measure_reference_cuda(args.steps, args.warmup)), flush=True)
    elif args.workload == "train":
        run_ours(args)
    else:
        run_workload(args)


if __name__ == "__main__":
    main()
