"""View-sharded data parallel step over NCCL on 2 GPUs of one box: each rank renders its own
view inside a CUDA graph whose backward pass issues the segment-wise gradient all-reduce; the reduced
gradient on every rank must equal the SUM of the two views' gradients computed locally (eagerly, no
collective) on that same rank -- so the check needs no cross-process comparison. Also exercises the
matched re-capture: one rank's binning overflow must make BOTH ranks re-capture (the overflow flag
travels in the last gradient segment). Skipped with fewer than 2 GPUs (`gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        from svgir_b200 import dist as D, pipeline, raster, scene
        r, w, dev = D.init_from_env()
        P, W, H, Ns = 6000, 160, 128, 16
        cloud = scene.make_surfels(P, seed=11)
        mats = scene.make_materials(cloud, Ns, seed=12, env_hw=(16, 32))
        cams = [pipeline.camera_from_scene(scene.look_at_camera(W, H, v, 4), dev) for v in range(4)]
        gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(i)).to(dev) for i in range(2)]
        bg = torch.zeros(3, device=dev)

        def model():
            pc = pipeline.model_from_scene(cloud, mats, dev)
            env = torch.from_numpy(mats["env_param"]).to(dev).requires_grad_(True)
            return pc, env

        # local reference: both views rendered eagerly on this rank, gradients summed by autograd
        pc_e, env_e = model()
        pc_g, env_g = model()
        params = pc_g.trainable() + [env_g]
        bucket = D.FlatGradBucket(params, segments=pipeline.reduce_segments(pc_g), extra_floats=1)
        runner = pipeline.GraphedTrainingStep(pc_g, env_g, bg, cams[0], gts[0], bucket=bucket, reduce_in_graph=True)
        worst = 0.0
        for step in range(3):
            views = [(2 * step + k) % 4 for k in range(world)]
            for t in pc_e.trainable() + [env_e]:
                t.grad = None
            for v in views:
                pipeline.training_step(cams[v], pc_e, env_e, bg, gts[v % 2], zero_grad=False)
            runner(cams[views[rank]], gts[views[rank] % 2])
            torch.cuda.synchronize()
            assert sorted(bucket.overlap_log) == [0, 1]
            for a, b in zip(params, pc_e.trainable() + [env_e]):
                rel = float((a.grad - b.grad).norm() / b.grad.norm().clamp_min(1e-20))
                worst = max(worst, rel)
        caps_before = runner.captures
        # Matched re-capture. Both ranks drop their graph; rank 1 re-captures with NO slack in its binning capacity.
        # Then the surfels grow 1.2x (scaling is a per-step input of the graph): rank 1 overflows, rank 0 (2x slack)
        # does not -- the flag in the last gradient segment must make BOTH re-capture and re-run together.
        old = (raster.ASYNC_SLACK, raster.ASYNC_MARGIN)
        if rank == 1:
            raster._CAP_HINT.clear()
            raster.ASYNC_SLACK, raster.ASYNC_MARGIN = 1.0, 16
        runner.graph = None
        runner(cams[rank], gts[rank])
        raster.ASYNC_SLACK, raster.ASYNC_MARGIN = old
        with torch.no_grad():
            pc_g.scaling.mul_(1.2)
            pc_e.scaling.mul_(1.2)
        runner(cams[rank], gts[rank])
        torch.cuda.synchronize()
        for t in pc_e.trainable() + [env_e]:
            t.grad = None
        for v in range(world):
            pipeline.training_step(cams[v], pc_e, env_e, bg, gts[v], zero_grad=False)
        for a, b in zip(params, pc_e.trainable() + [env_e]):
            worst = max(worst, float((a.grad - b.grad).norm() / b.grad.norm().clamp_min(1e-20)))
        q.put((rank, worst, runner.captures - caps_before, None))
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    except Exception as e:  # surface the failure instead of a hang
        import traceback
        q.put((rank, 1e9, -1, traceback.format_exc()))


def test_two_rank_graphed_step_overlapped_allreduce():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted(q.get(timeout=300) for _ in range(world))
    finally:   # never leave a rank spinning on the GPU
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for rank, worst, recaptures, err in res:
        assert err is None, err
        assert worst < 1e-3, (rank, worst)
        # one capture after the forced graph drop, one more on rank 1's overflow -- on BOTH ranks
        assert recaptures == 2, (rank, recaptures)


def _peer_worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        import torch.distributed as dist
        from svgir_b200 import dist as D, pipeline, scene
        r, w, dev = D.init_from_env()
        out = {}
        for mc in ("1", "0"):   # NVSwitch multicast path (when the box has it), then the peer load/store path
            os.environ["SVGIR_PEER_MULTICAST"] = mc
            peer = D.PeerAllReduce(dev)
            n = 1_000_003 * 4
            flat = peer.allocate(n)
            g = torch.Generator(device=dev).manual_seed(100 + rank)
            worst = 0.0
            for it in range(3):
                x = torch.randn(n, device=dev, generator=g)
                flat.copy_(x)
                ref = x.clone()
                dist.all_reduce(ref)            # NCCL as the checker
                peer.all_reduce()
                worst = max(worst, float((flat - ref).abs().max() / ref.abs().max()))
            # the same launch replayed from a CUDA graph (the flags return to zero after every launch)
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                peer.all_reduce()               # warm-up on the side stream
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            flat.fill_(float(rank + 1))
            with torch.cuda.graph(graph):
                peer.all_reduce()
            for it in range(3):
                flat.fill_(float(rank + 1 + it))
                torch.cuda.synchronize()
                dist.barrier()
                graph.replay()
                torch.cuda.synchronize()
                want = sum(float(k + 1 + it) for k in range(world))
                worst = max(worst, float((flat - want).abs().max()))
            out[mc] = (worst, peer.multicast)
            del graph
            torch.cuda.synchronize()
            dist.barrier()

        # the graphed data-parallel step with the peer all-reduce recorded at the end of the graph
        os.environ["SVGIR_PEER_MULTICAST"] = "1"
        P, W, H, Ns = 6000, 160, 128, 16
        cloud = scene.make_surfels(P, seed=11)
        mats = scene.make_materials(cloud, Ns, seed=12, env_hw=(16, 32))
        cams = [pipeline.camera_from_scene(scene.look_at_camera(W, H, v, 4), dev) for v in range(4)]
        gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(i)).to(dev) for i in range(2)]
        bg = torch.zeros(3, device=dev)

        def model():
            pc = pipeline.model_from_scene(cloud, mats, dev)
            env = torch.from_numpy(mats["env_param"]).to(dev).requires_grad_(True)
            return pc, env

        pc_e, env_e = model()
        pc_g, env_g = model()
        params = pc_g.trainable() + [env_g]
        peer = D.PeerAllReduce(dev)
        bucket = D.FlatGradBucket(params, extra_floats=1, alloc=peer.allocate, reducer=peer.all_reduce)
        runner = pipeline.GraphedTrainingStep(pc_g, env_g, bg, cams[0], gts[0], bucket=bucket, reduce_in_graph=True)
        step_worst = 0.0
        for step in range(3):
            views = [(2 * step + k) % 4 for k in range(world)]
            for t in pc_e.trainable() + [env_e]:
                t.grad = None
            for v in views:
                pipeline.training_step(cams[v], pc_e, env_e, bg, gts[v % 2], zero_grad=False)
            runner(cams[views[rank]], gts[views[rank] % 2])
            torch.cuda.synchronize()
            for a, b in zip(params, pc_e.trainable() + [env_e]):
                step_worst = max(step_worst, float((a.grad - b.grad).norm() / b.grad.norm().clamp_min(1e-20)))
        # segment-wise exchange: the rasteriser-side gradients are summed on a side stream (16 CTAs) while the shading
        # backward runs on the SMs left to it, the shading-side segment after it -- all inside the step's graph
        pc_s, env_s = model()
        params_s = pc_s.trainable() + [env_s]
        peer_s = D.PeerAllReduce(dev)
        bucket_s = D.FlatGradBucket(params_s, segments=pipeline.reduce_segments(pc_s), extra_floats=1,
                                    alloc=peer_s.allocate, segment_peer=peer_s)
        runner_s = pipeline.GraphedTrainingStep(pc_s, env_s, bg, cams[0], gts[0], bucket=bucket_s, reduce_in_graph=True)
        for step in range(3):
            views = [(2 * step + k + 1) % 4 for k in range(world)]
            for t in pc_e.trainable() + [env_e]:
                t.grad = None
            for v in views:
                pipeline.training_step(cams[v], pc_e, env_e, bg, gts[v % 2], zero_grad=False)
            runner_s(cams[views[rank]], gts[views[rank] % 2])
            torch.cuda.synchronize()
            assert runner_s.fused, "the peer-segment exchange is part of the fused step (fused_step.FusedTrainStep)"
            for a, b in zip(params_s, pc_e.trainable() + [env_e]):
                step_worst = max(step_worst, float((a.grad - b.grad).norm() / b.grad.norm().clamp_min(1e-20)))
        # matched re-capture of the fused step: rank 1 re-captures with no slack in its bins, then the surfels grow 1.2x
        # -- rank 1 overflows, rank 0 does not; the flag summed with the last gradient segment makes BOTH re-run
        from svgir_b200 import raster
        caps = runner_s.captures
        old = (raster.ASYNC_SLACK, raster.ASYNC_MARGIN)
        if rank == 1:
            raster._CAP_HINT.clear()   # the eager reference steps above left a roomy capacity hint for this shape
            raster.ASYNC_SLACK, raster.ASYNC_MARGIN = 1.0, 16
        runner_s.graph, runner_s.fs = None, None
        runner_s(cams[rank], gts[rank])
        raster.ASYNC_SLACK, raster.ASYNC_MARGIN = old
        with torch.no_grad():
            pc_s.scaling.mul_(1.2)
            pc_e.scaling.mul_(1.2)
        runner_s(cams[rank], gts[rank])
        torch.cuda.synchronize()
        assert runner_s.captures - caps == 2, (rank, runner_s.captures - caps)
        for t in pc_e.trainable() + [env_e]:
            t.grad = None
        for v in range(world):
            pipeline.training_step(cams[v], pc_e, env_e, bg, gts[v], zero_grad=False)
        for a, b in zip(params_s, pc_e.trainable() + [env_e]):
            step_worst = max(step_worst, float((a.grad - b.grad).norm() / b.grad.norm().clamp_min(1e-20)))
        # several views per rank (C4): the views before the last replay an accumulate-only graph, the LAST one carries the
        # exchange in its graph (on_overflow="raise": after the in-place all-reduce a re-run of one view would count the
        # other ranks' sums twice, so the step is redone). Rank 1 captures with no slack, then the surfels grow: its last
        # view overflows, BOTH ranks raise BinOverflow and redo the step.
        pc_m, env_m = model()
        with torch.no_grad():
            pc_m.scaling.mul_(1.2)
        params_m = pc_m.trainable() + [env_m]
        peer_m = D.PeerAllReduce(dev)
        bucket_m = D.FlatGradBucket(params_m, segments=pipeline.reduce_segments(pc_m), extra_floats=1,
                                    alloc=peer_m.allocate, segment_peer=peer_m)
        run_acc = pipeline.GraphedTrainingStep(pc_m, env_m, bg, cams[0], gts[0], bucket=bucket_m, zero_in_graph=False)
        if rank == 1:
            raster._CAP_HINT.clear()
            raster.ASYNC_SLACK, raster.ASYNC_MARGIN = 1.0, 16
        run_last = pipeline.GraphedTrainingStep(pc_m, env_m, bg, cams[0], gts[0], bucket=bucket_m, zero_in_graph=False,
                                                reduce_in_graph=True, on_overflow="raise")
        mine = [rank, rank + 2]
        redone = 0

        def multi_step():
            nonlocal redone
            for _ in range(4):
                bucket_m.zero()
                try:
                    run_acc(cams[mine[0]], gts[mine[0] % 2])
                    run_last(cams[mine[1]], gts[mine[1] % 2])
                    return
                except pipeline.BinOverflow:
                    redone += 1
            raise RuntimeError("did not converge")

        def check_multi():
            nonlocal step_worst
            torch.cuda.synchronize()
            for t in pc_e.trainable() + [env_e]:
                t.grad = None
            for v in range(4):
                pipeline.training_step(cams[v], pc_e, env_e, bg, gts[v % 2], zero_grad=False)
            for a, b in zip(params_m, pc_e.trainable() + [env_e]):
                step_worst = max(step_worst, float((a.grad - b.grad).norm() / b.grad.norm().clamp_min(1e-20)))

        multi_step()
        raster.ASYNC_SLACK, raster.ASYNC_MARGIN = old
        check_multi()
        assert redone == 0, redone
        with torch.no_grad():
            pc_m.scaling.mul_(1.15)
            pc_e.scaling.mul_(1.15)
        multi_step()
        check_multi()
        assert redone == 1, (rank, redone)   # rank 1's bins overflowed; rank 0 saw the summed flag
        q.put((rank, out, step_worst, None))
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)
    except Exception:
        import traceback
        q.put((rank, None, 1e9, traceback.format_exc()))


def test_two_rank_peer_memory_allreduce():
    """svgir_peer_allreduce (one kernel over NVLink peer memory) against NCCL on random data, replayed from a CUDA
    graph, and as the gradient exchange recorded at the end of the graphed data-parallel step."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted(q.get(timeout=300) for _ in range(world))
    finally:   # never leave a rank spinning on the GPU
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for rank, out, step_worst, err in res:
        assert err is None, err
        for mc, (worst, used_mc) in out.items():
            assert worst < 1e-6, (rank, mc, worst, used_mc)
        assert step_worst < 1e-3, (rank, step_worst)
