"""View-sharded data parallel host logic on CPU: gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from svgir_b200 import dist as D


def test_view_partition_covers_every_view_once():
    for n_views in (1, 7, 8, 200):
        for world in (1, 2, 4, 8):
            got = sorted(v for r in range(world) for v in D.views_for_rank(n_views, r, world))
            assert got == list(range(n_views))
            sizes = [len(D.views_for_rank(n_views, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_relight_grid_partition():
    # C5 shape: 200 views x 5 env maps over 8 ranks (eval_relighting_tensoIR.py:138-143, 303-331)
    items = [D.relight_grid_for_rank(200, 5, r, 8) for r in range(8)]
    flat = sorted(x for it in items for x in it)
    assert flat == [(e, v) for e in range(5) for v in range(200)]
    assert {len(it) for it in items} == {125}
    with pytest.raises(ValueError):
        D.relight_grid_for_rank(4, 2, 3, 2)


def test_flat_bucket_views_single_process():
    a = torch.randn(5, 3, requires_grad=True)
    b = torch.randn(7, requires_grad=True)
    bk = D.FlatGradBucket([a, b])
    assert bk.numel == 16 + 8 and bk.attached()
    (a.sum() * 2 + (b * b).sum()).backward()
    assert bk.attached(), "autograd must accumulate into the bucket views in place"
    assert torch.allclose(bk.view(0), torch.full((5, 3), 2.0))
    assert torch.allclose(bk.view(1), 2 * b.detach())
    bk.zero()
    assert float(bk.flat.abs().sum()) == 0.0
    a.grad = None  # something dropped the view: all_reduce() re-gathers and re-attaches
    (a.sum() * 3).backward()
    bk.all_reduce()
    assert bk.attached() and torch.allclose(a.grad, torch.full((5, 3), 3.0))


def test_segmented_bucket_issues_segments_in_backward_order():
    """Single process: segment k's reduce is issued as soon as all of its parameters have their gradient,
    i.e. in the order the backward pass finishes them, and the extra slot sits in the last segment."""
    a = torch.randn(6, 2, requires_grad=True)
    b = torch.randn(7, requires_grad=True)
    c = torch.randn(3, 3, requires_grad=True)
    bk = D.FlatGradBucket([a, b, c], segments=[[1], [0, 2]], extra_floats=1)
    assert bk.offsets[1] == 0 and bk.seg_bounds[0] == (0, 8)
    assert bk.seg_bounds[1][1] == bk.numel and bk.extra.numel() == 1
    assert bk.extra.data_ptr() == bk.flat.data_ptr() + 4 * bk.extra_offset
    # forward order a -> c -> b means backward finishes b first, then c, then a
    y = ((a * 2).sum() + (c * 3).sum()) * 1.0
    z = y + (b * 5).sum()
    bk.zero()
    bk.extra[0] = 1.0
    bk.begin_overlap()
    z.backward()
    assert bk.overlap_log == [0, 1]
    bk.finish_overlap()
    assert bk.attached()
    assert torch.allclose(a.grad, torch.full((6, 2), 2.0)) and torch.allclose(b.grad, torch.full((7,), 5.0))
    assert torch.allclose(c.grad, torch.full((3, 3), 3.0)) and float(bk.extra[0]) == 1.0
    # a parameter without a gradient this step: finish_overlap() still issues its segment
    bk.zero()
    bk.begin_overlap()
    (b * 1.0).sum().backward()
    assert bk.overlap_log == [0]
    bk.finish_overlap()
    assert bk.overlap_log == [0, 1]
    with pytest.raises(ValueError):
        D.FlatGradBucket([a, b, c], segments=[[0], [0, 2]])


def test_bucket_with_external_storage_and_custom_reducer():
    """The hooks PeerAllReduce plugs into (alloc = peer-mapped storage, reducer = the one-kernel all-reduce):
    gradients accumulate straight into the supplied storage, the reducer runs exactly once per step, after the
    backward pass, on the whole buffer; segmented buckets cannot take a whole-buffer reducer."""
    a = torch.randn(5, 3, requires_grad=True)
    b = torch.randn(7, requires_grad=True)
    arena = torch.full((64,), 7.0)
    calls = []

    def alloc(n):
        v = arena[:n]
        v.zero_()
        return v

    def reducer(flat):
        calls.append(flat.data_ptr())
        flat.mul_(2.0)   # stands in for "sum over 2 identical ranks"

    bk = D.FlatGradBucket([a, b], extra_floats=1, alloc=alloc, reducer=reducer)
    assert bk.flat.data_ptr() == arena.data_ptr() and bk.attached() and bk.extra is not None
    bk.begin_overlap()               # reducer mode arms no hooks: nothing may fire inside backward
    (a.sum() + (b * b).sum()).backward()
    assert calls == []
    bk.finish_overlap()
    assert calls == [arena.data_ptr()]
    assert torch.allclose(bk.view(0), torch.full((5, 3), 2.0)) and torch.allclose(bk.view(1), 4 * b.detach())
    bk.zero()
    (a.sum()).backward()
    bk.all_reduce()                  # the post-step entry point goes through the same reducer
    assert len(calls) == 2 and torch.allclose(bk.view(0), torch.full((5, 3), 2.0))
    with pytest.raises(ValueError):
        D.FlatGradBucket([a, b], segments=[[0], [1]], reducer=reducer)
    with pytest.raises(ValueError):
        D.FlatGradBucket([a, b], alloc=lambda n: torch.zeros(n + 4))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, dev = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and dev.type == "cpu"
    torch.manual_seed(0)  # replicas: identical parameters on every rank
    base = torch.randn(50, 12, requires_grad=True)
    rough = torch.randn(50, 4, requires_grad=True)
    env = torch.randn(8, 16, 3, requires_grad=True)
    bk = D.FlatGradBucket([base, rough, env])
    targets = [torch.full((50, 12), float(v)) for v in range(5)]

    def render_and_backward(v):  # stands in for pipeline.training_step(view v): per-view loss
        loss = ((base - targets[v]) ** 2).sum() * (v + 1) + (rough * (v + 1)).sum() + (env * env).sum()
        loss.backward()
        return loss

    total, mine = D.data_parallel_step(list(range(5)), render_and_backward, bk)
    assert mine == D.views_for_rank(5, rank, world)
    # the expected gradient is the sum over ALL five views, whichever rank rendered them
    exp_base = sum(2 * (base.detach() - targets[v]) * (v + 1) for v in range(5))
    exp_rough = torch.full((50, 4), float(sum(v + 1 for v in range(5))))
    exp_env = 5 * 2 * env.detach()
    ok = (torch.allclose(base.grad, exp_base, rtol=1e-5, atol=1e-5) and torch.allclose(rough.grad, exp_rough)
          and torch.allclose(env.grad, exp_env, rtol=1e-5, atol=1e-5) and bk.attached())
    # the same exchange, segment-wise from inside the backward pass (what GraphedTrainingStep captures)
    bk2 = D.FlatGradBucket([base, rough, env], segments=[[2, 1], [0]], extra_floats=1)
    bk2.zero()
    bk2.extra[0] = float(rank == 1)      # "rank 1 overflowed its bins"
    bk2.begin_overlap()
    loss = sum(((base - targets[v]) ** 2).sum() * (v + 1) + (rough * (v + 1)).sum() + (env * env).sum()
               for v in D.views_for_rank(5, rank, world))
    loss.backward()
    bk2.finish_overlap()
    ok = ok and (torch.allclose(base.grad, exp_base, rtol=1e-5, atol=1e-5) and torch.allclose(rough.grad, exp_rough)
                 and torch.allclose(env.grad, exp_env, rtol=1e-5, atol=1e-5) and bk2.attached()
                 and float(bk2.extra[0]) == 1.0 and sorted(bk2.overlap_log) == [0, 1])
    t = D.max_over_ranks(float(rank + 1), dev)
    q.put((rank, bool(ok), t, len(mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [2.0, 2.0]          # max over ranks
    assert sorted(r[3] for r in res) == [2, 3]        # 5 views dealt 3 + 2


def test_peer_allreduce_needs_a_process_group_and_has_no_fallback():
    """PeerAllReduce is the peer-memory (GPU box) exchange: outside an initialised multi-rank group it raises instead
    of silently degrading; callers (bench.py) then choose the NCCL path explicitly."""
    assert not dist.is_initialized()
    with pytest.raises(RuntimeError):
        D.PeerAllReduce(torch.device("cpu"))
    assert D.background_group() is None   # NCCL-only helper: no group outside NCCL


def test_graphed_step_accumulate_mode_needs_a_bucket():
    """GraphedTrainingStep(zero_in_graph=False) accumulates several views per step into the bucket: it needs one, and
    the reduction must then happen outside the graph (host-side argument check, no CUDA involved)."""
    from svgir_b200 import pipeline
    a = torch.randn(4, 3, requires_grad=True)
    bk = D.FlatGradBucket([a], extra_floats=1)
    with pytest.raises(ValueError):
        pipeline.GraphedTrainingStep(None, None, None, None, None, bucket=None, zero_in_graph=False)
    with pytest.raises(ValueError):
        pipeline.GraphedTrainingStep(None, None, None, None, None, bucket=bk, reduce_in_graph=True, zero_in_graph=False)
