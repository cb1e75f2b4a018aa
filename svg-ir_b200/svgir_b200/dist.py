"""View-sharded data parallelism for the hot path.

The reference is single-process / single-GPU (train.py:108-143 renders ONE view per iteration;
grep for nccl|distributed finds nothing), so nothing here is ported: this is the new harness the
north star asks for.  One process per GPU holds a full replica of the surfel parameters; a step
draws V views and rank r renders views r, r+G, ...; the per-surfel parameter gradients are summed
with ONE all-reduce over a flat fp32 buffer (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  Relighting sweeps shard the (view x env-map) grid round-robin with no collective.

PyTorch is used for the process group and device memory only.
"""
from __future__ import annotations

import os
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, torch.device]:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun contract) and joins the process
    group when WORLD_SIZE > 1. Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
    else:
        dev = torch.device("cpu")
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        be = backend or ("nccl" if dev.type == "cuda" else "gloo")
        if be == "nccl":
            dist.init_process_group(be, device_id=dev)
        else:
            dist.init_process_group(be)
    return rank, world, dev


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def rank_id() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def views_for_rank(n_views: int, rank: int, world: int) -> List[int]:
    """Rank r renders views r, r+G, r+2G, ... of the step's view list (SURVEY 8(e) 'Partitioning')."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_views, world))


def relight_grid_for_rank(n_views: int, n_envs: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """The relight sweep (eval_relighting_tensoIR.py:138-143, 303-331: env-map outer loop, test frames
    inner) as a flat (env, view) grid dealt round-robin; no collective is needed. Consecutive items of
    one rank share the env map as long as possible (the grid is env-major), so the per-env-map
    visibility / radiance precompute is redone at most ceil(n_envs) times per rank."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return [(i // n_views, i % n_views) for i in range(rank, n_views * n_envs, world)]


class FlatGradBucket:
    """One contiguous fp32 buffer holding the gradients of every trainable tensor.

    `.grad` of each parameter is a VIEW into the buffer, so autograd accumulates straight into it
    (no torch.cat / copy before the collective) and the optimiser reads the reduced gradient in
    place. `zero()` is one memset; `all_reduce()` is one collective per step.

    Overlap (`segments=[[param indices], ...]`, in the order the backward pass finishes them): the
    buffer is laid out segment by segment and `begin_overlap()` arms post-accumulate hooks; as soon
    as every parameter of a segment has received its gradient, that segment's all-reduce is issued
    asynchronously (NCCL's own stream), so it runs under the rest of the backward pass -- for the
    stage-2 step the rasteriser-side gradients (SH, opacity, scale, rotation: 56 of 87 floats per
    surfel) travel while the shading backward kernel is still running. `finish_overlap()` issues
    whatever is left and makes the current stream wait for all of it. The whole sequence can be
    captured into a CUDA graph (pipeline.GraphedTrainingStep)."""

    def __init__(self, params: Sequence[torch.Tensor], average: bool = False,
                 segments: Optional[Sequence[Sequence[int]]] = None, extra_floats: int = 0):
        params = [p for p in params if p is not None]
        if not params:
            raise ValueError("FlatGradBucket needs at least one parameter")
        dev = params[0].device
        for p in params:
            if p.device != dev or p.dtype != torch.float32:
                raise ValueError("all bucket parameters must be fp32 on one device")
        self.params = list(params)
        self.average = average
        if segments is None:
            segments = [list(range(len(self.params)))]
        seen = sorted(i for seg in segments for i in seg)
        if seen != list(range(len(self.params))):
            raise ValueError("segments must name every parameter exactly once")
        self.segments = [list(seg) for seg in segments if len(seg)]
        self.offsets = [0] * len(self.params)
        self.seg_bounds = []
        n = 0
        for seg in self.segments:
            lo = n
            for i in seg:
                self.offsets[i] = n
                n += (self.params[i].numel() + 3) // 4 * 4  # keep every view 16-byte aligned for vector loads
            self.seg_bounds.append((lo, n))
        # `extra_floats` scalars ride at the tail of the LAST segment (summed over ranks with it): the graphed
        # step puts its binning-overflow flag there so every rank takes the same re-capture decision.
        self.extra_offset = n
        if extra_floats:
            n += (int(extra_floats) + 3) // 4 * 4
            self.seg_bounds[-1] = (self.seg_bounds[-1][0], n)
        self.numel = n
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.extra = self.flat[self.extra_offset:self.extra_offset + int(extra_floats)] if extra_floats else None
        self.attach()
        self._work = None
        self._seg_of = {i: k for k, seg in enumerate(self.segments) for i in seg}
        self._hooks = []
        self._armed = False
        self._pending = []
        self._issued = []
        self._ready = []
        self.overlap_log = []   # segment issue order of the last overlapped step (tests / diagnostics)

    def attach(self):
        """(Re)binds p.grad to the buffer views (call again if something replaced .grad)."""
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o:o + p.numel()].view_as(p)

    def attached(self) -> bool:
        return all(p.grad is not None and p.grad.data_ptr() == self.flat.data_ptr() + 4 * o
                   for p, o in zip(self.params, self.offsets))

    def zero(self):
        self.flat.zero_()
        if not self.attached():
            self.attach()

    def view(self, i: int) -> torch.Tensor:
        p, o = self.params[i], self.offsets[i]
        return self.flat[o:o + p.numel()].view_as(p)

    def segment(self, k: int) -> torch.Tensor:
        lo, hi = self.seg_bounds[k]
        return self.flat[lo:hi]

    def _gather_if_detached(self):
        if not self.attached():  # something (e.g. zero_grad(set_to_none=True)) replaced .grad: gather
            for p, o in zip(self.params, self.offsets):
                if p.grad is not None:
                    self.flat[o:o + p.numel()].copy_(p.grad.reshape(-1))
                else:
                    self.flat[o:o + p.numel()].zero_()
            self.attach()

    def all_reduce(self, async_op: bool = False):
        """Sum (or mean) over ranks. With async_op the NCCL kernel runs on its own stream and overlaps
        whatever the caller enqueues next; call wait() before reading the gradients."""
        self._gather_if_detached()
        if world_size() == 1:
            return None
        if self.average:
            self.flat.div_(world_size())
        self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        return self._work

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None

    # ---- segment-wise all-reduce overlapped with the backward pass -----------------------------
    def _issue(self, k: int):
        if self._issued[k]:
            return
        self._issued[k] = True
        self.overlap_log.append(k)
        if world_size() == 1:
            return
        seg = self.segment(k)
        if self.average:
            seg.div_(world_size())
        self._pending.append(dist.all_reduce(seg, op=dist.ReduceOp.SUM, async_op=True))

    def _on_grad(self, i: int):
        if not self._armed:
            return
        k = self._seg_of[i]
        self._ready[k] += 1
        if self._ready[k] == len(self.segments[k]):
            self._issue(k)

    def begin_overlap(self):
        """Arms the hooks for ONE backward pass in which every parameter receives at most one
        accumulated gradient (one view per step). Gradients must accumulate into the attached views."""
        if not self._hooks:
            for i, p in enumerate(self.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(lambda _p, i=i: self._on_grad(i)))
        if not self.attached():
            self.attach()
        self._ready = [0] * len(self.segments)
        self._issued = [False] * len(self.segments)
        self._pending = []
        self.overlap_log = []
        self._armed = True

    def finish_overlap(self):
        """Issues the segments whose hooks did not all fire (parameters without a gradient this step),
        then makes the current stream wait for every outstanding all-reduce."""
        self._armed = False
        for k in range(len(self.segments)):
            self._issue(k)
        for w in self._pending:
            w.wait()
        self._pending = []

    @property
    def nbytes(self) -> int:
        return self.numel * 4


def data_parallel_step(step_views: Sequence[int], render_and_backward: Callable[[int], torch.Tensor],
                       bucket: FlatGradBucket, rank: Optional[int] = None, world: Optional[int] = None):
    """One multi-view step: zero the bucket, run `render_and_backward(view)` (which must leave its
    gradients accumulated in .grad) for this rank's share of `step_views`, then all-reduce.
    Returns (sum of local losses, list of local views)."""
    rank = rank_id() if rank is None else rank
    world = world_size() if world is None else world
    bucket.zero()
    mine = [step_views[i] for i in views_for_rank(len(step_views), rank, world)]
    total = None
    for v in mine:
        loss = render_and_backward(v)
        total = loss.detach() if total is None else total + loss.detach()
    bucket.all_reduce()
    return total, mine


def max_over_ranks(value_ms: float, device: torch.device) -> float:
    """Timing helper: device-measured milliseconds, max over ranks."""
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
