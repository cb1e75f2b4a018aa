"""View-sharded data parallelism for the hot path.

The reference is single-process / single-GPU (train.py:108-143 renders ONE view per iteration;
grep for nccl|distributed finds nothing), so nothing here is ported: this is the new harness the
north star asks for.  One process per GPU holds a full replica of the surfel parameters; a step
draws V views and rank r renders views r, r+G, ...; the per-surfel parameter gradients are summed
once per step over a flat fp32 buffer -- on the GPU box by the svgir peer-memory kernels
(`PeerAllReduce`: the buffer is peer-mapped, NVSwitch multicast or peer loads/stores over NVLink, the
rasteriser-side segment overlapped with the shading backward) or by NCCL, in the CPU tests by gloo.
Relighting sweeps shard the (view x env-map) grid round-robin with no collective.

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


def background_group(max_ctas: int = 4):
    """A second NCCL communicator over all ranks whose kernels use at most `max_ctas` CTAs: collectives issued on it
    run concurrently with a compute kernel at the cost of ~max_ctas of the 148 SMs (the default communicator's
    16-32 CTAs slowed the shading backward from 0.48 to 0.80 ms at N=8). Returns None outside NCCL."""
    if not dist.is_initialized() or dist.get_backend() != "nccl":
        return None
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = int(max_ctas)
    opts.config.min_ctas = 1
    return dist.new_group(backend="nccl", pg_options=opts)


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


class PeerAllReduce:
    """Flat fp32 gradient storage in a symmetric (peer-mapped) allocation + the one-kernel NVLink all-reduce over it
    (csrc/peer_allreduce.cu, `svgir_peer_allreduce`). The backward kernels write their gradients straight into this
    buffer; `all_reduce()` launches ONE kernel on the current stream that sums the buffers of all ranks in place
    (NVSwitch multicast ld_reduce / st when the allocation has a multicast mapping, 128-bit peer loads and stores
    otherwise; multicast is chosen from 4 ranks up). The launch is capturable: pipeline.GraphedTrainingStep records it at the end of the step's graph.

    torch.distributed._symmetric_memory is used for what PyTorch is here for: allocating device memory and exchanging
    the peer mappings at start-up. Raises if the box cannot provide peer-mapped memory; callers fall back to NCCL.
    SVGIR_PEER_MULTICAST=0 forces the peer load/store path."""

    def __init__(self, device: torch.device, group=None):
        if not dist.is_initialized() or dist.get_world_size() < 2:
            raise RuntimeError("PeerAllReduce needs an initialised process group with >= 2 ranks")
        from . import _lib
        if dist.get_world_size() > _lib.MAX_PEERS:
            raise RuntimeError("PeerAllReduce supports up to %d ranks of one box" % _lib.MAX_PEERS)
        self.device = device
        self.group = group if group is not None else dist.group.WORLD
        self.storage = None
        self.flat = None
        self.comm = None
        self.numel = 0
        self.multicast = False

    def allocate(self, numel: int) -> torch.Tensor:
        """Returns the zeroed flat buffer of `numel` floats (a view into the symmetric allocation)."""
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        n = (int(numel) + 3) // 4 * 4
        self.storage = symm_mem.empty(n + _lib.PEER_FLAG_WORDS, dtype=torch.float32, device=self.device)
        self.storage.zero_()
        hdl = symm_mem.rendezvous(self.storage, self.group)
        self.hdl = hdl
        world, rank = int(hdl.world_size), int(hdl.rank)
        comm = _lib.PeerComm()
        comm.world, comm.rank = world, rank
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        for i in range(world):
            comm.bufs[i] = ptrs[i]
            comm.flags[i] = ptrs[i] + 4 * n
        # multicast moves (1 + 1/N) x the buffer per NVLink direction, peer loads/stores 2 (N-1)/N x: multicast from 4 ranks
        # up (measured: N=2 0.32 vs 0.19 ms, N=8 0.26 vs 0.32 ms). SVGIR_PEER_MULTICAST=1 / 0 forces the choice.
        mc = 0
        want = os.environ.get("SVGIR_PEER_MULTICAST", "auto")
        if want == "1" or (want != "0" and world >= 4):
            try:
                mc = int(hdl.multicast_ptr or 0) if hdl.has_multicast_support else 0
            except Exception:
                mc = 0
        comm.multicast = mc or None
        self.multicast = bool(mc)
        self.comm, self.numel = comm, n
        self.flat = self.storage[:n]
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)   # every rank's flags are zeroed before the first launch
        torch.cuda.synchronize(self.device)
        return self.flat[:int(numel)]

    # ---- segment-wise exchange overlapped with the backward pass --------------------------------------------
    BG_GRID = int(os.environ.get("SVGIR_PEER_BG_GRID", "8"))   # CTAs (= SMs) of an overlapped segment's kernel

    def begin_segments(self):
        """Before the backward pass: the shading kernels launched from now on leave BG_GRID SMs to the overlapped
        all-reduce (their persistent grid would otherwise run its displaced CTAs as a second wave)."""
        from . import _lib
        self._joins = []
        _lib.lib().svgir_shade_reserve_sms(self.BG_GRID)

    def reduce_segment(self, k: int, lo: int, hi: int, last: bool):
        """Sums floats [lo, hi) of the buffer over ranks with flag bank k. Every segment but the last runs on a side
        stream (forked from / joined to the current one with events, so it is capturable) on BG_GRID CTAs, beside
        whatever the backward pass launches next; the last one runs on the current stream at full width."""
        from . import _lib
        L = _lib.lib()
        if k >= _lib.PEER_BANKS:
            raise ValueError("at most %d segments" % _lib.PEER_BANKS)
        lo4, hi4 = lo // 4 * 4, (hi + 3) // 4 * 4
        cur = torch.cuda.current_stream(self.device)
        if last:
            L.svgir_shade_reserve_sms(0)
            _lib.check(L.svgir_peer_allreduce_range(self.comm, lo4, hi4 - lo4, k, 0, cur.cuda_stream), "peer_allreduce")
            return
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(self.device)
        fork = torch.cuda.Event()
        fork.record(cur)
        self._side.wait_event(fork)
        _lib.check(L.svgir_peer_allreduce_range(self.comm, lo4, hi4 - lo4, k, self.BG_GRID, self._side.cuda_stream),
                   "peer_allreduce")
        join = torch.cuda.Event()
        join.record(self._side)
        self._joins.append(join)

    def end_segments(self):
        """After the last segment: the current stream waits for the overlapped ones."""
        from . import _lib
        _lib.lib().svgir_shade_reserve_sms(0)
        cur = torch.cuda.current_stream(self.device)
        for j in getattr(self, "_joins", []):
            cur.wait_event(j)
        self._joins = []

    def all_reduce(self, tensor: Optional[torch.Tensor] = None):
        """In-place sum over ranks of the whole flat buffer (the `tensor` argument, if given, must be a view of it)."""
        from . import _lib
        if tensor is not None and (tensor.data_ptr() < self.flat.data_ptr() or
                                   tensor.data_ptr() + 4 * tensor.numel() > self.flat.data_ptr() + 4 * self.numel):
            raise ValueError("PeerAllReduce.all_reduce: tensor is not a view of the symmetric buffer")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().svgir_peer_allreduce(self.comm, self.numel, stream), "peer_allreduce")


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
                 segments: Optional[Sequence[Sequence[int]]] = None, extra_floats: int = 0,
                 segment_groups: Optional[Sequence] = None, alloc: Optional[Callable[[int], torch.Tensor]] = None,
                 reducer: Optional[Callable[[torch.Tensor], None]] = None, segment_peer: Optional["PeerAllReduce"] = None):
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
        # optional process group per segment (None = default group): a segment whose all-reduce runs UNDER a compute
        # kernel goes over a communicator restricted to a few CTAs (`background_group`), so the collective does not
        # take SMs away from that kernel; the last, exposed segment uses the full-width default communicator
        self.segment_groups = list(segment_groups) if segment_groups is not None else [None] * len(self.segments)
        if len(self.segment_groups) != len(self.segments):
            raise ValueError("segment_groups must have one entry per segment")
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
        # `alloc(n)` supplies the flat storage (PeerAllReduce.allocate: peer-mapped memory) and `reducer(flat)` the
        # in-place sum over ranks (PeerAllReduce.all_reduce: one kernel on the current stream, no NCCL); with a
        # reducer the bucket must be a single segment, reduced after the backward pass
        # `segment_peer`: segment-wise exchange through PeerAllReduce.reduce_segment -- every segment but the last is
        # summed on a side stream while the backward pass continues (hooks fire as in the NCCL overlap mode)
        self.segment_peer = segment_peer
        if segment_peer is not None and reducer is not None:
            raise ValueError("use either reducer (whole buffer) or segment_peer (per segment)")
        self.reducer = reducer
        if reducer is not None and len(self.segments) != 1:
            raise ValueError("a custom reducer works on the whole bucket: use one segment")
        self.flat = alloc(n) if alloc is not None else torch.zeros(n, dtype=torch.float32, device=dev)
        if self.flat.numel() != n or self.flat.dtype != torch.float32 or self.flat.device != dev:
            raise ValueError("alloc must return %d fp32 elements on %s" % (n, dev))
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
        if self.reducer is not None:
            if self.average:
                self.flat.div_(world_size())
            self.reducer(self.flat)   # stream-ordered kernel: nothing to wait for
            return None
        if self.segment_peer is not None:
            if self.average:
                self.flat.div_(world_size())
            self.segment_peer.all_reduce(self.flat)
            return None
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
        if self.reducer is not None:   # whole-buffer reducer (it knows its own world)
            if self.average:
                self.flat.div_(world_size())
            self.reducer(self.flat)
            return
        if self.segment_peer is not None:
            lo, hi = self.seg_bounds[k]
            if self.average:
                self.flat[lo:hi].div_(world_size())
            self.segment_peer.reduce_segment(k, lo, hi, last=(k == len(self.segments) - 1))
            return
        if world_size() == 1:
            return
        seg = self.segment(k)
        if self.average:
            seg.div_(world_size())
        self._pending.append(dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.segment_groups[k], async_op=True))

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
        if not self._hooks and self.reducer is None:   # a custom reducer runs once, from finish_overlap()
            for i, p in enumerate(self.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(lambda _p, i=i: self._on_grad(i)))
        if not self.attached():
            self.attach()
        self._ready = [0] * len(self.segments)
        self._issued = [False] * len(self.segments)
        self._pending = []
        self.overlap_log = []
        self._armed = True
        if self.segment_peer is not None:
            self.segment_peer.begin_segments()

    def finish_overlap(self):
        """Issues the segments whose hooks did not all fire (parameters without a gradient this step),
        then makes the current stream wait for every outstanding all-reduce."""
        self._armed = False
        for k in range(len(self.segments)):
            self._issue(k)
        for w in self._pending:
            w.wait()
        self._pending = []
        if self.segment_peer is not None:
            self.segment_peer.end_segments()

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
