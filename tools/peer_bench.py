#!/usr/bin/env python
"""Back-to-back timing of the gradient all-reduce alone (torchrun, one rank per GPU): svgir_peer_allreduce with and
without NVSwitch multicast vs NCCL, on a buffer of the C3-train bucket size (104 MB). Prints ms per call (max over ranks)."""
import os, sys, json
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "svg-ir_b200"))
from svgir_b200 import dist as D

rank, world, dev = D.init_from_env()
n = int(os.environ.get("PEER_BENCH_NUMEL", 26106148))
res = {}
def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 4)
x = torch.randn(n, device=dev)
res["nccl"] = timed(lambda: dist.all_reduce(x))
for mc in ("1", "0"):
    os.environ["SVGIR_PEER_MULTICAST"] = mc
    peer = D.PeerAllReduce(dev)
    flat = peer.allocate(n)
    flat.normal_()
    for weak in ("1", "0"):
        for grid in ("128", "64", "32"):
            os.environ["SVGIR_PEER_WEAK"], os.environ["SVGIR_PEER_GRID"] = weak, grid
            res[("peer_mc" if peer.multicast else "peer_ldst") + ("_weak" if weak == "1" else "_sys") + "_g" + grid] = timed(peer.all_reduce)
    os.environ.pop("SVGIR_PEER_WEAK"); os.environ.pop("SVGIR_PEER_GRID")
if rank == 0:
    print(json.dumps({"world": world, "bytes": n * 4, "ms_per_call": res,
                      "algbw_GBs": {k: round(n * 4 / v / 1e6, 1) for k, v in res.items()}}), flush=True)
torch.cuda.synchronize(); dist.barrier()
os._exit(0)
