"""Ad-hoc: where does the CUDA LBVH differ from the oracle?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import numpy as np, torch, util
from svgir_b200 import bvh
from oracle import bvh as OB
from test_bvh_oracle_cpu import torch_leaf_boxes
for P in (2, 3, 1000):
    c = util.make_bvh_case(P, 4, seed=100 + P)
    t = {k: torch.from_numpy(v).cuda() for k, v in c.items()}
    nodes, aabbs = bvh.leaf_aabbs(t["means"], t["scales"], t["rotations"])
    a_before = aabbs.cpu().numpy().copy()
    n0, a0 = OB.init(c["means"], c["scales"], c["rotations"])
    print(P, "leaf boxes: gpu vs oracle mismatches", int((a_before != a0).sum()), "max abs", float(np.abs(a_before - a0).max()))
    tn, ta = torch_leaf_boxes(*(torch.from_numpy(c[k]) for k in ("means", "scales", "rotations")))
    print("   torch-cpu vs torch-gpu leaf mismatches", int((ta.numpy() != a_before).sum()))
    tree = bvh.Bvh(nodes, aabbs)
    torch.cuda.synchronize()
    on, oa, om = OB.create(c["means"], c["scales"], c["rotations"])
    a = tree.aabbs.cpu().numpy()
    bad = np.argwhere(a != oa)
    print("   after build: node mism", int((tree.nodes.cpu().numpy() != on).sum()), "aabb mism", len(bad), bad[:6].tolist())
    if len(bad):
        i = bad[0][0]; print("   row", i, a[i], oa[i])
    print("   morton mism", int((tree.morton.cpu().numpy().astype(np.uint64) != om).sum()))
