"""CPU: the C restatement of the reference LBVH (oracle/bvh_oracle.c) -- structural invariants of the
build, the torch evaluation of RayTracer.__init__ it must equal bit for bit, tree trace == flat trace,
and the golden vectors captured from the reference BVH kernels on a B200."""
import os

import numpy as np
import pytest
import torch

import util
from oracle import bvh as OB

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def torch_leaf_boxes(means3D, scales, rotations):
    """RayTracer.__init__ (submodules/bvh/__init__.py:29-57) + build_rotation
    (utils/general_utils.py:82-103) with the reference's torch expressions, device-agnostic."""
    r = rotations
    sumsq = r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3]
    # the reference runs on CUDA, where torch.sqrt is the correctly rounded sqrt.rn; torch's vectorised
    # CPU sqrt is not (it differs from IEEE in ~1 % of the inputs), so the CPU evaluation uses numpy's
    norm = torch.sqrt(sumsq) if sumsq.is_cuda else torch.from_numpy(np.sqrt(sumsq.numpy()))
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - r * z)
    R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y)
    R[:, 2, 1] = 2 * (y * z + r * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    P = means3D.shape[0]
    nodes = torch.full((2 * P - 1, 5), -1, device=r.device).int()
    nodes[:P - 1, 4] = 0
    nodes[P - 1:, 4] = 1
    aabbs = torch.zeros(2 * P - 1, 6, device=r.device).float()
    aabbs[:, :3] = 100000
    aabbs[:, 3:] = -100000
    a, b, c = R[:, :, 0], R[:, :, 1], R[:, :, 2]
    m = 3
    sa, sb, sc = m * scales[:, 0], m * scales[:, 1], m * scales[:, 2]
    corners = []
    for s1 in (1, -1):
        for s2 in (1, -1):
            for s3 in (1, -1):
                v = means3D + a * sa[:, None] if s1 > 0 else means3D - a * sa[:, None]
                v = v + b * sb[:, None] if s2 > 0 else v - b * sb[:, None]
                v = v + c * sc[:, None] if s3 > 0 else v - c * sc[:, None]
                corners.append(v)
    st = torch.stack(corners)
    aabbs[P - 1:] = torch.cat([st.min(0).values, st.max(0).values], dim=-1)
    return nodes, aabbs


def assert_boxes_equal_up_to_reference_race(ours, ref, P):
    """The reference's bottom-up merge (construct.cu:232-263) has no __threadfence between a thread's box
    store and its atomicCAS on the parent's flag, so the second arrival occasionally reads a sibling box
    that is still (partly) the +-100000 initialisation and stores an incomplete merge. Leaf rows must be
    bit-equal; an internal row may differ only by the reference's box being a strict subset of the true
    merge, and only for a vanishing fraction of the nodes."""
    ours, ref = np.asarray(ours), np.asarray(ref)
    assert (ours[P - 1:] == ref[P - 1:]).all()
    bad = np.unique(np.argwhere(ours != ref)[:, 0])
    assert len(bad) <= max(2, 2e-3 * P), len(bad)
    for i in bad:
        assert (ours[i, :3] <= ref[i, :3]).all() and (ours[i, 3:] >= ref[i, 3:]).all()
    return len(bad)


def check_tree(nodes, aabbs, morton, P):
    assert nodes.shape == (2 * P - 1, 5)
    assert nodes[0, 0] == -1 and nodes[0, 4] == P
    code = morton.astype(np.uint64)
    assert (code[1:] > code[:-1]).all(), "64-bit codes must be strictly ascending"
    objs = nodes[P - 1:, 3]
    assert sorted(objs.tolist()) == list(range(P)), "every surfel is exactly one leaf"
    assert ((code & np.uint64((1 << 31) - 1)).astype(np.int64) == objs).all()
    assert (nodes[P - 1:, 4] == 1).all() and (nodes[:P - 1, 3] == -1).all()
    l, r = nodes[:P - 1, 1], nodes[:P - 1, 2]
    assert (nodes[l, 0] == np.arange(P - 1)).all() and (nodes[r, 0] == np.arange(P - 1)).all()
    assert (nodes[:P - 1, 4] == nodes[l, 4] + nodes[r, 4]).all()
    assert (aabbs[:P - 1, :3] == np.minimum(aabbs[l, :3], aabbs[r, :3])).all()
    assert (aabbs[:P - 1, 3:] == np.maximum(aabbs[l, 3:], aabbs[r, 3:])).all()
    children = np.concatenate([l, r])
    assert sorted(children.tolist()) == list(range(1, 2 * P - 1)), "every non-root node has one parent"


@pytest.mark.parametrize("P,surfel", [(2, True), (3, True), (777, True), (5000, False)])
def test_build_invariants_and_leaf_boxes(P, surfel):
    c = util.make_bvh_case(P, 4, seed=P, surfel=surfel)
    nodes0, aabbs0 = OB.init(c["means"], c["scales"], c["rotations"])
    tn, ta = torch_leaf_boxes(*(torch.from_numpy(c[k]) for k in ("means", "scales", "rotations")))
    assert (nodes0 == tn.numpy()).all()
    assert (aabbs0 == ta.numpy()).all(), "leaf boxes must equal the torch evaluation bit for bit"
    nodes, aabbs, morton = OB.build(nodes0, aabbs0)
    check_tree(nodes, aabbs, morton, P)
    # leaf boxes are a permutation of the input boxes
    assert (aabbs[P - 1:] == aabbs0[P - 1:][nodes[P - 1:, 3]]).all()


def test_duplicate_centres_share_codes_but_build():
    c = util.make_bvh_case(64, 4, seed=3)
    c["means"][:] = c["means"][0]          # all Morton codes equal: the index tie-break must carry the build
    c["scales"][:] = c["scales"][0]
    c["rotations"][:] = c["rotations"][0]
    nodes, aabbs, morton = OB.create(c["means"], c["scales"], c["rotations"])
    check_tree(nodes, aabbs, morton, 64)
    assert len(set((morton >> np.uint64(31)).tolist())) == 1


def test_single_surfel_tree_and_trace():
    c = util.make_bvh_case(1, 16, seed=5)
    nodes, aabbs, morton = OB.create(c["means"], c["scales"], c["rotations"])
    assert nodes.tolist() == [[-1, -1, -1, 0, 1]]
    ci = OB.inverse_covariance(c["scales"], c["rotations"])
    cnt, vis = OB.trace_opacity(nodes, aabbs, c["rays_o"], c["rays_d"], c["means"], ci, c["opacity"], c["normals"])
    assert cnt.shape == (16,) and ((vis == 0) | (vis >= 0.9)).all()


@pytest.mark.parametrize("surfel", [True, False])
def test_tree_trace_equals_flat_trace(surfel):
    c = util.make_bvh_case(3000, 3000, seed=21, surfel=surfel)
    nodes, aabbs, morton = OB.create(c["means"], c["scales"], c["rotations"])
    ci = OB.inverse_covariance(c["scales"], c["rotations"])
    args = (nodes, aabbs, c["rays_o"], c["rays_d"], c["means"], ci, c["opacity"], c["normals"])
    cnt, vis = OB.trace_opacity(*args)
    cnt_b, vis_b = OB.trace_opacity(*args, brute=True)
    assert ((vis == 0) | (vis >= 0.9)).all()
    assert (vis == 0).mean() > 0.02 and (cnt > 0).mean() > 0.05, "the case must exercise occlusion"
    same = (vis == 0) == (vis_b == 0)
    assert same.mean() > 0.999                       # 0.9-threshold ties only
    assert (cnt[same] == cnt_b[same]).all()
    np.testing.assert_allclose(vis[same], vis_b[same], rtol=2e-6, atol=0)


def test_oracle_matches_reference_bvh_golden():
    path = os.path.join(GOLD, "ref_bvh_small.npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet")
    g = dict(np.load(path))
    from golden.make_golden_gpu import BVH_CASE
    c = util.make_bvh_case(**BVH_CASE)
    nodes, aabbs, morton = OB.create(c["means"], c["scales"], c["rotations"])
    assert (nodes == g["nodes"]).all() and (morton == g["morton"]).all()
    assert_boxes_equal_up_to_reference_race(aabbs, g["aabbs"], BVH_CASE["P"])
    ci = OB.inverse_covariance(c["scales"], c["rotations"])
    assert util.rel_l2(ci, g["cov_inv"]) < 1e-5
    cnt, vis = OB.trace_opacity(nodes, aabbs, c["rays_o"], c["rays_d"], c["means"], g["cov_inv"], c["opacity"], c["normals"])
    same = (vis == 0) == (g["visibility"] == 0)
    assert same.mean() > 0.995
    assert (cnt[same] != g["contributes"][same]).mean() < 5e-3
    ok = same & (cnt == g["contributes"])
    # surfels: Sigma^-1 ~ 1/s_z^2 = 1e8 along the normal makes `power` ill-conditioned (e_n ~ 1 ulp of the
    # position, squared, times 1e8): the reference build's FMA contraction vs this file's separately rounded
    # arithmetic moves alpha by up to ~1e-3. The CUDA kernel is held to 1e-5 against the reference kernels
    # themselves (tests/test_bvh_gpu.py); the CPU restatement can only be held to the conditioning.
    assert np.abs(vis[ok] - g["visibility"][ok]).max() < 5e-3
    assert np.abs(vis[ok] - g["visibility"][ok]).mean() < 1e-4
