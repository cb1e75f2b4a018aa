"""Parity at BASELINE.json's FULL sizes (SURVEY.md 8(d) configs C2..C5 shapes).

Two kinds of checks per shape:
 * side by side with the UNMODIFIED reference rasteriser (oracle/_ref/libsvgss_ref.so, built from the
   sources under /root/reference by oracle/Makefile and shipped to the GPU box as a binary): sort keys,
   point list, tile ranges, radii, n_contrib bit-exact; images / G-buffers <= 1e-5 abs; gradients
   <= 1e-3 relative (north_star tolerances);
 * size-independent properties that need no reference: keys sorted, ranges tile the key list exactly,
   every key's tile id equals the range that holds it, opacity image == 1 - prod(1-alpha) bound,
   sum(weights) == sum(opacity image) (both are sum_i alpha_i T_i), backward linear in pixel gradients.
"""
import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu

GRADS = ("dL_dmeans2D", "dL_dopacity", "dL_dcolors", "dL_dfeatures", "dL_dvfeatures", "dL_dmeans3D", "dL_dcov3D",
         "dL_dsh", "dL_dscales", "dL_drotations")

# (tag, P, W, H, S, VS, backward)
SHAPES = [
    ("C3-train", 300_000, 800, 800, 4, 52, True),
    ("C3-eval", 300_000, 800, 800, 7, 64, False),
    ("C4-view", 1_000_000, 800, 800, 4, 52, True),
    ("C5-view", 2_000_000, 1920, 1080, 7, 64, False),
]


def _tiles(W, H):
    return ((W + 15) // 16), ((H + 15) // 16)


def _check_binning_properties(st, W, H):
    R = st.num_rendered
    tx, ty = _tiles(W, H)
    keys = st.t["sorted_keys"][:R]
    assert R > 0
    assert bool((keys[1:] >= keys[:-1]).all()), "keys not sorted"
    tile = (keys >> 32).to(torch.int64)
    assert int(tile.min()) >= 0 and int(tile.max()) < tx * ty
    rng = st.t["ranges"].to(torch.int64)
    cnt = torch.bincount(tile, minlength=tx * ty)
    nonempty = cnt > 0
    assert bool(((rng[:, 1] - rng[:, 0]) == cnt)[nonempty].all()), "range length != tile population"
    assert bool((rng[~nonempty] == 0).all()), "untouched tiles must stay (0,0) (rasterizer_impl.cu:340)"
    start = torch.cumsum(cnt, 0) - cnt
    assert bool((rng[:, 0] == start)[nonempty].all()), "ranges do not tile the key list"
    # point_list indexes visible surfels only
    pl = st.t["point_list"][:R].to(torch.int64)
    assert int(pl.min()) >= 0


@pytest.mark.parametrize("tag,P,W,H,S,VS,backward", SHAPES, ids=[s[0] for s in SHAPES])
def test_full_size_vs_reference_and_properties(tag, P, W, H, S, VS, backward):
    from oracle import ref_cuda
    case = util.make_case(P, W, H, S=S, VS=VS, seed=1234)
    g = util.pixel_grads(case) if backward else None
    out, st, bw = util.run_ours(case, backward=backward, grads=g)
    _check_binning_properties(st, W, H)
    # sum_i w_i over all (pixel, surfel) pairs, computed two ways
    wsum = float(out["weights"].double().sum())
    osum = float(out["opacity"].double().sum())
    assert abs(wsum - osum) <= 2e-4 * max(osum, 1.0), (wsum, osum)
    assert float(out["opacity"].max()) <= 1.0 and float(out["opacity"].min()) >= 0.0
    vis = out["radii"] > 0
    assert bool((out["weights"].reshape(-1)[~vis] == 0).all())

    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libsvgss_ref.so not in this snapshot (properties checked)")
    r, rout, rbw = util.run_ref(case, backward=backward, grads=g)
    R = rout["num_rendered"]
    tx, ty = _tiles(W, H)
    assert R == st.num_rendered
    assert bool((out["radii"] == rout["radii"]).all())
    assert bool((r.state("keys", (R,), torch.int64) == st.t["sorted_keys"][:R]).all())
    assert bool((r.state("point_list", (R,), torch.int32) == st.t["point_list"][:R]).all())
    assert bool((r.state("ranges", (tx * ty, 2), torch.int32) == st.t["ranges"]).all())
    assert bool((r.state("n_contrib", (W * H,), torch.int32) == st.t["n_contrib"]).all())
    for k in ("color", "normal", "depth", "opacity", "feature", "vfeature"):
        assert float((out[k] - rout[k]).abs().max()) <= 1e-5, (tag, k)
    assert util.rel_l2(out["weights"].cpu().numpy(), rout["weights"].cpu().numpy().reshape(-1, 1)) < 1e-5
    if backward:
        for k in GRADS:
            a = bw[k].cpu().numpy()
            b = rbw[k].cpu().numpy().reshape(a.shape)
            assert util.rel_l2(a, b) < 1e-3, (tag, k, util.rel_l2(a, b))


def test_full_size_backward_is_linear_in_pixel_gradients():
    """dL/dtheta(a*g1 + b*g2) == a*dL/dtheta(g1) + b*dL/dtheta(g2) at C3-train size: every backward kernel is
    linear in the upstream image gradients (backward.cu:530-934 has no g-dependent branch)."""
    from svgir_b200 import raster
    case = util.make_case(300_000, 800, 800, S=4, VS=52, seed=1234)
    g1, g2 = util.pixel_grads(case, seed=7), util.pixel_grads(case, seed=8)
    a, b = 0.75, -1.5
    g3 = {k: (a * g1[k] + b * g2[k]).astype(np.float32) for k in g1}
    out, st, bw1 = util.run_ours(case, grads=g1)
    res = []
    for g in (g2, g3):
        gt = {k: torch.from_numpy(v).cuda() for k, v in g.items()}
        res.append(raster.backward(st, out["radii"], gt, want_debug=True))
    torch.cuda.synchronize()
    bw2, bw3 = res
    for k in GRADS:
        lin = a * bw1[k].double() + b * bw2[k].double()
        err = float((bw3[k].double() - lin).norm() / lin.norm().clamp_min(1e-30))
        assert err < 1e-3, (k, err)
