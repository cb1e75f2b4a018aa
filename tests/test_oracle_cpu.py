"""CPU: pins the C oracle (oracle/svgss_oracle.c) against golden vectors captured from the
UNMODIFIED reference CUDA extension on a B200 (tests/golden/make_golden_gpu.py), plus closed-form
and property checks of the compositing equations."""
import os

import numpy as np
import pytest

import util
from golden.make_golden_gpu import CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("tag", ["train", "eval"])
def test_oracle_matches_reference_cuda_golden(tag):
    g = dict(np.load(os.path.join(GOLD, f"ref_svgss_{tag}.npz")))
    case = util.make_case(**CASES[tag])
    grads = util.pixel_grads(case)
    fw, bw = util.run_oracle(case, grads=grads)
    # integer / index work: bit-exact
    assert fw["num_rendered"] == int(g["num_rendered"])
    assert (fw["radii"] == g["out_radii"]).all()
    assert (fw["keys"] == g["keys"]).all()
    assert (fw["point_list"] == g["point_list"]).all()
    assert (fw["ranges"] == g["ranges"]).all()
    # floating point images: 1e-5 abs; the oracle's expf is glibc's, the reference's CUDA's, so allow a
    # vanishing fraction of alpha-threshold flips
    assert (fw["n_contrib"] != g["n_contrib"]).mean() < 1e-3
    for k, ok in (("color", "color"), ("normal", "normal_img"), ("depth", "depth"), ("opacity", "opacity"),
                  ("feature", "feature"), ("vfeature", "vfeature")):
        bad = np.abs(fw[ok] - g["out_" + k]) > 1e-5 + 1e-5 * np.abs(g["out_" + k])
        assert bad.mean() < 1e-3, (k, bad.mean())
    assert np.abs(fw["final_T"] - g["final_T"]).max() < 1e-5
    np.testing.assert_allclose(fw["weights"], g["out_weights"], rtol=1e-4, atol=1e-5)
    # gradients: 1e-3 relative
    for k in ("dL_dmeans2D", "dL_dopacity", "dL_dcolors", "dL_dfeatures", "dL_dvfeatures", "dL_dmeans3D",
              "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations", "dL_dnormal", "dL_ddepth"):
        ref = g["grad_" + k].reshape(bw[k].shape)
        assert util.rel_l2(bw[k], ref) < 1e-3, (k, util.rel_l2(bw[k], ref))
    ref_conic = g["grad_dL_dconic"].reshape(-1, 4)
    assert util.rel_l2(bw["dL_dconic"], ref_conic) < 1e-3


def _single_surfel_case():
    """One opaque surfel facing the camera at the image centre."""
    from svgir_b200 import scene
    cam = scene.look_at_camera(64, 64, 0, n_views=1, distance=4.0)
    fwd = -cam.campos / np.linalg.norm(cam.campos)
    n = -fwd
    helper = np.array([1.0, 0.0, 0.0]) if abs(n[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
    a = np.cross(helper, n); a /= np.linalg.norm(a)
    b = np.cross(n, a)
    q = scene._rotmat_to_quat(np.stack([a, b, n], 1)[None]).astype(np.float32)
    cl = scene.SurfelCloud(np.zeros((1, 3), np.float32), np.array([[0.05, 0.05, 1e-6]], np.float32), q,
                           np.array([[0.9]], np.float32), np.zeros((1, 16, 3), np.float32), n[None].astype(np.float32))
    return dict(cloud=cl, cam=cam, features=np.array([[2.0]], np.float32),
                vfeatures=np.array([[1.0, 1.0, 1.0, 1.0]], np.float32), colors=np.array([[0.2, 0.4, 0.6]], np.float32),
                S=1, VS=4, bg=np.array([1.0, 0.0, 0.0], np.float32), config=np.array([1, 1, 1], np.float32))


def test_single_surfel_closed_form():
    case = _single_surfel_case()
    fw, _ = util.run_oracle(case, backward=False)
    assert fw["radii"][0] > 0
    # pixel nearest the projected centre: alpha = min(.99, o*exp(power)), colour = a*c + (1-a)*bg
    mx, my = fw["means2D"][0]
    px, py = int(round(mx)), int(round(my))
    con = fw["conic_opacity"][0]
    dx, dy = mx - px, my - py
    power = -0.5 * (con[0] * dx * dx + con[2] * dy * dy) - con[1] * dx * dy
    alpha = min(0.99, 0.9 * np.exp(power))
    np.testing.assert_allclose(fw["opacity"][0, py, px], alpha, rtol=1e-5)
    np.testing.assert_allclose(fw["color"][:, py, px], alpha * np.array([0.2, 0.4, 0.6]) + (1 - alpha) * np.array([1.0, 0, 0]), rtol=1e-5)
    # the four bilinear weights sum to one, so a constant vfeature row composites like a flat feature
    np.testing.assert_allclose(fw["vfeature"][0, py, px], alpha * 1.0, rtol=1e-5)
    np.testing.assert_allclose(fw["feature"][0, py, px], alpha * 2.0, rtol=1e-5)
    # normalised depth of a single surfel is its own per-pixel depth, close to the centre depth
    np.testing.assert_allclose(fw["depth"][0, py, px], fw["depths"][0], rtol=1e-2)
    # far corner: untouched -> background and T clamped to 1-1e-6 (forward.cu:671)
    np.testing.assert_allclose(fw["color"][:, 0, 0], [0.999999, 0, 0], atol=1e-6)
    assert fw["n_contrib"].reshape(64, 64)[0, 0] == 0


def test_empty_scene_and_binning_properties():
    from oracle import svgss as O
    case = util.make_case(3000, 130, 70, seed=77)
    fw, _ = util.run_oracle(case, backward=False)
    R = fw["num_rendered"]
    keys = fw["keys"]
    assert (np.diff(keys.astype(np.uint64)) >= 0).all()                      # sortedness
    assert sorted(fw["keys_unsorted"].tolist()) == keys.tolist()              # permutation
    assert R == int(fw["tiles_touched"].astype(np.int64).sum())
    rg = fw["ranges"].astype(np.int64)
    nz = rg[:, 1] > rg[:, 0]
    assert int((rg[nz, 1] - rg[nz, 0]).sum()) == R                            # ranges partition the list
    assert (rg[~nz] == 0).all()                                               # empty tiles are (0,0)
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    for t in np.nonzero(nz)[0][:50]:
        assert (tiles[rg[t, 0]:rg[t, 1]] == t).all()
    # stable order: equal keys keep ascending surfel index
    same = keys[1:] == keys[:-1]
    assert (fw["point_list"][1:][same] > fw["point_list"][:-1][same]).all()


def test_backward_is_linear_in_pixel_gradients():
    case = util.make_case(1200, 64, 48, seed=13)
    g1 = util.pixel_grads(case, seed=1)
    g2 = util.pixel_grads(case, seed=2)
    gs = {k: g1[k] + 2.0 * g2[k] for k in g1}
    _, b1 = util.run_oracle(case, grads=g1)
    _, b2 = util.run_oracle(case, grads=g2)
    _, bs = util.run_oracle(case, grads=gs)
    for k in ("dL_dmeans3D", "dL_dvfeatures", "dL_dopacity", "dL_dsh", "dL_drotations"):
        assert util.rel_l2(bs[k], b1[k] + 2.0 * b2[k]) < 1e-4, k
