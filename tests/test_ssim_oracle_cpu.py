"""CPU: the SSIM oracle (oracle/ssim_oracle.py) against goldens produced by the reference's own `ssim`
(utils/loss_utils.py:32-62) and torch autograd (tests/golden/make_golden_ssim.py)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_ssim_oracle_matches_reference_golden():
    from oracle import ssim_oracle as so
    g = np.load(os.path.join(HERE, "golden", "ref_ssim.npz"))
    for tag in ("a", "b", "c"):
        x, y = g[f"{tag}_img1"], g[f"{tag}_img2"]
        assert abs(so.ssim(x, y) - float(g[f"{tag}_ssim"])) < 2e-6
        grad = so.ssim_grad(x, y)
        ref = g[f"{tag}_grad"]
        assert np.abs(grad - ref).max() < 2e-6 * max(1.0, np.abs(ref).max() * 1e3), tag


def test_ssim_oracle_properties():
    from oracle import ssim_oracle as so
    rng = np.random.default_rng(3)
    x = rng.uniform(0, 1, (3, 20, 24)).astype(np.float32)
    assert abs(so.ssim(x, x) - 1.0) < 1e-6            # identical images
    assert np.abs(so.ssim_grad(x, x)).max() < 1e-6     # ... are a stationary point
    w = so.gaussian_window()
    assert w.shape == (11,) and abs(float(w.sum()) - 1.0) < 1e-6 and np.allclose(w, w[::-1])
    # finite-difference check of one gradient entry
    y = rng.uniform(0, 1, x.shape).astype(np.float32)
    gidx = (1, 7, 9)
    e = 1e-3
    xp, xm = x.copy(), x.copy()
    xp[gidx] += e
    xm[gidx] -= e
    fd = (so.ssim(xp, y) - so.ssim(xm, y)) / (2 * e)
    assert abs(fd - so.ssim_grad(x, y)[gidx]) < 1e-5
