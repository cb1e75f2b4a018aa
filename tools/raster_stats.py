"""Work statistics of the compositing kernels on the bench scene (CPU, via the test oracle): how many
(warp, instance) evaluations hit, hits per hit-warp, and what a per-warp footprint cull would remove."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "svg-ir_b200")]
from svgir_b200 import scene
from oracle import svgss as O

P, W, H = int(sys.argv[1]) if len(sys.argv) > 1 else 300000, 800, 800
cl = scene.make_surfels(P, seed=1234)
cam = scene.look_at_camera(W, H, 0, 8)
fw = O.forward(cam, cl.means3D, cl.opacity, cl.scales, cl.rotations, None, None, shs=cl.shs)
R = fw["num_rendered"]; print("R", R, "P_vis", int((fw["radii"] > 0).sum()))
ranges, pl = fw["ranges"], fw["point_list"]
m2, co = fw["means2D"], fw["conic_opacity"]
ncon = fw["n_contrib"].reshape(H, W)
gx = (W + 15) // 16
tot_wi = hit_wi = hits = cull_keep = fwd_wi = 0
nh_hist = np.zeros(33, np.int64)
rng = np.random.default_rng(0)
tiles = rng.choice(len(ranges), 400, replace=False)
for t in tiles:
    a, b = ranges[t]
    if b <= a: continue
    ids = pl[a:b]
    tx, ty = (t % gx) * 16, (t // gx) * 16
    px = (tx + np.arange(16))[None, :].repeat(16, 0).reshape(-1).astype(np.float32)
    py = (ty + np.arange(16))[:, None].repeat(16, 1).reshape(-1).astype(np.float32)
    nc = ncon[ty:ty + 16, tx:tx + 16].reshape(-1)
    tmax = int(nc.max())
    ids = ids[:tmax]
    dx = m2[ids, 0][:, None] - px[None]; dy = m2[ids, 1][:, None] - py[None]
    A, B, Cc, o = co[ids, 0][:, None], co[ids, 1][:, None], co[ids, 2][:, None], co[ids, 3][:, None]
    power = -0.5 * (A * dx * dx + Cc * dy * dy) - B * dx * dy
    alpha = np.minimum(0.99, o * np.exp(power))
    hit = (power <= 0) & (alpha >= 1 / 255) & (np.arange(len(ids))[:, None] < nc[None])
    if os.environ.get("WARP_8x4"):
        hw = hit.reshape(len(ids), 4, 4, 2, 8).transpose(0, 1, 3, 2, 4).reshape(len(ids), 8, 32)  # warp = 8 wide x 4 tall
    else:
        hw = hit.reshape(len(ids), 8, 32)
    nh = hw.sum(-1)
    tot_wi += nh.size; hit_wi += int((nh > 0).sum()); hits += int(nh.sum())
    nh_hist += np.bincount(nh.reshape(-1), minlength=33)
    # per-warp footprint cull: alpha >= 1/255 <=> quadratic form <= 2 ln(255 o) =: L ; y-extent of that ellipse
    L = 2 * np.log(np.maximum(255 * co[ids, 3], 1e-9))
    det = co[ids, 0] * co[ids, 2] - co[ids, 1] ** 2
    ey = np.sqrt(np.maximum(L, 0) * co[ids, 0] / det); ex = np.sqrt(np.maximum(L, 0) * co[ids, 2] / det)
    y0, y1 = m2[ids, 1] - ey, m2[ids, 1] + ey
    x0, x1 = m2[ids, 0] - ex, m2[ids, 0] + ex
    wy0 = ty + 2 * np.arange(8)[None]; wy1 = wy0 + 1
    keep = (L[:, None] > 0) & (y1[:, None] >= wy0) & (y0[:, None] <= wy1) & (x1[:, None] >= tx) & (x0[:, None] <= tx + 15)
    cull_keep += int(keep.sum())
    if not os.environ.get('WARP_8x4'): assert not ((nh > 0) & ~keep).any()
print("warp-instances evaluated (bwd, to tile max n_contrib):", tot_wi, "per tile", tot_wi / len(tiles) / 8)
print("  with >=1 hit: %.3f" % (hit_wi / tot_wi), " avg hits per hit-warp: %.2f" % (hits / max(hit_wi, 1)))
print("  kept by per-warp bbox cull: %.3f (of those hit: %.3f)" % (cull_keep / tot_wi, hit_wi / max(cull_keep, 1)))
print("  blended pairs per pixel: %.1f" % (hits / (len(tiles) * 256)))
print("  nh histogram", nh_hist.tolist())
