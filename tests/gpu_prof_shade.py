"""Ad-hoc: run the fused shading fwd+bwd a few times at bench shape (for ncu), print golden error stats."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import numpy as np, torch
from svgir_b200 import scene, shading

N, Ns, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
if len(sys.argv) > 4:
    GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for tag in ("train_small", "eval_small"):
        g = dict(np.load(os.path.join(GOLD, f"ref_shading_{tag}.npz")))
        t = {k[3:]: torch.tensor(v).cuda() for k, v in g.items() if k.startswith("in_")}
        r = shading.shade_surfels(t["base_color"], t["roughness"], t["shading_normals"], t["viewdirs"], t["radiance"],
                                  (t["env_param"], shading.MODE_LEARNABLE), t["visibility"], t["incident_dirs"], t["incident_areas"])
        for k in ("pbr", "specular", "direct"):
            a, b, b64 = r[k].cpu().numpy().astype(np.float64), g["out_" + k].astype(np.float64), g["out64_" + k]
            ea, eb = np.abs(a - b64) / (1e-5 / 3e-5 + np.abs(b64)), np.abs(b - b64) / (1e-5 / 3e-5 + np.abs(b64))
            print(tag, k, "ours: max %.2e p99 %.2e l2 %.2e | ref32: max %.2e p99 %.2e l2 %.2e" % (
                ea.max(), np.quantile(ea, 0.99), np.linalg.norm(a - b64) / np.linalg.norm(b64),
                eb.max(), np.quantile(eb, 0.99), np.linalg.norm(b - b64) / np.linalg.norm(b64)))
cl = scene.make_surfels(N, seed=3)
m = scene.make_materials(cl, Ns, seed=4)
cam = scene.look_at_camera(800, 800, 0)
d = lambda a: torch.from_numpy(a).cuda()
t = {k: d(v) for k, v in m.items()}
vd = torch.nn.functional.normalize(d(cam.campos)[None] - d(cl.means3D), dim=-1)
for k in ("base_color", "roughness", "shading_normals", "env_param"):
    t[k].requires_grad_(True)
view = d(cam.viewmatrix[:3, :3].copy())
for _ in range(reps):
    f, vf = shading.shade_and_pack(t["base_color"], t["roughness"], t["shading_normals"], vd, t["radiance"],
                                   (t["env_param"], shading.MODE_LEARNABLE), t["visibility"], t["incident_dirs"],
                                   t["incident_areas"], view, is_training=True)
    (f.sum() + vf.sum()).backward()
torch.cuda.synchronize()
print("done")
