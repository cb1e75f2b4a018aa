"""Ad-hoc: CUDA-event timing of the fused shading kernels (library timers) at bench shape, with and
without the env-map gradient."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
from svgir_b200 import scene, shading, _lib

N, Ns = int(sys.argv[1]), int(sys.argv[2])
cl = scene.make_surfels(N, seed=3)
m = scene.make_materials(cl, Ns, seed=4)
cam = scene.look_at_camera(800, 800, 0)
d = lambda a: torch.from_numpy(a).cuda()
t = {k: d(v) for k, v in m.items()}
vd = torch.nn.functional.normalize(d(cam.campos)[None] - d(cl.means3D), dim=-1)
view = d(cam.viewmatrix[:3, :3].copy())
for env_grad in (True, False):
    for k in ("base_color", "roughness", "shading_normals"):
        t[k].requires_grad_(True)
    t["env_param"].requires_grad_(env_grad)
    for it in range(6):
        if it == 2:
            _lib.timing_collect(reset=True); _lib.timing_enable(True)
        f, vf = shading.shade_and_pack(t["base_color"], t["roughness"], t["shading_normals"], vd, t["radiance"],
                                       (t["env_param"], shading.MODE_LEARNABLE), t["visibility"], t["incident_dirs"],
                                       t["incident_areas"], view, is_training=True)
        (f.sum() + vf.sum()).backward()
    _lib.timing_enable(False)
    fw, bw = _lib.timing_collect("shade_fwd"), _lib.timing_collect("shade_bwd")
    print("env_grad=%s  shade_fwd %.4f ms  shade_bwd %.4f ms" % (env_grad, fw[0] / fw[1], bw[0] / bw[1]))
    _lib.timing_collect(reset=True)
