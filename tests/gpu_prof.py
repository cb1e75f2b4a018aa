"""Ad-hoc: run ours fwd+bwd a few times (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch, util
from svgir_b200 import raster
P, W, H, S, VS, N = [int(x) for x in sys.argv[1:7]]
case = util.make_case(P, W, H, S=S, VS=VS)
g = util.pixel_grads(case)
cam = case["cam"]; t = util.to_cuda(case)
s = raster.RasterSettings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=t["bg"],
                          scale_modifier=1.0, viewmatrix=t["viewmatrix"], projmatrix=t["projmatrix"], sh_degree=3,
                          campos=t["campos"], patch_bbox=t["patch_bbox"], config=t["config"])
gt = {k: torch.from_numpy(v).cuda() for k, v in g.items()}
for _ in range(N):
    o, st = raster.forward(s, t["means3D"], t["opacity"], t["scales"], t["rotations"], None, t["shs"], None, t["features"], t["vfeatures"])
    raster.backward(st, o["radii"], gt)
torch.cuda.synchronize()
