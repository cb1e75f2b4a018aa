"""Ad-hoc: parity + CUDA-event timing of ours vs the reference CUDA rasteriser at bench shapes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import numpy as np, torch
import util
from svgir_b200 import raster

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main(P=300000, W=800, H=800, S=4, VS=52):
    case = util.make_case(P, W, H, S=S, VS=VS)
    g = util.pixel_grads(case)
    out, st, bw = util.run_ours(case, grads=g)
    R = st.num_rendered
    print("P", P, "R", R, "visible", int((out["radii"] > 0).sum()))
    nc = st.t["n_contrib"]
    print("n_contrib mean", nc.float().mean().item(), "max", nc.max().item())
    rg = st.t["ranges"].long(); cnt = (rg[:,1]-rg[:,0])
    print("tiles nonempty", int((cnt>0).sum()), "max per tile", int(cnt.max()), "mean nonempty", float(cnt[cnt>0].float().mean()))
    from oracle import ref_cuda
    r, rout, rbw = util.run_ref(case, grads=g)
    print("R ref", rout["num_rendered"], "radii mismatches", int((out["radii"] != rout["radii"]).sum()))
    if R == rout["num_rendered"]:
        rkeys = r.state("keys", (R,), torch.int64); rpl = r.state("point_list", (R,), torch.int32)
        T = ((W+15)//16)*((H+15)//16)
        rranges = r.state("ranges", (T, 2), torch.int32)
        print("keys equal:", bool((rkeys == st.t["sorted_keys"][:R]).all()), "point_list equal:", bool((rpl == st.t["point_list"][:R]).all()),
              "ranges equal:", bool((rranges == st.t["ranges"]).all()))
    rnc = r.state("n_contrib", (H*W,), torch.int32)
    print("n_contrib mismatches vs ref", int((rnc != st.t["n_contrib"]).sum()))
    for k in ("color", "normal", "depth", "opacity", "feature", "vfeature", "weights"):
        d = (out[k] - rout[k]).abs()
        print(f"fwd {k:9s} max abs diff {d.max().item():.3e} bit-equal {bool((out[k]==rout[k]).all())}")
    for k in ("dL_dmeans2D", "dL_dopacity", "dL_dcolors", "dL_dfeatures", "dL_dvfeatures", "dL_dmeans3D",
              "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"):
        a = bw[k].cpu().numpy(); b = rbw[k].cpu().numpy().reshape(a.shape)
        print(f"bwd {k:14s} rel l2 vs ref {util.rel_l2(a,b):.3e}")
    # ---- timing ----
    cam = case["cam"]; t = util.to_cuda(case)
    s = raster.RasterSettings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=t["bg"],
                              scale_modifier=1.0, viewmatrix=t["viewmatrix"], projmatrix=t["projmatrix"], sh_degree=3,
                              campos=t["campos"], patch_bbox=t["patch_bbox"], config=t["config"])
    gt = {k: torch.from_numpy(v).cuda() for k, v in g.items()}
    holder = {}
    def ours_fwd():
        holder["o"], holder["s"] = raster.forward(s, t["means3D"], t["opacity"], t["scales"], t["rotations"], None, t["shs"], None, t["features"], t["vfeatures"])
    def ours_bwd():
        raster.backward(holder["s"], holder["o"]["radii"], gt)
    def ours_both():
        ours_fwd(); ours_bwd()
    print("ours fwd ms", timeit(ours_fwd)); print("ours bwd ms", timeit(ours_bwd)); print("ours fwd+bwd ms", timeit(ours_both))
    P_ = P
    z = lambda *sh: torch.zeros(sh, device="cuda")
    def ref_fwd():
        r.forward(bg=t["bg"], means3D=t["means3D"], features=t["features"], vfeatures=t["vfeatures"], colors=None, opacity=t["opacity"],
                  scales=t["scales"], rotations=t["rotations"], scale_modifier=1.0, viewmatrix=t["viewmatrix"], projmatrix=t["projmatrix"],
                  prcppoint=t["prcppoint"], patchbbox=t["patch_bbox"], tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, H=H, W=W, sh=t["shs"],
                  degree=3, campos=t["campos"], config=t["config"])
    def ref_bwd():
        r.backward(gt["dL_dcolor"], gt["dL_dnormal"], gt["dL_ddepth"], gt["dL_dopacity"], gt["dL_dfeature"], gt["dL_dvfeature"])
    def ref_both():
        ref_fwd(); ref_bwd()
    print("ref fwd ms", timeit(ref_fwd)); print("ref bwd ms", timeit(ref_bwd)); print("ref fwd+bwd ms", timeit(ref_both))

if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    main(*a)
