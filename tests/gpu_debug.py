"""Ad-hoc differential run (ours vs C oracle vs reference CUDA) used while developing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import numpy as np, torch
import util

def main(P=10000, W=200, H=200):
    case = util.make_case(P, W, H)
    g = util.pixel_grads(case)
    ofw, obw = util.run_oracle(case, grads=g)
    out, st, bw = util.run_ours(case, grads=g)
    print("R ours", st.num_rendered, "oracle", ofw["num_rendered"])
    rad = out["radii"].cpu().numpy()
    print("radii mismatches vs oracle:", int((rad != ofw["radii"]).sum()))
    keys = st.t["sorted_keys"][:st.num_rendered].cpu().numpy().astype(np.uint64)
    pl = st.t["point_list"][:st.num_rendered].cpu().numpy().astype(np.uint32)
    if st.num_rendered == ofw["num_rendered"]:
        print("keys equal:", bool((keys == ofw["keys"]).all()), "point_list equal:", bool((pl == ofw["point_list"]).all()))
        print("ranges equal:", bool((st.t["ranges"].cpu().numpy().astype(np.uint32) == ofw["ranges"]).all()))
    for k, ok in (("color", "color"), ("normal", "normal_img"), ("depth", "depth"), ("opacity", "opacity"),
                  ("feature", "feature"), ("vfeature", "vfeature"), ("weights", "weights")):
        a = out[k].cpu().numpy(); b = ofw[ok]
        print(f"fwd {k:9s} max abs diff vs oracle {np.abs(a-b).max():.3e}  frac>1e-5 {(np.abs(a-b)>1e-5).mean():.2e}")
    nc = st.t["n_contrib"].cpu().numpy().astype(np.uint32)
    print("n_contrib mismatches vs oracle", int((nc != ofw["n_contrib"]).sum()))
    for k in ("dL_dmeans2D", "dL_dconic", "dL_dopacity", "dL_dcolors", "dL_dnormal", "dL_ddepth", "dL_dfeatures",
              "dL_dvfeatures", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"):
        a = bw[k].cpu().numpy().reshape(obw[k].shape); b = obw[k]
        print(f"bwd {k:14s} rel l2 vs oracle {util.rel_l2(a,b):.3e}")
    from oracle import ref_cuda
    if ref_cuda.available():
        r, rout, rbw = util.run_ref(case, grads=g)
        print("R ref", rout["num_rendered"])
        print("radii mismatches vs ref:", int((out["radii"] != rout["radii"]).sum()))
        R = rout["num_rendered"]
        rkeys = r.state("keys", (R,), torch.int64)
        rpl = r.state("point_list", (R,), torch.int32)
        T = ((W+15)//16)*((H+15)//16)
        rranges = r.state("ranges", (T, 2), torch.int32)
        if R == st.num_rendered:
            print("keys equal ref:", bool((rkeys == st.t["sorted_keys"][:R]).all()), "point_list equal ref:",
                  bool((rpl == st.t["point_list"][:R]).all()), "ranges equal ref:", bool((rranges == st.t["ranges"]).all()))
        rnc = r.state("n_contrib", (H*W,), torch.int32)
        print("n_contrib mismatches vs ref", int((rnc != st.t["n_contrib"]).sum()))
        for k in ("color", "normal", "depth", "opacity", "feature", "vfeature", "weights"):
            d = (out[k] - rout[k]).abs()
            print(f"fwd {k:9s} max abs diff vs ref {d.max().item():.3e} frac>1e-5 {(d>1e-5).float().mean().item():.2e}  bit-equal {bool((out[k]==rout[k]).all())}")
        for k in ("dL_dmeans2D", "dL_dopacity", "dL_dcolors", "dL_dfeatures", "dL_dvfeatures", "dL_dmeans3D",
                  "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations", "dL_dnormal", "dL_ddepth"):
            a = bw[k].cpu().numpy(); b = rbw[k].cpu().numpy().reshape(a.shape)
            print(f"bwd {k:14s} rel l2 vs ref {util.rel_l2(a,b):.3e}")
        # oracle vs ref
        print("oracle radii mismatches vs ref:", int((torch.from_numpy(ofw['radii']).cuda() != rout["radii"]).sum()))
        if R == ofw["num_rendered"]:
            print("oracle keys equal ref:", bool((rkeys.cpu().numpy().astype(np.uint64) == ofw["keys"]).all()))

if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    main(*a)
