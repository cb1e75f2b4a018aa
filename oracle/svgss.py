"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of oracle/svgss_oracle.c.

`forward()` / `backward()` mirror the data flow of the reference's
CudaRasterizer::Rasterizer::forward / ::backward
(svgss_rasterization/cuda_rasterizer/rasterizer_impl.cu:209-382, 386-523) on host arrays and
return every intermediate the parity tests compare (radii, keys, point_list, ranges, n_contrib ...).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("svgss_oracle.c", "bvh_oracle.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_bin.restype = C.c_long
    return _LIB


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous"
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def forward(cam, means3D, opacities, scales, rotations, features, vfeatures, *, shs=None,
            colors_precomp=None, sh_degree=3, bg=(0, 0, 0), scale_modifier=1.0,
            config=(1, 1, 1), variant=0, stop_after=None):
    """Runs preprocess -> binning -> forward compositing on the CPU. Returns a dict."""
    L = lib()
    means3D, opacities, scales, rotations = map(_f32, (means3D, opacities, scales, rotations))
    features = _f32(features if features is not None else np.zeros((means3D.shape[0], 0)))
    vfeatures = _f32(vfeatures if vfeatures is not None else np.zeros((means3D.shape[0], 0)))
    shs, colors_precomp = _f32(shs), _f32(colors_precomp)
    P = means3D.shape[0]
    S, VS = features.shape[1], vfeatures.shape[1]
    M = 0 if shs is None else shs.shape[1]
    W, H = cam.W, cam.H
    cfg = np.asarray(config, np.float32)
    bg = np.asarray(bg, np.float32)
    o = dict(P=P, S=S, VS=VS, W=W, H=H, M=M)
    o["radii"] = np.zeros(P, np.int32)
    o["means2D"] = np.zeros((P, 2), np.float32)
    o["depths"] = np.zeros(P, np.float32)
    o["cov3D"] = np.zeros((P, 6), np.float32)
    o["rgb"] = np.zeros((P, 3), np.float32)
    o["clamped"] = np.zeros((P, 3), np.uint8)
    o["normal"] = np.zeros((P, 3), np.float32)
    o["conic_opacity"] = np.zeros((P, 4), np.float32)
    o["Jinv"] = np.zeros((P, 10), np.float32)
    o["viewCos"] = np.zeros(P, np.float32)
    o["lambda"] = np.zeros((P, 2), np.float32)
    o["tiles_touched"] = np.zeros(P, np.uint32)
    L.oracle_preprocess(
        C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(scales), C.c_float(scale_modifier),
        _p(rotations), _p(opacities), _p(shs), None, _p(colors_precomp), _p(cam.viewmatrix),
        _p(cam.projmatrix), _p(cam.patch_bbox), _p(cam.campos), C.c_int(W), C.c_int(H),
        C.c_float(cam.tanfovx), C.c_float(cam.tanfovy), _p(cfg), C.c_int(len(cfg)), C.c_int(variant),
        _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["rgb"]),
        _p(o["clamped"]), _p(o["normal"]), _p(o["conic_opacity"]), _p(o["Jinv"]), _p(o["viewCos"]),
        _p(o["lambda"]), _p(o["tiles_touched"]))
    if stop_after == "preprocess":
        return o
    Rn = int(o["tiles_touched"].astype(np.int64).sum())
    gx, gy = (W + 15) // 16, (H + 15) // 16
    o["keys_unsorted"] = np.zeros(max(Rn, 1), np.uint64)
    o["vals_unsorted"] = np.zeros(max(Rn, 1), np.uint32)
    o["keys"] = np.zeros(max(Rn, 1), np.uint64)
    o["point_list"] = np.zeros(max(Rn, 1), np.uint32)
    o["ranges"] = np.zeros((gx * gy, 2), np.uint32)
    R2 = L.oracle_bin(C.c_int(P), _p(o["means2D"]), _p(o["depths"]), _p(o["radii"]), C.c_int(W),
                      C.c_int(H), _p(o["keys_unsorted"]), _p(o["vals_unsorted"]), _p(o["keys"]),
                      _p(o["point_list"]), _p(o["ranges"]))
    assert R2 == Rn
    o["num_rendered"] = Rn
    for k in ("keys_unsorted", "vals_unsorted", "keys", "point_list"):
        o[k] = o[k][:Rn]
    if stop_after == "bin":
        return o
    colors = colors_precomp if colors_precomp is not None else o["rgb"]
    HW = H * W
    o["final_T"] = np.zeros(HW, np.float32)
    o["final_D"] = np.zeros(HW, np.float32)
    o["n_contrib"] = np.zeros(HW, np.uint32)
    o["color"] = np.zeros((3, H, W), np.float32)
    o["normal_img"] = np.zeros((3, H, W), np.float32)
    o["depth"] = np.zeros((1, H, W), np.float32)
    o["opacity"] = np.zeros((1, H, W), np.float32)
    o["feature"] = np.zeros((S, H, W), np.float32)
    o["vfeature"] = np.zeros((VS // 4, H, W), np.float32)
    o["weights"] = np.zeros((P, 1), np.float32)
    L.oracle_render_fwd(
        C.c_int(variant), C.c_int(W), C.c_int(H), C.c_int(S), C.c_int(VS), _p(o["ranges"]),
        _p(o["point_list"]), _p(o["means2D"]), _p(features), _p(vfeatures), _p(colors),
        _p(o["normal"]), _p(o["depths"]), _p(o["conic_opacity"]), _p(o["Jinv"]), _p(o["lambda"]),
        _p(bg), _p(cfg), C.c_int(len(cfg)), _p(o["final_T"]), _p(o["final_D"]), _p(o["n_contrib"]),
        _p(o["color"]), _p(o["normal_img"]), _p(o["depth"]), _p(o["opacity"]), _p(o["feature"]),
        _p(o["vfeature"]), _p(o["weights"]))
    o["_inputs"] = dict(cam=cam, means3D=means3D, opacities=opacities, scales=scales,
                        rotations=rotations, features=features, vfeatures=vfeatures, shs=shs,
                        colors_precomp=colors_precomp, sh_degree=sh_degree, bg=bg,
                        scale_modifier=scale_modifier, cfg=cfg, variant=variant, colors=colors)
    return o


def backward(fw, dL_dcolor, dL_dnormal, dL_ddepth, dL_dopac, dL_dfeature, dL_dvfeature,
             backward_geometry=True):
    """Gradient of the forward `fw` (dict returned by forward()). Returns a dict of float32 arrays
    named like the reference's 13-tuple (rasterize_points.cu:264)."""
    L = lib()
    i = fw["_inputs"]
    cam = i["cam"]
    P, S, VS, W, H, M = (fw[k] for k in ("P", "S", "VS", "W", "H", "M"))
    g = [np.ascontiguousarray(x, np.float32) for x in
         (dL_dcolor, dL_dnormal, dL_ddepth, dL_dopac, dL_dfeature, dL_dvfeature)]
    d = dict(mean2D=np.zeros((P, 3)), conic=np.zeros((P, 4)), opacity=np.zeros((P, 1)),
             colors=np.zeros((P, 3)), normal=np.zeros((P, 3)), depth=np.zeros((P, 1)),
             features=np.zeros((P, S)), vfeatures=np.zeros((P, VS)))
    L.oracle_render_bwd(
        C.c_int(i["variant"]), C.c_int(W), C.c_int(H), C.c_int(S), C.c_int(VS), _p(fw["ranges"]),
        _p(fw["point_list"]), _p(fw["means2D"]), _p(i["features"]), _p(i["vfeatures"]),
        _p(i["colors"]), _p(fw["normal"]), _p(fw["depths"]), _p(fw["conic_opacity"]), _p(fw["Jinv"]),
        _p(fw["lambda"]), _p(i["bg"]), _p(i["cfg"]), C.c_int(len(i["cfg"])), _p(fw["final_T"]),
        _p(fw["final_D"]), _p(fw["n_contrib"]), _p(g[0]), _p(g[1]), _p(g[2]), _p(g[3]), _p(g[4]),
        _p(g[5]), C.c_int(1 if backward_geometry else 0), _p(d["mean2D"]), _p(d["conic"]),
        _p(d["opacity"]), _p(d["colors"]), _p(d["normal"]), _p(d["depth"]), _p(d["features"]),
        _p(d["vfeatures"]))
    r = {k: v.astype(np.float32) for k, v in d.items()}
    out = dict(dL_dmeans2D=r["mean2D"], dL_dconic=r["conic"], dL_dopacity=r["opacity"],
               dL_dcolors=r["colors"], dL_dnormal=r["normal"], dL_ddepth=r["depth"],
               dL_dfeatures=r["features"], dL_dvfeatures=r["vfeatures"])
    out["dL_dmeans3D"] = np.zeros((P, 3), np.float32)
    out["dL_dcov3D"] = np.zeros((P, 6), np.float32)
    out["dL_dsh"] = np.zeros((P, M, 3), np.float32)
    out["dL_dscales"] = np.zeros((P, 3), np.float32)
    out["dL_drotations"] = np.zeros((P, 4), np.float32)
    fx = W / (2.0 * cam.tanfovx)
    fy = H / (2.0 * cam.tanfovy)
    L.oracle_preprocess_bwd(
        C.c_int(P), C.c_int(i["sh_degree"]), C.c_int(M), _p(i["means3D"]), _p(fw["radii"]),
        _p(i["shs"]), _p(fw["clamped"]), _p(i["scales"]), _p(i["rotations"]),
        C.c_float(i["scale_modifier"]), _p(fw["cov3D"]), _p(cam.viewmatrix), _p(cam.projmatrix),
        C.c_float(fx), C.c_float(fy), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy),
        _p(cam.campos), _p(i["cfg"]), C.c_int(len(i["cfg"])), C.c_int(i["variant"]),
        _p(out["dL_dmeans2D"]), _p(out["dL_dconic"]), _p(out["dL_dcolors"]), _p(out["dL_dnormal"]),
        _p(np.ascontiguousarray(out["dL_ddepth"].reshape(-1))), _p(out["dL_dmeans3D"]),
        _p(out["dL_dcov3D"]), _p(out["dL_dsh"]), _p(out["dL_dscales"]), _p(out["dL_drotations"]))
    return out
