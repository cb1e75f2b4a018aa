"""TEST INFRASTRUCTURE ONLY -- ctypes driver of oracle/_ref/libsvgss_ref.so, i.e. the UNMODIFIED
reference CUDA rasteriser (svgss_rasterization/cuda_rasterizer/*.cu) behind the C harness in
oracle/ref_harness_svgss.cu. Used by the GPU differential tests, the golden-vector generator and
bench.py's `--impl reference_cuda` leg. Needs a GPU."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def available(name="svgss") -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", f"lib{name}_ref.so"))


def _lib(name="svgss"):
    if name not in _LIBS:
        L = C.CDLL(os.path.join(_HERE, "_ref", f"lib{name}_ref.so"))
        getattr(L, f"ref_{name}_create").restype = C.c_void_p
        getattr(L, f"ref_{name}_state").restype = C.c_void_p
        getattr(L, f"ref_{name}_state").argtypes = [C.c_void_p, C.c_char_p]
        getattr(L, f"ref_{name}_destroy").argtypes = [C.c_void_p]
        _LIBS[name] = L
    return _LIBS[name]


def _p(t):
    return C.c_void_p(0 if t is None or t.numel() == 0 else t.data_ptr())


class RefSvgss:
    """One forward (+ optional backward) of the reference stage-2 rasteriser."""

    def __init__(self):
        self.L = _lib("svgss")
        self.h = C.c_void_p(self.L.ref_svgss_create())

    def close(self):
        if self.h:
            self.L.ref_svgss_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, *, bg, means3D, features, vfeatures, colors, opacity, scales, rotations,
                scale_modifier, viewmatrix, projmatrix, prcppoint, patchbbox, tanfovx, tanfovy, H, W, sh,
                degree, campos, config, debug=False):
        dev = means3D.device
        P, S, VS = means3D.shape[0], features.shape[1], vfeatures.shape[1]
        M = sh.shape[1] if sh is not None and sh.numel() else 0
        f = dict(dtype=torch.float32, device=dev)
        o = dict(color=torch.zeros((3, H, W), **f), normal=torch.zeros((3, H, W), **f),
                 depth=torch.zeros((1, H, W), **f), opacity=torch.zeros((1, H, W), **f),
                 feature=torch.zeros((S, H, W), **f), vfeature=torch.zeros((VS // 4, H, W), **f),
                 weights=torch.zeros((P, 1), **f), radii=torch.zeros((P,), dtype=torch.int32, device=dev))
        self.args = dict(bg=bg, means3D=means3D, features=features, vfeatures=vfeatures, colors=colors,
                         opacity=opacity, scales=scales, rotations=rotations, scale_modifier=scale_modifier,
                         viewmatrix=viewmatrix, projmatrix=projmatrix, prcppoint=prcppoint,
                         patchbbox=patchbbox, tanfovx=tanfovx, tanfovy=tanfovy, H=H, W=W, sh=sh, degree=degree,
                         campos=campos, config=config, P=P, S=S, VS=VS, M=M)
        torch.cuda.synchronize()
        R = self.L.ref_svgss_forward(
            self.h, P, S, VS, degree, M, _p(bg), W, H, _p(means3D), _p(sh), _p(colors), _p(features),
            _p(vfeatures), _p(opacity), _p(scales), C.c_float(scale_modifier), _p(rotations), _p(None),
            _p(viewmatrix), _p(projmatrix), _p(prcppoint), _p(patchbbox), _p(campos), C.c_float(tanfovx),
            C.c_float(tanfovy), 0, _p(config), _p(o["color"]), _p(o["normal"]), _p(o["depth"]),
            _p(o["opacity"]), _p(o["feature"]), _p(o["vfeature"]), _p(o["weights"]), _p(o["radii"]),
            int(debug))
        if R < 0:
            raise RuntimeError("reference forward failed")
        o["num_rendered"] = R
        self.out = o
        return o

    def state(self, name, shape, dtype):
        """Copy of one internal state array of the reference (see ref_harness_svgss.cu)."""
        ptr = self.L.ref_svgss_state(self.h, name.encode())
        if not ptr:
            raise KeyError(name)
        n = 1
        for s in shape:
            n *= s
        out = torch.empty(shape, dtype=dtype, device="cuda")
        if n:
            nbytes = n * out.element_size()
            rt = C.CDLL("libcudart.so")
            rt.cudaMemcpy(C.c_void_p(out.data_ptr()), C.c_void_p(ptr), C.c_size_t(nbytes), 3)
        torch.cuda.synchronize()
        return out

    def backward(self, dL_dcolor, dL_dnormal, dL_ddepth, dL_dopac, dL_dfeature, dL_dvfeature):
        a = self.args
        P, S, VS, M = a["P"], a["S"], a["VS"], a["M"]
        dev = a["means3D"].device
        f = dict(dtype=torch.float32, device=dev)
        g = dict(dL_dmeans2D=torch.zeros((P, 3), **f), dL_dconic=torch.zeros((P, 2, 2), **f),
                 dL_dopacity=torch.zeros((P, 1), **f), dL_dcolors=torch.zeros((P, 3), **f),
                 dL_dfeatures=torch.zeros((P, S), **f), dL_dvfeatures=torch.zeros((P, VS), **f),
                 dL_dnormal=torch.zeros((P, 3), **f), dL_ddepth=torch.zeros((P, 1), **f),
                 dL_dmeans3D=torch.zeros((P, 3), **f), dL_dcov3D=torch.zeros((P, 6), **f),
                 dL_dsh=torch.zeros((P, M, 3), **f), dL_dscales=torch.zeros((P, 3), **f),
                 dL_drotations=torch.zeros((P, 4), **f), dL_dviewmat=torch.zeros((4, 4), **f),
                 dL_dprojmat=torch.zeros((4, 4), **f), dL_dcampos=torch.zeros((3,), **f))
        torch.cuda.synchronize()
        rc = self.L.ref_svgss_backward(
            self.h, P, S, VS, a["degree"], M, _p(a["bg"]), a["W"], a["H"], _p(a["means3D"]), _p(a["sh"]),
            _p(a["features"]), _p(a["vfeatures"]), _p(a["colors"]), _p(a["scales"]),
            C.c_float(a["scale_modifier"]), _p(a["rotations"]), _p(None), _p(a["viewmatrix"]),
            _p(a["projmatrix"]), _p(a["campos"]), _p(a["prcppoint"]), _p(a["patchbbox"]),
            C.c_float(a["tanfovx"]), C.c_float(a["tanfovy"]), _p(self.out["radii"]), _p(dL_dcolor),
            _p(dL_dnormal), _p(dL_ddepth), _p(dL_dopac), _p(dL_dfeature), _p(dL_dvfeature),
            _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]),
            _p(g["dL_dfeatures"]), _p(g["dL_dvfeatures"]), _p(g["dL_dnormal"]), _p(g["dL_ddepth"]),
            _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]),
            _p(g["dL_drotations"]), _p(g["dL_dviewmat"]), _p(g["dL_dprojmat"]), _p(g["dL_dcampos"]), 0,
            _p(a["config"]))
        if rc != 0:
            raise RuntimeError("reference backward failed")
        return g


class RefRgss:
    """One forward (+ optional backward) of the reference stage-1 rasteriser
    (rgss-rasterization/cuda_rasterizer, harness oracle/ref_harness_rgss.cu)."""

    def __init__(self):
        self.L = _lib("rgss")
        self.h = C.c_void_p(self.L.ref_rgss_create())

    def close(self):
        if self.h:
            self.L.ref_rgss_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, *, bg, means3D, features, colors, opacity, scales, rotations, scale_modifier, viewmatrix,
                projmatrix, tanfovx, tanfovy, cx, cy, H, W, sh, degree, campos, computer_pseudo_normal=False,
                debug=False):
        dev = means3D.device
        P, S = means3D.shape[0], features.shape[1]
        M = sh.shape[1] if sh is not None and sh.numel() else 0
        f = dict(dtype=torch.float32, device=dev)
        o = dict(color=torch.zeros((3, H, W), **f), normal=torch.zeros((3, H, W), **f),
                 opacity=torch.zeros((1, H, W), **f), depth=torch.zeros((1, H, W), **f),
                 feature=torch.zeros((S, H, W), **f), pseudo_normal=torch.zeros((3, H, W), **f),
                 surface_xyz=torch.zeros((3, H, W), **f), weights=torch.zeros((P, 1), **f),
                 radii=torch.zeros((P,), dtype=torch.int32, device=dev))
        self.args = dict(bg=bg, means3D=means3D, features=features, colors=colors, opacity=opacity, scales=scales,
                         rotations=rotations, scale_modifier=scale_modifier, viewmatrix=viewmatrix,
                         projmatrix=projmatrix, tanfovx=tanfovx, tanfovy=tanfovy, H=H, W=W, sh=sh, degree=degree,
                         campos=campos, P=P, S=S, M=M)
        torch.cuda.synchronize()
        R = self.L.ref_rgss_forward(
            self.h, P, S, degree, M, _p(bg), W, H, _p(means3D), _p(sh), _p(colors), _p(features), _p(opacity),
            _p(scales), C.c_float(scale_modifier), _p(rotations), _p(None), _p(viewmatrix), _p(projmatrix),
            _p(campos), C.c_float(tanfovx), C.c_float(tanfovy), C.c_float(cx), C.c_float(cy), 0,
            int(computer_pseudo_normal), _p(o["color"]), _p(o["normal"]), _p(o["opacity"]), _p(o["depth"]),
            _p(o["feature"]), _p(o["pseudo_normal"]), _p(o["surface_xyz"]), _p(o["weights"]), _p(o["radii"]),
            int(debug))
        if R < 0:
            raise RuntimeError("reference rgss forward failed")
        o["num_rendered"] = R
        self.out = o
        return o

    def state(self, name, shape, dtype):
        ptr = self.L.ref_rgss_state(self.h, name.encode())
        if not ptr:
            raise KeyError(name)
        out = torch.empty(shape, dtype=dtype, device="cuda")
        if out.numel():
            rt = C.CDLL("libcudart.so")
            rt.cudaMemcpy(C.c_void_p(out.data_ptr()), C.c_void_p(ptr), C.c_size_t(out.numel() * out.element_size()), 3)
        torch.cuda.synchronize()
        return out

    def backward(self, dL_dcolor, dL_dnormal, dL_dopac, dL_ddepth, dL_dfeature, backward_geometry=True):
        a = self.args
        P, S, M = a["P"], a["S"], a["M"]
        f = dict(dtype=torch.float32, device=a["means3D"].device)
        g = dict(dL_dmeans2D=torch.zeros((P, 3), **f), dL_dconic=torch.zeros((P, 2, 2), **f),
                 dL_dopacity=torch.zeros((P, 1), **f), dL_dcolors=torch.zeros((P, 3), **f),
                 dL_dnormal=torch.zeros((P, 3), **f), dL_ddepth=torch.zeros((P, 1), **f),
                 dL_dfeatures=torch.zeros((P, S), **f), dL_dmeans3D=torch.zeros((P, 3), **f),
                 dL_dcov3D=torch.zeros((P, 6), **f), dL_dsh=torch.zeros((P, M, 3), **f),
                 dL_dscales=torch.zeros((P, 3), **f), dL_drotations=torch.zeros((P, 4), **f))
        torch.cuda.synchronize()
        rc = self.L.ref_rgss_backward(
            self.h, P, S, a["degree"], M, _p(a["bg"]), a["W"], a["H"], _p(a["means3D"]), _p(a["sh"]),
            _p(a["features"]), _p(a["colors"]), _p(a["scales"]), C.c_float(a["scale_modifier"]), _p(a["rotations"]),
            _p(None), _p(a["viewmatrix"]), _p(a["projmatrix"]), _p(a["campos"]), C.c_float(a["tanfovx"]),
            C.c_float(a["tanfovy"]), _p(self.out["radii"]), _p(dL_dcolor), _p(dL_dnormal), _p(dL_dopac),
            _p(dL_ddepth), _p(dL_dfeature), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]),
            _p(g["dL_dcolors"]), _p(g["dL_dnormal"]), _p(g["dL_ddepth"]), _p(g["dL_dfeatures"]),
            _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]),
            int(backward_geometry), 0)
        if rc != 0:
            raise RuntimeError("reference rgss backward failed")
        return g


class RefReq:
    """The reference's SH render_equation kernels (rgss-rasterization/render_equation.cu) behind
    oracle/ref_harness_req.cu. Inputs: dict from oracle.render_equation_sh_oracle.make_inputs (CUDA tensors)."""
    ORDER = ("base_color", "roughness", "metallic", "normals", "viewdirs", "incidents_shs", "direct_shs", "visibility_shs")

    def __init__(self):
        import torch  # noqa: F401  (libtorch must be loaded before the harness)
        self.L = C.CDLL(os.path.join(_HERE, "_ref", "libreq_ref.so"))

    def _sizes(self, t):
        return t["base_color"].shape[0], t["incidents_shs"].shape[1], t["direct_shs"].shape[1], t["visibility_shs"].shape[1]

    def forward(self, t, sample_num, is_training=False, rand_float=None):
        P, Si, Sd, Sv = self._sizes(t)
        f = dict(dtype=torch.float32, device=t["base_color"].device)
        o = dict(incident_dirs=torch.zeros((P, sample_num, 3), **f), pbr=torch.zeros((P, 3), **f),
                 diffuse_light=torch.zeros((P, 3), **f))
        torch.cuda.synchronize()
        rc = self.L.ref_req_forward(P, Si, Sd, Sv, *[_p(t[k]) for k in self.ORDER], sample_num, int(is_training),
                                    _p(rand_float), _p(o["incident_dirs"]), _p(o["pbr"]), _p(o["diffuse_light"]))
        if rc != 0:
            raise RuntimeError("reference render_equation forward failed")
        return o

    def forward_complex(self, t, sample_num):
        P, Si, Sd, Sv = self._sizes(t)
        f = dict(dtype=torch.float32, device=t["base_color"].device)
        z = lambda *s: torch.zeros(s, **f)
        o = dict(incident_dirs=z(P, sample_num, 3), pbr=z(P, 3), incident_lights=z(P, sample_num, 3),
                 local_incident_lights=z(P, sample_num, 3), global_incident_lights=z(P, sample_num, 3),
                 incident_visibility=z(P, sample_num, 1), diffuse_light=z(P, 3), local_diffuse_light=z(P, 3),
                 accum=z(P, 1), rgb_d=z(P, 3), rgb_s=z(P, 3))
        torch.cuda.synchronize()
        rc = self.L.ref_req_forward_complex(P, Si, Sd, Sv, *[_p(t[k]) for k in self.ORDER], sample_num,
                                            *[_p(o[k]) for k in ("incident_dirs", "pbr", "incident_lights",
                                                                 "local_incident_lights", "global_incident_lights",
                                                                 "incident_visibility", "diffuse_light",
                                                                 "local_diffuse_light", "accum", "rgb_d", "rgb_s")])
        if rc != 0:
            raise RuntimeError("reference render_equation forward_complex failed")
        return o

    def backward(self, t, sample_num, incident_dirs, g_pbr, g_dl):
        P, Si, Sd, Sv = self._sizes(t)
        f = dict(dtype=torch.float32, device=t["base_color"].device)
        z = lambda *s: torch.zeros(s, **f)
        g = dict(dL_dbase_color=z(P, 3), dL_droughness=z(P, 1), dL_dmetallic=z(P, 1), dL_dnormals=z(P, 3),
                 dL_dviewdirs=z(P, 3), dL_dincidents_shs=z(P, Si, 3), dL_ddirect_shs=z(1, Sd, 3),
                 dL_dvisibility_shs=z(P, Sv, 1))
        torch.cuda.synchronize()
        rc = self.L.ref_req_backward(P, Si, Sd, Sv, *[_p(t[k]) for k in self.ORDER], sample_num, _p(incident_dirs),
                                     _p(g_pbr), _p(g_dl), *[_p(v) for v in g.values()])
        if rc != 0:
            raise RuntimeError("reference render_equation backward failed")
        return g


class RefBvh:
    """The reference LBVH (submodules/bvh construct.cu / trace.cu, unmodified) behind
    oracle/ref_harness_bvh.cu; mirrors RayTracer (submodules/bvh/__init__.py:28-71)."""

    def __init__(self, nodes: torch.Tensor, aabbs: torch.Tensor, means3D, scales, rotations):
        self.L = C.CDLL(os.path.join(_HERE, "_ref", "libbvh_ref.so"))
        P = means3D.shape[0]
        self.nodes, self.aabbs = nodes.contiguous().clone(), aabbs.contiguous().clone()
        self.morton = torch.zeros((P,), dtype=torch.int64, device=means3D.device)
        rc = self.L.ref_bvh_build(P, _p(means3D.contiguous()), _p(scales.contiguous()), _p(rotations.contiguous()),
                                  _p(self.nodes), _p(self.aabbs), _p(self.morton))
        if rc != 0:
            raise RuntimeError(f"ref_bvh_build failed ({rc})")

    def trace_opacity(self, rays_o, rays_d, means3D, cov_inv, opacity, normals):
        rays_o, rays_d = rays_o.contiguous(), rays_d.contiguous()
        shape = rays_o.shape[:-1]
        contrib = torch.zeros(shape, dtype=torch.int32, device=rays_o.device)
        vis = torch.ones(shape, dtype=torch.float32, device=rays_o.device)
        rc = self.L.ref_bvh_trace_opacity(rays_o.numel() // 3, _p(self.nodes), _p(self.aabbs), _p(rays_o), _p(rays_d),
                                          _p(means3D.contiguous()), _p(cov_inv.contiguous()), _p(opacity.contiguous()),
                                          _p(normals.contiguous()), _p(contrib), _p(vis))
        if rc != 0:
            raise RuntimeError(f"ref_bvh_trace_opacity failed ({rc})")
        return contrib, vis
