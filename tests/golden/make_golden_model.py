"""Generates tests/golden/ref_model.npz by running the REFERENCE's own GaussianModel code on CPU:

  A  densify_and_prune (scene/gaussian_model.py:1229-1250 -> densify_and_clone :1189, densify_and_split :1136,
     densification_postfix :1098, cat_tensors_to_optimizer :1064, prune_points :1038, _prune_optimizer :1020) on a seeded
     300-surfel PBR model with a live torch.optim.Adam (training_setup :737-773) that has taken two steps;
  B  replace_nangrad_to_zero (:775-800) on gradients with NaNs;
  C  get_radiance_loss (:544-575) up to -- and after -- its one Slang call: `renderer.render_irradiance_sample` is
     replaced by a recorder that returns a seeded tensor, so max_idx, the envmap = direct_light(dirs) * areas, the
     transposed normal layout, the gathered target and the L1 value are the reference's own;
  D  the chunking of update_radiace (:487-497): chunk boundaries and the per-chunk torch.rand draws of
     sample_incident_rays (its tracer calls are replaced by recorders);
  E  checkpoints: ref_checkpoint.pth = torch.save((capture(), iteration)) of a reference model with a live optimiser
     (:195-225), and the reverse direction checked HERE: a checkpoint written by svgir_b200.io.save_checkpoint with a
     FusedAdam.state_dict() is restored by the reference's restore() (:227-268), tensors and Adam state equal, and the
     restored optimiser steps;
  F  PLY: the structured array the reference's save_ply builds (:855-881) for a stage-1 model, recorded at the
     PlyElement.describe call; the reference's load_ply (:891-1003) reading files written by svgir_b200.io.save_ply
     (served through svgir_b200.io.read_ply_vertices, `plyfile` being absent) reproduces the model, and its PBR branch
     equals io.load_ply(reference_roughness_quirk=True); the reference's own writer raises for a PBR model.

Run in the build container only:   python tests/golden/make_golden_model.py

The reference hard-codes device="cuda" in tensor factories and `.cuda()` calls; there is no GPU here, so those are
redirected to the CPU for the duration of the run (torch factories wrapped, Tensor.cuda = identity). Absent third-party
modules (plyfile, simple_knn, custom_knn, slangtorch, kornia, ...) are stubs; none is touched by the code above
except `torch.normal`, which is wrapped to RECORD the standard-normal draws
z = (sample - mean) / std so that the checker can be fed the same ones.
"""
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

_FACTORIES = ("tensor", "zeros", "ones", "empty", "full", "rand", "randn", "arange", "linspace", "zeros_like", "ones_like",
              "rand_like", "randint", "eye", "as_tensor")


def redirect_cuda_to_cpu():
    for name in _FACTORIES:
        orig = getattr(torch, name)

        def wrap(*a, __orig=orig, **k):
            if "device" in k and str(k["device"]).startswith("cuda"):
                k["device"] = "cpu"
            return __orig(*a, **k)
        setattr(torch, name, wrap)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.empty_cache = lambda: None
    torch.cuda.synchronize = lambda *a, **k: None


def import_reference():
    for name in ["plyfile", "simple_knn", "simple_knn._C", "custom_knn", "custom_knn._C", "slangtorch", "trimesh", "pyexr",
                 "nvdiffrast", "nvdiffrast.torch", "kornia", "kornia.filters", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.colors", "matplotlib.cm", "imageio", "imageio.plugins", "imageio.plugins.freeimage", "cv2",
                 "lpips", "dearpygui", "dearpygui.dearpygui", "svgss_rasterization", "pbgi", "pbgi.renderer", "submodules",
                 "submodules.bvh", "torchvision", "torchvision.utils", "torchvision.transforms",
                 "torchvision.transforms.functional", "PIL", "PIL.Image", "open3d", "scipy.spatial.transform"]:
        if name not in sys.modules:
            m = mock.MagicMock(name=name)
            m.__path__ = []
            m.__spec__ = None
            sys.modules[name] = m
    sys.path.insert(0, REF)
    import torch.utils.cpp_extension as cpp_ext
    cpp_ext.load = lambda *a, **k: mock.MagicMock(name="jit_ext")
    import scene.gaussian_model as gm
    import scene.direct_light_map as dlm
    return gm, dlm


class Opt:   # arguments/__init__.py:78-98 (OptimizationParams defaults)
    position_lr_init, position_lr_final, position_lr_delay_mult, position_lr_max_steps = 0.00016, 0.0000016, 0.01, 30_000
    normal_lr, sh_lr, opacity_lr, scaling_lr, rotation_lr = 0.01, 0.0025, 0.05, 0.005, 0.001
    base_color_lr, roughness_lr, light_lr, light_rest_lr, visibility_lr, visibility_rest_lr = 0.01, 0.01, 0.001, 0.0001, 0.0025, 0.0025
    percent_dense = 0.001


GROUPS = ("xyz", "normal", "rotation", "scaling", "opacity", "f_dc", "f_rest", "base_color", "roughness", "incidents_dc",
          "incidents_rest", "visibility_dc", "visibility_rest")
ATTR = {"xyz": "_xyz", "normal": "_normal", "rotation": "_rotation", "scaling": "_scaling", "opacity": "_opacity",
        "f_dc": "_shs_dc", "f_rest": "_shs_rest", "base_color": "_base_color", "roughness": "_roughness",
        "incidents_dc": "_incidents_dc", "incidents_rest": "_incidents_rest", "visibility_dc": "_visibility_dc",
        "visibility_rest": "_visibility_rest"}


def make_model(gm, P, seed):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    m = gm.GaussianModel(3, render_type="render_relight")
    m.spatial_lr_scale = 1.3
    shapes = {"xyz": (P, 3), "normal": (P, 12), "rotation": (P, 4), "scaling": (P, 3), "opacity": (P, 1), "f_dc": (P, 1, 3),
              "f_rest": (P, 15, 3), "base_color": (P, 12), "roughness": (P, 4), "incidents_dc": (P, 1, 3),
              "incidents_rest": (P, 3, 3), "visibility_dc": (P, 1, 1), "visibility_rest": (P, 3, 1)}
    for k, sh in shapes.items():
        v = rn(*sh)
        if k == "scaling":
            v = torch.log(0.004 + 0.03 * torch.rand(*sh, generator=g))     # straddles percent_dense * extent
        if k == "opacity":
            v = 2.5 * rn(*sh)                                                # some below min_opacity
        if k == "normal":
            v = 0.1 * v
        setattr(m, ATTR[k], torch.nn.Parameter(v.float()))
    m.max_radii2D = torch.zeros(P)
    return m, g


def group_params(m):
    return {grp["name"]: grp["params"][0] for grp in m.optimizer.param_groups}


def main():
    redirect_cuda_to_cpu()
    gm, dlm = import_reference()
    out = {}

    # ---- A: densification ---------------------------------------------------------------------------------------------
    P = 300
    m, g = make_model(gm, P, 11)
    m.training_setup(Opt())
    for step in range(2):                      # give Adam a state to carry through the surgery
        for k, p in group_params(m).items():
            p.grad = 0.1 * torch.randn(p.shape, generator=g)
        m.optimizer.step()
    gp = group_params(m)
    for k in GROUPS:
        out["A_in_" + k] = gp[k].detach().numpy().copy()
        st = m.optimizer.state[gp[k]]
        out["A_in_exp_avg_" + k] = st["exp_avg"].numpy().copy()
        out["A_in_exp_avg_sq_" + k] = st["exp_avg_sq"].numpy().copy()
    m.xyz_gradient_accum = 0.0006 * torch.rand(P, 1, generator=g)
    m.denom = torch.randint(0, 4, (P, 1), generator=g).float()           # zeros -> NaN / inf averages
    m.xyz_gradient_accum[m.denom == 0] = 0.0                              # never-visible surfels: 0 / 0
    m.normal_gradient_accum = torch.zeros(P, 1)
    m.weights_accum = torch.rand(P, 1, generator=g) * (torch.rand(P, 1, generator=g) > 0.1)
    m.max_radii2D = 40 * torch.rand(P, generator=g)
    for k in ("xyz_gradient_accum", "denom", "normal_gradient_accum", "weights_accum", "max_radii2D"):
        out["A_in_" + k] = getattr(m, k).numpy().copy()
    cfg = dict(max_grad=0.0002, min_opacity=0.05, extent=25.0, max_screen_size=20, max_grad_normal=0.1)   # 0.001 * extent straddles the scales
    zs = []
    orig_normal = torch.normal

    def recording_normal(mean, std, **k):
        z = torch.randn(mean.shape, generator=g)
        zs.append(z.clone())
        return mean + std * z
    torch.normal = recording_normal
    m.densify_and_prune(cfg["max_grad"], cfg["min_opacity"], cfg["extent"], cfg["max_screen_size"], cfg["max_grad_normal"])
    torch.normal = orig_normal
    assert len(zs) == 1
    out["A_z"] = zs[0].numpy()
    out["A_cfg"] = np.array([cfg["max_grad"], cfg["min_opacity"], cfg["extent"], cfg["max_screen_size"], cfg["max_grad_normal"],
                             Opt.percent_dense, 1e-5], np.float64)
    gp = group_params(m)
    for k in GROUPS:
        out["A_out_" + k] = gp[k].detach().numpy().copy()
        st = m.optimizer.state[gp[k]]
        out["A_out_exp_avg_" + k] = st["exp_avg"].numpy().copy()
        out["A_out_exp_avg_sq_" + k] = st["exp_avg_sq"].numpy().copy()
        assert getattr(m, ATTR[k]) is gp[k] or k in ("base_color", "roughness", "incidents_dc", "incidents_rest",
                                                     "visibility_dc", "visibility_rest")
    for k in ("xyz_gradient_accum", "denom", "normal_gradient_accum", "weights_accum", "max_radii2D"):
        out["A_out_" + k] = getattr(m, k).numpy().copy()
    print("A: %d -> %d surfels, split draws %s" % (P, gp["xyz"].shape[0], tuple(zs[0].shape)))

    # ---- B: NaN patches ------------------------------------------------------------------------------------------------
    m2, g2 = make_model(gm, 40, 12)
    m2.training_setup(Opt())
    for k, p in group_params(m2).items():
        gr = torch.randn(p.shape, generator=g2)
        gr[torch.rand(p.shape, generator=g2) < 0.15] = float("nan")
        p.grad = gr
        out["B_in_" + k] = gr.numpy().copy()
    m2.replace_nangrad_to_zero()
    for k, p in group_params(m2).items():
        out["B_out_" + k] = p.grad.numpy().copy()

    # ---- C: get_radiance_loss around its Slang call -------------------------------------------------------------------
    N, S = 64, 16
    m3, g3 = make_model(gm, N, 13)
    rn = lambda *s: torch.randn(*s, generator=g3)
    m3._incident_dirs = torch.nn.functional.normalize(rn(N, S, 3), dim=-1)
    m3._incident_areas = torch.full((N, S, 1), 2 * np.pi)
    vis = torch.rand(N, S, 1, generator=g3)
    vis[vis < 0.3] = 0.0
    vis[:8] = 1.0                                                        # fully visible surfels: all scores tie at +-0
    m3._visibility_tracing = vis
    m3._radiances = torch.rand(N, S, 3, generator=g3)
    m3._radiances[3, :, 1] = float("nan")
    m3._radiance_ratio = torch.tensor(0.7)
    light = dlm.DirectLightMap(16, 3.0)
    with torch.no_grad():
        light.env.copy_(3.0 * torch.rand(light.env.shape, generator=g3))
    cam = types.SimpleNamespace(camera_center=torch.tensor([0.4, -1.1, 2.3]))
    rec = {}
    fake_irr = torch.rand(N, 3, generator=g3)

    def render_irradiance_sample(Nn, Ss, max_idx, envmap, dirs, xyz, scaling, rotation, normal, albedo, roughness, metallic,
                                 opacity, features):
        rec.update(N=Nn, S=Ss, max_idx=max_idx.clone(), envmap=envmap.detach().clone(), normal=normal.detach().clone(),
                   albedo=albedo.detach().clone(), roughness=roughness.detach().clone(), metallic=metallic.detach().clone())
        return fake_irr
    m3.renderer = types.SimpleNamespace(render_irradiance_sample=render_irradiance_sample)
    loss = gm.GaussianModel.get_radiance_loss(m3, cam, light)
    for k in ("xyz", "rotation", "normal", "base_color", "roughness"):
        out["C_in_" + k] = getattr(m3, ATTR[k]).detach().numpy().copy()
    out.update(C_in_incident_dirs=m3._incident_dirs.numpy(), C_in_incident_areas=m3._incident_areas.numpy(),
               C_in_visibility=vis.numpy(), C_in_radiances=m3._radiances.numpy(), C_in_ratio=np.float32(0.7),
               C_in_env_param=light.env.detach().numpy().copy(), C_in_campos=cam.camera_center.numpy(),
               C_in_fake_irradiance=fake_irr.numpy(),
               C_out_max_idx=rec["max_idx"].numpy(), C_out_envmap=rec["envmap"].numpy(), C_out_normal=rec["normal"].numpy(),
               C_out_albedo=rec["albedo"].numpy(), C_out_roughness=rec["roughness"].numpy(),
               C_out_metallic=rec["metallic"].numpy(), C_out_geo_normal=m3.get_geo_normal.detach().numpy(),
               C_out_shading_normal=m3.get_shading_normal.detach().numpy(), C_out_loss=np.float64(float(loss)))
    print("C: loss %.6f, max_idx histogram head %s" % (float(loss), np.bincount(rec["max_idx"].numpy().ravel())[:6]))

    # ---- D: update_radiace chunking ------------------------------------------------------------------------------------
    for tag, Pn, Sn in (("a", 1000, 64), ("b", 1001, 64), ("c", 50, 24)):
        m4, g4 = make_model(gm, Pn, 14)
        calls = []

        def render_radiance_with_sampling_SH(xyz, dirs, cov_inv, sample_num, calls=calls):
            calls.append((int(xyz.shape[0]), dirs.clone()))
            n = xyz.shape[0]
            return (torch.zeros(n, sample_num, 3), torch.ones(n, sample_num, 1), torch.zeros(n, sample_num, 1, dtype=torch.int32),
                    torch.zeros(n, sample_num, 2))
        m4.renderer = types.SimpleNamespace(set_proxy_from_gaussian_model=lambda pc: None, build_bvh=lambda: None,
                                            render_radiance_with_sampling_SH=render_radiance_with_sampling_SH)
        gm.RayTracer = lambda *a, **k: None
        torch.manual_seed(77)
        m4.update_radiace(Sn)
        out["D_%s_chunks" % tag] = np.array([c[0] for c in calls], np.int64)
        out["D_%s_PS" % tag] = np.array([Pn, Sn], np.int64)
        if tag == "c":
            out["D_c_rotation"] = m4._rotation.detach().numpy().copy()
            out["D_c_dirs"] = torch.cat([c[1] for c in calls], 0).numpy()
            torch.manual_seed(77)
            out["D_c_rand"] = np.concatenate([torch.rand(c[0], 1).numpy() for c in calls], 0)
        print("D[%s]: chunks %s" % (tag, [c[0] for c in calls]))

    # ---- E: checkpoint interchange ------------------------------------------------------------------------------------
    # E1: a checkpoint written by the reference (train.py: torch.save((gaussians.capture(), iteration), path))
    m5, g5 = make_model(gm, 20, 15)
    m5.training_setup(Opt())
    for k, p in group_params(m5).items():
        p.grad = 0.1 * torch.randn(p.shape, generator=g5)
    m5.optimizer.step()
    m5._radiances = torch.rand(20, 8, 3, generator=g5)
    m5.max_radii2D = torch.rand(20, generator=g5)
    m5.weights_accum = torch.rand(20, 1, generator=g5)
    torch.save((m5.capture(), 4321), os.path.join(HERE, "ref_checkpoint.pth"))
    for k in GROUPS:
        out["E_" + k] = getattr(m5, ATTR[k]).detach().numpy().copy()
        out["E_exp_avg_" + k] = m5.optimizer.state[group_params(m5)[k]]["exp_avg"].numpy().copy()
    out["E_radiances"] = m5._radiances.numpy().copy()
    out["E_max_radii2D"] = m5.max_radii2D.numpy().copy()
    out["E_weights_accum"] = m5.weights_accum.numpy().copy()
    # E2: a checkpoint written by THIS repo (svgir_b200.io.save_checkpoint + FusedAdam.state_dict) restored by the reference
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "svg-ir_b200"))
    from svgir_b200 import io as sio, optim as sopt
    import tempfile
    names = {"xyz": "xyz", "normal": "normal", "f_dc": "shs_dc", "f_rest": "shs_rest", "scaling": "scaling", "rotation": "rotation",
             "opacity": "opacity", "base_color": "base_color", "roughness": "roughness", "incidents_dc": "incidents_dc",
             "incidents_rest": "incidents_rest", "visibility_dc": "visibility_dc", "visibility_rest": "visibility_rest"}
    g6 = torch.Generator().manual_seed(16)
    mine = {names[k]: torch.randn(tuple(getattr(m5, ATTR[k]).shape), generator=g6) for k in GROUPS}
    order = [grp["name"] for grp in m5.optimizer.param_groups]
    fa = sopt.FusedAdam.__new__(sopt.FusedAdam)     # its constructor insists on CUDA tensors; only state_dict() is needed here
    fa.param_groups = [{"name": k, "params": [mine[names[k]]], "lr": 1e-3 * (i + 1)} for i, k in enumerate(order)]
    fa.betas, fa.eps, fa.state = (0.9, 0.999), 1e-15, {}
    for k in order:
        fa.state[k] = {"exp_avg": torch.randn(mine[names[k]].shape, generator=g6), "exp_avg_sq": torch.rand(mine[names[k]].shape, generator=g6)}
    fa.step_count = 17
    model = dict(mine, active_sh_degree=3, max_radii2D=torch.rand(20, generator=g6), weights_accum=torch.rand(20, 1, generator=g6),
                 xyz_gradient_accum=torch.rand(20, 1, generator=g6), normal_gradient_accum=torch.zeros(20, 1),
                 denom=torch.ones(20, 1), opt_dict=fa.state_dict(), spatial_lr_scale=2.5,
                 radiances=torch.rand(20, 8, 3, generator=g6), radiance_ratio=torch.tensor(0.9))
    with tempfile.TemporaryDirectory() as td:
        sio.save_checkpoint(os.path.join(td, "chkpnt.pth"), model, 99)
        model_args, it = torch.load(os.path.join(td, "chkpnt.pth"), weights_only=False)
    m6 = gm.GaussianModel(3, render_type="render_relight")
    m6.restore(model_args, Opt(), is_training=True, restore_optimizer=True)
    ok = it == 99 and m6.spatial_lr_scale == 2.5
    for k in GROUPS:
        ok = ok and torch.equal(getattr(m6, ATTR[k]).detach(), mine[names[k]])
        stt = m6.optimizer.state[group_params(m6)[k]]
        ok = ok and torch.equal(stt["exp_avg"], fa.state[k]["exp_avg"]) and torch.equal(stt["exp_avg_sq"], fa.state[k]["exp_avg_sq"])
        ok = ok and float(stt["step"]) == 17.0
    ok = ok and torch.equal(m6._radiances, model["radiances"]) and torch.equal(m6.weights_accum, model["weights_accum"])
    for k, p in group_params(m6).items():          # the restored optimiser must be able to step
        p.grad = torch.zeros_like(p)
    m6.optimizer.step()
    assert ok, "the reference could not restore a checkpoint written by svgir_b200.io"
    out["E_reference_restored_our_checkpoint"] = np.int64(1)
    print("E: reference checkpoint written; the reference restored ours and stepped its optimiser")

    # ---- F: PLY ---------------------------------------------------------------------------------------------------------
    # `plyfile` is absent; what the reference hands to it (PlyElement.describe) and asks from it (PlyData.read) is the
    # whole interface, so the stand-ins below RECORD the structured array of save_ply and SERVE a file through this
    # repo's own reader to load_ply.
    rec_ply = {}

    class PlyElement:
        @staticmethod
        def describe(arr, name):
            rec_ply["names"], rec_ply["data"], rec_ply["element"] = list(arr.dtype.names), np.array(arr.tolist(), np.float32), name
            return (arr, name)

    class PlyData:
        def __init__(self, els=None):
            self.els = els

        def write(self, path):
            rec_ply["written"] = path

        @staticmethod
        def read(path):
            v = sio.read_ply_vertices(path)

            class El:
                properties = [types.SimpleNamespace(name=n) for n in v]

                def __getitem__(self, k):
                    return v[k]
            return types.SimpleNamespace(elements=[El()])
    gm.PlyElement, gm.PlyData = PlyElement, PlyData
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        # F1: stage-1 model (no PBR attributes): the reference's writer, recorded
        m7 = gm.GaussianModel(3, render_type="render")
        g7 = torch.Generator().manual_seed(17)
        Pn = 12
        for k, sh in {"xyz": (Pn, 3), "normal": (Pn, 12), "rotation": (Pn, 4), "scaling": (Pn, 3), "opacity": (Pn, 1),
                      "f_dc": (Pn, 1, 3), "f_rest": (Pn, 15, 3)}.items():
            setattr(m7, ATTR[k], torch.nn.Parameter(torch.randn(*sh, generator=g7)))
        m7.save_ply(os.path.join(td, "ref", "point_cloud.ply"))
        out["F1_names"] = np.array(rec_ply["names"])
        out["F1_data"] = rec_ply["data"]
        for k in ("xyz", "rotation", "scaling", "opacity", "f_dc", "f_rest"):
            out["F1_in_" + k] = getattr(m7, ATTR[k]).detach().numpy().copy()
        out["F1_geo_normal"] = m7.get_geo_normal.detach().numpy().copy()
        # F2: the same model written by THIS repo's writer and read back by the reference's load_ply
        ours = {"xyz": m7._xyz, "shs_dc": m7._shs_dc, "shs_rest": m7._shs_rest, "opacity": m7._opacity, "scaling": m7._scaling,
                "rotation": m7._rotation}
        sio.save_ply(os.path.join(td, "ours.ply"), ours, geo_normal=m7.get_geo_normal)
        m8 = gm.GaussianModel(3, render_type="render")
        m8.load_ply(os.path.join(td, "ours.ply"))
        ok = all(torch.equal(getattr(m8, ATTR[k]).detach(), getattr(m7, ATTR[k]).detach()) for k in
                 ("xyz", "rotation", "scaling", "opacity", "f_dc", "f_rest"))
        ok = ok and torch.equal(m8._normal.detach(), m7.get_geo_normal.detach().repeat(1, 4))
        assert ok, "the reference's load_ply did not reproduce the model from a file written by svgir_b200.io.save_ply"
        out["F2_reference_loaded_our_ply"] = np.int64(1)
        # F3: with PBR attributes the reference's writer names 12 roughness columns and supplies 4
        m9, _ = make_model(gm, Pn, 18)
        try:
            m9.save_ply(os.path.join(td, "ref_pbr", "point_cloud.ply"))
            out["F3_reference_pbr_save_raises"] = np.int64(0)
        except ValueError as e:
            out["F3_reference_pbr_save_raises"] = np.int64(1)
            print("F3: reference save_ply with PBR attributes raises: %s" % str(e)[:90])
        # F4: a PBR file written by this repo, read by the reference's load_ply: roughness comes back as the normal_* block
        pbr = {"xyz": m9._xyz, "shs_dc": m9._shs_dc, "shs_rest": m9._shs_rest, "opacity": m9._opacity, "scaling": m9._scaling,
               "rotation": m9._rotation, "base_color": m9._base_color, "normal": m9._normal, "roughness": m9._roughness,
               "incidents_dc": m9._incidents_dc, "incidents_rest": torch.randn(Pn, 15, 3), "visibility_dc": m9._visibility_dc,
               "visibility_rest": torch.randn(Pn, 15, 1)}
        sio.save_ply(os.path.join(td, "ours_pbr.ply"), pbr, geo_normal=m9.get_geo_normal)
        m10 = gm.GaussianModel(3, render_type="render_relight")
        m10.load_ply(os.path.join(td, "ours_pbr.ply"))
        mine = sio.load_ply(os.path.join(td, "ours_pbr.ply"), use_pbr=True, reference_roughness_quirk=True)
        names2 = {"xyz": "xyz", "rotation": "rotation", "scaling": "scaling", "opacity": "opacity", "f_dc": "shs_dc", "f_rest": "shs_rest",
                  "base_color": "base_color", "roughness": "roughness", "incidents_dc": "incidents_dc",
                  "incidents_rest": "incidents_rest", "visibility_dc": "visibility_dc", "visibility_rest": "visibility_rest"}
        ok = all(torch.equal(getattr(m10, ATTR[k]).detach(), mine[v]) for k, v in names2.items())
        assert ok, "io.load_ply(reference_roughness_quirk=True) differs from the reference's load_ply"
        assert torch.equal(m10._roughness.detach(), m9._normal.detach())        # the quirk itself
        out["F4_reference_pbr_load_equals_ours_with_quirk"] = np.int64(1)
    print("F: PLY names/data recorded (%d properties); reference loaded our files" % len(rec_ply["names"]))

    np.savez_compressed(os.path.join(HERE, "ref_model.npz"), **out)
    print("wrote ref_model.npz: %d arrays, %.1f KB" % (len(out), os.path.getsize(os.path.join(HERE, "ref_model.npz")) / 1e3))


if __name__ == "__main__":
    main()
