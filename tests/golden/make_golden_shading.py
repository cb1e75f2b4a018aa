"""Generates tests/golden/ref_shading_*.npz by running the REFERENCE's own shading code
(/root/reference/gaussian_renderer/svgss.py: rendering_equation4, GGX_specular4;
scene/direct_light_map.py: DirectLightMap.direct_light; scene/envmap.py: EnvLight.direct_light)
on CPU with seeded inputs.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_shading.py

Missing third-party imports of the reference (plyfile, simple_knn, slangtorch, kornia, ...) are
replaced by empty stub modules; none of them is touched by the functions used here.
"""
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def import_reference():
    for name in ["plyfile", "simple_knn", "simple_knn._C", "custom_knn", "custom_knn._C", "slangtorch", "trimesh",
                 "pyexr", "nvdiffrast", "nvdiffrast.torch", "kornia", "kornia.filters", "matplotlib",
                 "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "imageio", "imageio.plugins",
                 "imageio.plugins.freeimage", "cv2", "tqdm", "lpips", "dearpygui", "dearpygui.dearpygui",
                 "svgss_rasterization", "pbgi", "pbgi.renderer", "submodules", "submodules.bvh"]:
        if name not in sys.modules:
            m = mock.MagicMock(name=name)
            m.__path__ = []
            m.__spec__ = None
            sys.modules[name] = m
    sys.path.insert(0, REF)
    # gaussian_renderer/rgss_rasterization.py:10-24 JIT-compiles the stage-1 extension on import;
    # the shading functions do not need it.
    import torch.utils.cpp_extension as cpp_ext
    cpp_ext.load = lambda *a, **k: mock.MagicMock(name="jit_ext")
    import gaussian_renderer.svgss as ref_svgss  # noqa
    import scene.direct_light_map as ref_dlm  # noqa
    return ref_svgss, ref_dlm


def make_inputs(N, Ns, seed, env_hw=(32, 64)):
    sys.path.insert(0, os.path.join(ROOT, "svg-ir_b200"))
    from svgir_b200 import scene
    cl = scene.make_surfels(N, seed=seed)
    m = scene.make_materials(cl, Ns, seed=seed + 1, env_hw=env_hw)
    cam = scene.look_at_camera(200, 200, 0)
    viewdirs = cam.campos[None] - cl.means3D
    viewdirs = viewdirs / np.linalg.norm(viewdirs, axis=-1, keepdims=True)
    m["viewdirs"] = viewdirs.astype(np.float32)
    m["view3x3"] = cam.viewmatrix[:3, :3].copy()
    return m


def main():
    ref_svgss, ref_dlm = import_reference()

    class Env:  # DirectLightMap without its hard-coded .cuda() constructor (direct_light_map.py:11-16)
        direct_light = ref_dlm.DirectLightMap.direct_light
        get_env = ref_dlm.DirectLightMap.get_env

        def __init__(self, p):
            self.env = p

    for tag, N, Ns, seed in (("train_small", 257, 64, 11), ("eval_small", 48, 384, 12), ("tiny", 5, 7, 13)):
        m = make_inputs(N, Ns, seed)
        t = {k: torch.tensor(v, requires_grad=k in ("base_color", "roughness", "shading_normals", "viewdirs",
                                                   "radiance", "env_param")) for k, v in m.items()}
        env = Env(t["env_param"])
        pbr, extra = ref_svgss.rendering_equation4(
            t["base_color"], t["roughness"], t["shading_normals"], t["viewdirs"], t["radiance"], env,
            visibility_precompute=t["visibility"], incident_dirs_precompute=t["incident_dirs"],
            incident_areas_precompute=t["incident_areas"])
        g = torch.Generator().manual_seed(seed)
        outs = dict(pbr=pbr, diffuse_light=extra["diffuse_light"], specular=extra["specular"],
                    direct=extra["direct"], indirect=extra["indirect"])
        cot = {k: torch.randn(v.shape, generator=g) for k, v in outs.items()}
        loss = sum((outs[k] * cot[k]).sum() for k in outs)
        loss.backward()
        save = {"in_" + k: v for k, v in m.items()}
        save.update({"out_" + k: v.detach().numpy() for k, v in outs.items()})
        save.update({"out_incident_lights": extra["incident_lights"].detach().numpy(),
                     "out_global_incident_lights": extra["global_incident_lights"].detach().numpy()})
        save.update({"cot_" + k: v.numpy() for k, v in cot.items()})
        # the same reference code evaluated in float64: the "truth" against which the fp32 rounding
        # noise of the reference itself (and of the CUDA kernels) is measured in the tests. The GGX
        # term NoH^2(a^2-1)+1 cancels catastrophically for small roughness, so the fp32 reference is
        # itself only ~1e-4..1e-3 relative on the specular peak of some surfels.
        t64 = {k: torch.tensor(v, dtype=torch.float64) for k, v in m.items()}
        pbr64, extra64 = ref_svgss.rendering_equation4(
            t64["base_color"], t64["roughness"], t64["shading_normals"], t64["viewdirs"], t64["radiance"],
            Env(t64["env_param"]), visibility_precompute=t64["visibility"],
            incident_dirs_precompute=t64["incident_dirs"], incident_areas_precompute=t64["incident_areas"])
        for k, v in dict(pbr=pbr64, diffuse_light=extra64["diffuse_light"], specular=extra64["specular"],
                         direct=extra64["direct"], indirect=extra64["indirect"]).items():
            save["out64_" + k] = v.detach().numpy()
        for k in ("base_color", "roughness", "shading_normals", "viewdirs", "radiance", "env_param"):
            save["grad_" + k] = t[k].grad.numpy()
        np.savez_compressed(os.path.join(HERE, f"ref_shading_{tag}.npz"), **save)
        print(tag, {k: float(np.abs(v).mean()) for k, v in save.items() if k.startswith("out_p") or k.startswith("grad_e")})

    # stand-alone env lookups (learnable map and the 32x64-resampled HDR path)
    rng = np.random.default_rng(5)
    d = rng.standard_normal((4000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d[:6] = np.array([[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0]], np.float32)
    envp = (3.0 * rng.uniform(0, 1, (1, 16, 32, 3))).astype(np.float32)
    out_l = ref_dlm.DirectLightMap.direct_light(Env(torch.tensor(envp)), torch.tensor(d)).numpy()
    hdr = (rng.uniform(0, 4, (24, 48, 3)) ** 2).astype(np.float32)
    # EnvLight.direct_light body (scene/envmap.py:54-72) needs only self.envmap / self.transform
    import importlib
    src = open(os.path.join(REF, "scene/envmap.py")).read()
    src = src.replace("imageio.plugins.freeimage.download()", "")
    mod = types.ModuleType("ref_envmap")
    sys.modules.setdefault("utils.graphics_utils", importlib.import_module("utils.graphics_utils"))
    exec(compile(src, "ref_envmap", "exec"), mod.__dict__)
    el = mod.EnvLight.__new__(mod.EnvLight)
    torch.nn.Module.__init__(el)
    el.envmap = torch.tensor(hdr)
    el.transform = None
    out_h = mod.EnvLight.direct_light(el, torch.tensor(d)).numpy()
    tr = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], np.float32)
    out_ht = mod.EnvLight.direct_light(el, torch.tensor(d), transform=torch.tensor(tr)).numpy()
    np.savez_compressed(os.path.join(HERE, "ref_envlight.npz"), dirs=d, env_param=envp, out_learnable=out_l,
                        hdr=hdr, out_hdr=out_h, transform=tr, out_hdr_transformed=out_ht)
    print("envlight", out_l.mean(), out_h.mean())


if __name__ == "__main__":
    main()
