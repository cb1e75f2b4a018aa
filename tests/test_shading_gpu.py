"""GPU parity of the fused render_equation kernels (csrc/shading.cu) through the C ABI against
(a) golden vectors produced by the reference's own Python code and (b) the torch restatement."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ("base_color", "roughness", "shading_normals", "viewdirs", "radiance", "env_param")


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("tag", ["tiny", "train_small", "eval_small"])
def test_shade_matches_reference_golden(tag):
    from svgir_b200 import shading
    g = dict(np.load(os.path.join(GOLD, f"ref_shading_{tag}.npz")))
    t = {k[3:]: torch.tensor(v).cuda() for k, v in g.items() if k.startswith("in_")}
    for k in NAMES:
        t[k].requires_grad_(True)
    r = shading.shade_surfels(t["base_color"], t["roughness"], t["shading_normals"], t["viewdirs"], t["radiance"],
                              (t["env_param"], shading.MODE_LEARNABLE), t["visibility"], t["incident_dirs"],
                              t["incident_areas"])
    # forward. North-star tolerance: 1e-5 absolute on the O(1) G-buffer inputs. Two fp32 effects make
    # an element-wise 1e-5 check against the reference's own fp32 output meaningless for a few
    # surfels: (i) the GGX term NoH^2(a^2-1)+1 cancels catastrophically for small roughness, so the
    # reference evaluated in fp32 is itself only 1e-4..3e-4 relative on some specular peaks (measured
    # against the same reference code run in float64, stored as out64_* by make_golden_shading.py);
    # (ii) acos(d.z) in the lat-long lookup (direct_light_map.py:76) is ill-conditioned at the poles.
    # So the kernel is held to the float64 truth with the fp32 reference's own error as the yardstick, in
    # units of the north-star tolerance e = |x - x64| / (1e-5 + 3e-5 |x64|): relative L2 error, the 99th
    # percentile of e and max e all <= 3x the reference's (floors: 2e-6, 1, 3), and every element within
    # 1e-3 rel of the fp32 reference.
    for k in ("pbr", "diffuse_light", "specular", "direct", "indirect"):
        a, b, b64 = r[k].detach().cpu().numpy(), g["out_" + k], g["out64_" + k]
        ref_err = _rel(b, b64)
        assert _rel(a, b64) <= max(3.0 * ref_err, 2e-6), (k, _rel(a, b64), ref_err)
        tol = 1e-5 + 3e-5 * np.abs(b64)
        e, e_ref = np.abs(a - b64) / tol, np.abs(b - b64) / tol
        assert np.quantile(e, 0.99) <= max(3.0 * np.quantile(e_ref, 0.99), 1.0), (k, np.quantile(e, 0.99), np.quantile(e_ref, 0.99))
        assert e.max() <= max(3.0 * e_ref.max(), 3.0), (k, e.max(), e_ref.max())
        np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-4, err_msg=k)
    np.testing.assert_allclose(r["mean_incident_lights"].detach().cpu().numpy(), g["out_incident_lights"].mean(-2),
                               rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(r["mean_visibility"].detach().cpu().numpy(), g["in_visibility"].mean(-2), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(r["mean_local_lights"].detach().cpu().numpy(), g["in_radiance"].mean(-2), rtol=1e-5, atol=1e-6)
    loss = sum((r[k] * torch.tensor(g["cot_" + k]).cuda()).sum() for k in ("pbr", "diffuse_light", "specular", "direct", "indirect"))
    loss.backward()
    for k in NAMES:  # gradients: 1e-3 relative (north star)
        err = _rel(t[k].grad.cpu().numpy(), g["grad_" + k])
        assert err < 1e-3, (k, err)


def test_shade_compat_wrapper_and_lazy_lights():
    from svgir_b200 import shading
    g = dict(np.load(os.path.join(GOLD, "ref_shading_tiny.npz")))
    t = {k[3:]: torch.tensor(v).cuda() for k, v in g.items() if k.startswith("in_")}
    pbr, extra = shading.rendering_equation4(
        t["base_color"], t["roughness"], t["shading_normals"], t["viewdirs"], t["radiance"],
        (t["env_param"], shading.MODE_LEARNABLE), visibility_precompute=t["visibility"],
        incident_dirs_precompute=t["incident_dirs"], incident_areas_precompute=t["incident_areas"])
    np.testing.assert_allclose(pbr.cpu().numpy(), g["out_pbr"], rtol=3e-5, atol=1e-5)
    for k in ("incident_dirs", "incident_lights", "local_incident_lights", "global_incident_lights",
              "incident_visibility", "diffuse_light", "specular", "direct", "indirect"):
        assert k in extra
    np.testing.assert_allclose(extra["incident_lights"].cpu().numpy(), g["out_incident_lights"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(extra["global_incident_lights"].cpu().numpy(), g["out_global_incident_lights"], rtol=2e-5, atol=1e-5)


def test_direct_light_learnable_and_hdr():
    from svgir_b200 import shading
    g = dict(np.load(os.path.join(GOLD, "ref_envlight.npz")))
    d = torch.tensor(g["dirs"]).cuda()
    envp = torch.tensor(g["env_param"]).cuda().requires_grad_(True)
    out = shading.direct_light((envp, shading.MODE_LEARNABLE), d)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["out_learnable"], rtol=2e-5, atol=2e-5)
    # gradient w.r.t. the env parameter against torch autograd of the restatement
    from oracle import shading_oracle as SO
    cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    (out * cot.cuda()).sum().backward()
    pe = torch.tensor(g["env_param"], requires_grad=True)
    (SO.direct_light_learnable(pe, torch.tensor(g["dirs"])) * cot).sum().backward()
    assert _rel(envp.grad.cpu().numpy(), pe.grad.numpy()) < 1e-4

    class HDR:  # duck-typed EnvLight (scene/envmap.py:26-34)
        envmap = torch.tensor(g["hdr"]).cuda()
        transform = None
    out = shading.direct_light(HDR(), d)
    np.testing.assert_allclose(out.cpu().numpy(), g["out_hdr"], rtol=2e-5, atol=2e-5)
    out = shading.direct_light(HDR(), d, transform=torch.tensor(g["transform"]).cuda())
    np.testing.assert_allclose(out.cpu().numpy(), g["out_hdr_transformed"], rtol=2e-5, atol=2e-5)


def test_shade_metallic_against_restatement():
    """Optional per-vertex metallic: compare with a torch restatement using
    f_d=(1-m) base/pi, F0=0.04(1-m)+base*m (legacy render_equation.cu:55-190)."""
    import math
    from svgir_b200 import shading
    from oracle import shading_oracle as SO
    g = dict(np.load(os.path.join(GOLD, "ref_shading_tiny.npz")))
    t = {k[3:]: torch.tensor(v) for k, v in g.items() if k.startswith("in_")}
    torch.manual_seed(0)
    met = torch.rand(t["roughness"].shape)
    # restatement
    Lg = SO.direct_light_learnable(t["env_param"], t["incident_dirs"]).clamp(0, 64) * t["visibility"]
    L = t["radiance"] + Lg
    ndi = (t["shading_normals"][:, None] * t["incident_dirs"][:, :, None]).sum(-1, keepdim=True).clamp(min=0)
    D = SO.ggx_specular4(t["shading_normals"], t["viewdirs"], t["incident_dirs"], t["roughness"], fresnel=0.0)  # (p)*D
    D1 = SO.ggx_specular4(t["shading_normals"], t["viewdirs"], t["incident_dirs"], t["roughness"], fresnel=1.0)  # D
    base = t["base_color"].reshape(-1, 3, 4).transpose(1, 2)  # [n,4,3]
    m = met[:, :, None]
    F0 = 0.04 * (1 - m) + base * m
    fs = F0[:, None] * D1 + (1 - F0[:, None]) * D           # [n,Ns,4,3]
    fd = ((1 - m) * base / math.pi)[:, None]
    T = L[:, :, None] * t["incident_areas"][:, :, None] * ndi
    pbr = ((fd + fs) * T).mean(1).transpose(1, 2).reshape(-1, 12)
    tc = {k: v.cuda() for k, v in t.items()}
    r = shading.shade_surfels(tc["base_color"], tc["roughness"], tc["shading_normals"], tc["viewdirs"], tc["radiance"],
                              (tc["env_param"], shading.MODE_LEARNABLE), tc["visibility"], tc["incident_dirs"],
                              tc["incident_areas"], metallic=met.cuda())
    np.testing.assert_allclose(r["pbr"].cpu().numpy(), pbr.numpy(), rtol=5e-5, atol=1e-5)


@pytest.mark.parametrize("is_training", [True, False])
def test_shade_and_pack_matches_torch_packing(is_training):
    """Fused shading + render_view packing (svgss.py:141-166) against the un-fused kernel followed by the
    reference's torch cat / matmul packing, values and gradients."""
    from svgir_b200 import shading
    g = dict(np.load(os.path.join(GOLD, "ref_shading_train_small.npz")))
    names = ("base_color", "roughness", "shading_normals", "viewdirs", "env_param")
    view = torch.tensor(g["in_view3x3"]).cuda()

    def inputs():
        t = {k[3:]: torch.tensor(v).cuda() for k, v in g.items() if k.startswith("in_")}
        for k in names:
            t[k].requires_grad_(True)
        return t

    t = inputs()
    r = shading.shade_surfels(t["base_color"], t["roughness"], t["shading_normals"], t["viewdirs"], t["radiance"],
                              (t["env_param"], shading.MODE_LEARNABLE), t["visibility"], t["incident_dirs"],
                              t["incident_areas"])
    nview = (t["shading_normals"] @ view).transpose(1, 2).reshape(t["shading_normals"].shape[0], -1)
    if is_training:
        f_ref = torch.cat([r["mean_visibility"], r["mean_local_lights"]], -1)
        vf_ref = torch.cat([r["pbr"], t["base_color"], nview, t["roughness"], r["diffuse_light"]], -1)
    else:
        f_ref = torch.cat([r["mean_incident_lights"], r["mean_local_lights"], r["mean_visibility"]], -1)
        vf_ref = torch.cat([r["pbr"], t["base_color"], nview, t["roughness"], r["direct"], r["indirect"]], -1)
    gen = torch.Generator().manual_seed(5)
    cf, cvf = torch.randn(f_ref.shape, generator=gen).cuda(), torch.randn(vf_ref.shape, generator=gen).cuda()
    ((f_ref * cf).sum() + (vf_ref * cvf).sum()).backward()

    t2 = inputs()
    f, vf = shading.shade_and_pack(t2["base_color"], t2["roughness"], t2["shading_normals"], t2["viewdirs"], t2["radiance"],
                                   (t2["env_param"], shading.MODE_LEARNABLE), t2["visibility"], t2["incident_dirs"],
                                   t2["incident_areas"], view, is_training=is_training)
    np.testing.assert_allclose(f.detach().cpu().numpy(), f_ref.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(vf.detach().cpu().numpy(), vf_ref.detach().cpu().numpy(), rtol=2e-4, atol=1e-5)
    ((f * cf).sum() + (vf * cvf).sum()).backward()
    for k in names:
        assert _rel(t2[k].grad.cpu().numpy(), t[k].grad.cpu().numpy()) < 1e-4, k


@pytest.mark.parametrize("N,Ns,chunk", [(300_000, 64, 50_000), (100_000, 384, 12_500)])
def test_shade_full_size_differential(N, Ns, chunk):
    """the goldens pin the kernels at N <= 257. Here the fused kernels run at the bench sizes (C3-train:
    300k surfels x 64 samples; one C3-eval chunk: 100k x 384) against the torch graph of the reference's
    rendering_equation4 + direct_light (oracle/shading_oracle.py, itself pinned to goldens made by the reference's own
    Python) evaluated on the same GPU in float32 AND float64, chunk by chunk. Values: held to the float64 evaluation
    with the float32 graph's own error as the yardstick (same rule as test_shade_matches_reference_golden: relative L2,
    99th percentile and max of |x - x64| / (1e-5 + 3e-5 |x64|) at most 3x the fp32 graph's, floors 2e-6 / 1 / 3).
    Gradients: 1e-3 relative L2 against the fp32 graph's autograd (north star)."""
    from svgir_b200 import scene, shading
    from oracle import shading_oracle as SO
    dev = torch.device("cuda:0")
    cloud = scene.make_surfels(N, seed=77)
    mats = scene.make_materials_torch(cloud, Ns, 78, dev)
    campos = torch.tensor([0.3, -3.6, 1.7], device=dev)
    viewdirs = torch.nn.functional.normalize(campos - torch.from_numpy(cloud.means3D).to(dev), dim=-1)
    gen = torch.Generator(dev).manual_seed(79)
    cots = {k: torch.randn((N, 12), device=dev, generator=gen) / N for k in ("pbr", "diffuse_light", "specular")}
    names = ("base_color", "roughness", "shading_normals", "env_param")
    t = {k: mats[k].clone().requires_grad_(True) for k in names}
    vd = viewdirs.clone().requires_grad_(True)
    r = shading.shade_surfels(t["base_color"], t["roughness"], t["shading_normals"], vd, mats["radiance"],
                              (t["env_param"], shading.MODE_LEARNABLE), mats["visibility"], mats["incident_dirs"],
                              mats["incident_areas"])
    sum((r[k] * cots[k]).sum() for k in cots).backward()
    ours = {k: r[k].detach() for k in cots}
    ours_g = {k: t[k].grad.clone() for k in names}
    ours_g["viewdirs"] = vd.grad.clone()

    ref32 = {k: torch.empty((N, 12), device=dev) for k in cots}
    ref64 = {k: torch.empty((N, 12), device=dev, dtype=torch.float64) for k in cots}
    ref_g = {"base_color": torch.empty_like(mats["base_color"]), "roughness": torch.empty_like(mats["roughness"]),
             "shading_normals": torch.empty_like(mats["shading_normals"]), "viewdirs": torch.empty_like(viewdirs),
             "env_param": torch.zeros_like(mats["env_param"])}
    for lo in range(0, N, chunk):
        sl = slice(lo, min(lo + chunk, N))
        for dt in (torch.float32, torch.float64):
            c = {k: mats[k][sl].to(dt).detach().requires_grad_(dt == torch.float32)
                 for k in ("base_color", "roughness", "shading_normals")}
            envp = mats["env_param"].to(dt).detach().requires_grad_(dt == torch.float32)
            v = viewdirs[sl].to(dt).detach().requires_grad_(dt == torch.float32)
            with torch.set_grad_enabled(dt == torch.float32):
                pbr, extra = SO.rendering_equation4(
                    c["base_color"], c["roughness"], c["shading_normals"], v, mats["radiance"][sl].to(dt),
                    lambda d: SO.direct_light_learnable(envp, d), mats["visibility"][sl].to(dt),
                    mats["incident_dirs"][sl].to(dt), mats["incident_areas"][sl].to(dt))
            out = {"pbr": pbr, "diffuse_light": extra["diffuse_light"], "specular": extra["specular"]}
            if dt == torch.float32:
                sum((out[k] * cots[k][sl]).sum() for k in cots).backward()
                for k in cots:
                    ref32[k][sl] = out[k].detach()
                for k in ("base_color", "roughness", "shading_normals"):
                    ref_g[k][sl] = c[k].grad
                ref_g["viewdirs"][sl] = v.grad
                ref_g["env_param"] += envp.grad
            else:
                for k in cots:
                    ref64[k][sl] = out[k]
            del c, envp, v, pbr, extra, out
        torch.cuda.empty_cache()

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    for k in cots:
        a, b, b64 = ours[k].double(), ref32[k].double(), ref64[k]
        ref_err = rel(b, b64)
        assert rel(a, b64) <= max(3.0 * ref_err, 2e-6), (k, rel(a, b64), ref_err)
        tol = 1e-5 + 3e-5 * b64.abs()
        e, e_ref = ((a - b64).abs() / tol).flatten(), ((b - b64).abs() / tol).flatten()
        q, q_ref = float(torch.quantile(e[:: max(1, e.numel() // 4_000_000)], 0.99)), float(torch.quantile(e_ref[:: max(1, e.numel() // 4_000_000)], 0.99))
        assert q <= max(3.0 * q_ref, 1.0), (k, q, q_ref)
        assert float(e.max()) <= max(3.0 * float(e_ref.max()), 3.0), (k, float(e.max()), float(e_ref.max()))
    for k, gref in ref_g.items():
        assert rel(ours_g[k], gref) < 1e-3, (k, rel(ours_g[k], gref))


def test_env_tap_cache_is_bit_identical_and_follows_in_place_updates():
    """svgir_env_taps + svgir_shade_in.env_taps: the cached texel corner / weights are the kernels' own env_coords
    results, so outputs and gradients are bit-equal with and without the cache; re-sampling the directions in place
    is picked up through the tensor version (the buffer is refreshed in place, as graphs that baked it need)."""
    from svgir_b200 import shading
    g = dict(np.load(os.path.join(GOLD, "ref_shading_train_small.npz")))
    names = ("base_color", "roughness", "shading_normals", "viewdirs", "env_param")

    def run(t, cache):
        shading.ENV_TAP_CACHE = cache
        for k in names:
            t[k].grad = None
        r = shading.shade_surfels(t["base_color"], t["roughness"], t["shading_normals"], t["viewdirs"], t["radiance"],
                                  (t["env_param"], shading.MODE_LEARNABLE), t["visibility"], t["incident_dirs"],
                                  t["incident_areas"])
        (r["pbr"].sum() + 0.5 * r["specular"].sum() + r["mean_incident_lights"].sum()).backward()
        shading.ENV_TAP_CACHE = True
        return [r[k].detach().clone() for k in ("pbr", "diffuse_light", "specular")] + [t[k].grad.clone() for k in names]

    t = {k[3:]: torch.tensor(v).cuda() for k, v in g.items() if k.startswith("in_")}
    for k in names:
        t[k].requires_grad_(True)
    shading._TAP_CACHE.clear()
    a = run(t, False)
    assert not shading._TAP_CACHE
    b = run(t, True)
    assert len(shading._TAP_CACHE) == 1
    for x, y in zip(a[:3] + a[3:4] + a[5:7], b[:3] + b[3:4] + b[5:7]):     # forward + deterministic gradients: bit-equal
        assert torch.equal(x, y)
    for x, y in zip(a, b):                                                   # atomically accumulated ones: order only
        assert _rel(y.cpu().numpy(), x.cpu().numpy()) < 1e-6
    taps = next(iter(shading._TAP_CACHE.values()))[0]
    ptr = taps.data_ptr()
    want = shading.env_taps(t["incident_dirs"], t["env_param"].shape[-3], t["env_param"].shape[-2])
    assert torch.equal(taps, want)
    # the directions are re-sampled in place (update_radiace): same buffer, new content
    with torch.no_grad():
        d = t["incident_dirs"]
        d.copy_(torch.nn.functional.normalize(d + 0.3 * torch.roll(d, 1, dims=1), dim=-1))
    c_off = run(t, False)
    c_on = run(t, True)
    assert len(shading._TAP_CACHE) == 1 and next(iter(shading._TAP_CACHE.values()))[0].data_ptr() == ptr
    assert torch.equal(c_off[0], c_on[0]) and not torch.equal(c_on[0], b[0])
    # tap words decode to the lookup's texel corner: an in-range direction lands inside the map
    He, We = int(t["env_param"].shape[-3]), int(t["env_param"].shape[-2])
    w0 = taps.reshape(-1, 3)[:, 0].contiguous().view(torch.int32)
    x0 = ((w0 & 0xffff) ^ 0x8000) - 0x8000
    y0 = w0 >> 16
    assert int(x0.min()) >= -1 and int(x0.max()) <= We - 1 and int(y0.min()) >= -1 and int(y0.max()) <= He - 1
