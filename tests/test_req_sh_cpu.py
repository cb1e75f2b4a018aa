"""CPU: the SH render_equation restatement (oracle/render_equation_sh_oracle.py) against the golden
captured from the reference kernels on a B200 (tests/golden/make_golden_req_gpu.py), plus internal
consistency of the two backward variants."""
import os

import numpy as np
import pytest
import torch

from oracle import render_equation_sh_oracle as RO

GPATH = os.path.join(os.path.dirname(__file__), "golden", "ref_req_sh_small.npz")
ORDER = ("base_color", "roughness", "metallic", "normals", "viewdirs", "incidents_shs", "direct_shs", "visibility_shs")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_matches_reference_kernel_golden(tag):
    G = np.load(GPATH)
    P, Si, Sd, Sv, Ns = [int(x) for x in G[f"{tag}_meta"]]
    t = [torch.from_numpy(G[f"{tag}_in_{k}"]) for k in ORDER]
    o = RO.forward_complex(*t, Ns)
    for k in ("pbr", "incident_dirs", "incident_lights", "local_incident_lights", "global_incident_lights",
              "incident_visibility", "diffuse_light", "local_diffuse_light", "accum", "rgb_d", "rgb_s"):
        assert _rel(o[k].numpy(), G[f"{tag}_fc_{k}"]) < 2e-5, k
    for k in ("pbr", "incident_dirs", "diffuse_light"):
        assert _rel(o[k].numpy(), G[f"{tag}_fw_{k}"]) < 2e-5, k
    ot = RO.forward_complex(*t, Ns, rand_float=torch.from_numpy(G[f"{tag}_rand"]))
    for k in ("pbr", "incident_dirs", "diffuse_light"):
        assert _rel(ot[k].numpy(), G[f"{tag}_fwt_{k}"]) < 1e-4, k
    bw = RO.backward_legacy(*t, Ns, torch.from_numpy(G[f"{tag}_fw_incident_dirs"]), torch.from_numpy(G[f"{tag}_g_pbr"]),
                            torch.from_numpy(G[f"{tag}_g_dl"]))
    for k, v in bw.items():
        if k != "dL_ddirect_shs":
            assert _rel(v.numpy(), G[f"{tag}_bw_{k}"]) < 1e-4, k


def test_legacy_and_analytic_backward_agree_where_the_reference_is_right():
    """Gradients the reference's slips do not touch (base colour, metallic, roughness... all flow through f_d / f_s
    and the lighting) are identical in both variants when no clamp is active; dL_dnormals differs by design."""
    t = RO.make_inputs(64, 16, 16, 16, seed=3)
    # keep every SH-lit quantity strictly inside its clamp range
    t["incidents_shs"] = t["incidents_shs"] * 0.02
    t["incidents_shs"][:, 0] = 1.5
    t["direct_shs"] = t["direct_shs"] * 0.02
    t["visibility_shs"] = t["visibility_shs"] * 0.02
    ins = [t[k] for k in ORDER]
    Ns = 24
    dirs = RO.kernel_dirs(t["normals"], Ns)
    g = torch.Generator().manual_seed(1)
    gp, gd = torch.randn(64, 3, generator=g), torch.randn(64, 3, generator=g)
    a = RO.backward_legacy(*ins, Ns, dirs, gp, gd)
    b = RO.backward_analytic(*ins, Ns, dirs, gp, gd)
    for k in ("dL_dbase_color", "dL_dmetallic", "dL_droughness", "dL_dincidents_shs", "dL_ddirect_shs", "dL_dvisibility_shs"):
        assert _rel(a[k].numpy(), b[k].numpy()) < 1e-4, k
    assert _rel(a["dL_dnormals"].numpy(), b["dL_dnormals"].numpy()) > 1e-2  # the :406 overwrite drops the transport term
