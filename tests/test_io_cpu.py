"""Checkpoint / PLY formats of the reference's GaussianModel (svgir_b200.io; gaussian_model.py:195-268, 825-1003).
`plyfile` is absent, so the PLY checks are: byte-level layout of the writer (the header + packed float32 rows plyfile
produces for PlyElement.describe(float32 structured array)), the reader on little-/big-endian/ascii variants and with a
trailing face element, the reference's column conventions (channel-major SH blocks), and capture/restore round trips."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "svg-ir_b200"))


def _raw(P=37, pbr=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    m = {"xyz": r(P, 3), "shs_dc": r(P, 1, 3), "shs_rest": r(P, 15, 3), "opacity": r(P, 1), "scaling": r(P, 3), "rotation": r(P, 4)}
    if pbr:
        m.update({"base_color": r(P, 12), "normal": r(P, 12), "roughness": r(P, 4), "incidents_dc": r(P, 1, 3),
                  "incidents_rest": r(P, 15, 3), "visibility_dc": r(P, 1, 1), "visibility_rest": r(P, 15, 1)})
    return m


def test_ply_layout_and_roundtrip(tmp_path):
    from svgir_b200 import io
    m = _raw()
    gn = torch.randn(37, 3)
    path = str(tmp_path / "point_cloud" / "iteration_7" / "point_cloud.ply")
    io.save_ply(path, m, geo_normal=gn)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().split("\n")
    assert lines[0] == "ply" and lines[1] == "format binary_little_endian 1.0" and lines[2] == "element vertex 37"
    props = [l.split()[2] for l in lines if l.startswith("property float ")]
    assert props[:6] == ["x", "y", "z", "nx", "ny", "nz"] and props[6:9] == ["f_dc_0", "f_dc_1", "f_dc_2"]
    assert props[9] == "f_rest_0" and props[53] == "f_rest_44" and props[54] == "opacity"
    assert props[55:62] == ["scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert props[62] == "base_color_0" and props[74] == "normal_0" and props[86] == "roughness_0" and props[90] == "incidents_dc_0"
    assert len(props) == 62 + 12 + 12 + 4 + 3 + 45 + 1 + 15 and len(body) == 37 * 4 * len(props)
    rows = np.frombuffer(body, "<f4").reshape(37, len(props))
    # channel-major SH: f_rest_k = shs_rest[:, k % 15, k // 15]  (transpose(1, 2).flatten, gaussian_model.py:862)
    assert np.array_equal(rows[:, 9 + 17], m["shs_rest"][:, 2, 1].numpy())
    assert np.array_equal(rows[:, 3:6], gn.numpy())
    back = io.load_ply(path, max_sh_degree=3, use_pbr=True)
    for k in m:
        assert torch.equal(back[k], m[k]), k
    assert torch.equal(back["geo_normal"], gn)
    quirk = io.load_ply(path, use_pbr=True, reference_roughness_quirk=True)
    assert torch.equal(quirk["roughness"], m["normal"])       # gaussian_model.py:947-953 reads the normal_* columns


def test_ply_stage1_and_foreign_variants(tmp_path):
    from svgir_b200 import io
    m = _raw(P=11, pbr=False, seed=3)
    p1 = str(tmp_path / "a.ply")
    io.save_ply(p1, m)
    back = io.load_ply(p1)
    assert torch.equal(back["shs_rest"], m["shs_rest"]) and back["normal"].shape == (11, 12) and back["active_sh_degree"] == 3
    # the same content big-endian with doubles for xyz and a face element after the vertices, and as ascii
    v = io.read_ply_vertices(p1)
    names = list(v)
    dt = np.dtype([(n, ">f8" if n in "xyz" else ">f4") for n in names])
    rec = np.zeros(11, dt)
    for n in names:
        rec[n] = v[n]
    p2 = str(tmp_path / "b.ply")
    with open(p2, "wb") as f:
        f.write(("ply\nformat binary_big_endian 1.0\ncomment made by a test\nelement vertex 11\n" +
                 "".join("property %s %s\n" % ("double" if n in "xyz" else "float", n) for n in names) +
                 "element face 1\nproperty list uchar int vertex_indices\nend_header\n").encode())
        f.write(rec.tobytes())
        f.write(bytes([3]) + np.array([0, 1, 2], ">i4").tobytes())
    b2 = io.load_ply(p2)
    for k in m:
        assert torch.allclose(b2[k], m[k]), k
    p3 = str(tmp_path / "c.ply")
    with open(p3, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 11\n" + "".join("property float %s\n" % n for n in names) + "end_header\n")
        for i in range(11):
            f.write(" ".join(repr(float(v[n][i])) for n in names) + "\n")
    b3 = io.load_ply(p3)
    for k in m:
        assert torch.allclose(b3[k], m[k]), k


def test_checkpoint_capture_restore_roundtrip(tmp_path):
    from svgir_b200 import io
    m = _raw(P=9)
    model = dict(m)
    model.update({"active_sh_degree": 3, "max_radii2D": torch.zeros(9), "weights_accum": torch.rand(9, 1),
                  "xyz_gradient_accum": torch.rand(9, 1), "normal_gradient_accum": torch.zeros(9, 1), "denom": torch.ones(9, 1),
                  "opt_dict": {"state": {}, "param_groups": []}, "spatial_lr_scale": 2.5, "radiances": torch.rand(9, 64, 3),
                  "radiance_ratio": 1.0})
    cap = io.capture(model)
    assert len(cap) == 23 and cap[0] == 3 and cap[1] is model["xyz"] and cap[2] is model["normal"] and cap[14] == 2.5
    assert cap[15] is model["base_color"] and cap[21] is model["radiances"]
    path = str(tmp_path / "chkpnt50000.pth")
    io.save_checkpoint(path, model, 50000)
    back, it = io.load_checkpoint(path)
    assert it == 50000
    for k in ("xyz", "normal", "shs_rest", "roughness", "visibility_rest", "radiances", "weights_accum"):
        assert torch.equal(back[k], model[k]), k
    # the 16-entry stage-0 layout (create_from_ckpt from_gs, gaussian_model.py:603-620)
    gs = [3, m["xyz"], m["shs_dc"], m["shs_rest"], m["scaling"], m["rotation"], m["opacity"], torch.zeros(9), torch.zeros(9, 1),
          torch.zeros(9, 1), torch.zeros(9, 1), torch.zeros(9, 1), torch.ones(9, 1), {}, 1.0, None]
    r = io.restore(gs)
    assert r["xyz"] is m["xyz"] and r["shs_dc"] is m["shs_dc"] and "normal" not in r


def test_activations_follow_the_reference_getters():
    from svgir_b200 import io
    m = _raw(P=5)
    a = io.surfel_model_tensors(m, base_color_scale=torch.tensor([1.0, 0.5, 2.0]))
    assert torch.allclose(a["opacity"], torch.sigmoid(m["opacity"])) and torch.allclose(a["scaling"], torch.exp(m["scaling"]))
    assert torch.allclose(a["rotation"].norm(dim=-1), torch.ones(5)) and a["shs"].shape == (5, 16, 3)
    assert torch.allclose(a["base_color"][:, 4:8], (torch.sigmoid(m["base_color"][:, 4:8]) * 0.77 + 0.03) * 0.5)
    assert float(a["roughness"].min()) >= 0.09 and float(a["roughness"].max()) <= 0.99
    assert a["shading_normal"].shape == (5, 4, 3) and torch.allclose(a["shading_normal"].norm(dim=-1), torch.ones(5, 4), atol=1e-6)
    # geo normal = third column of the rotation matrix of the normalised quaternion
    q = torch.nn.functional.normalize(m["rotation"], dim=-1)
    r_, x, y, z = q.unbind(-1)
    assert torch.allclose(a["geo_normal"][:, 2], 1 - 2 * (x * x + y * y), atol=1e-6)
