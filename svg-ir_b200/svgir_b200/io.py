"""Checkpoint and PLY formats of the reference's GaussianModel, so that models trained by the
reference load into this implementation (and back) for end-to-end comparisons.

* PLY: `save_ply` / `load_ply` (scene/gaussian_model.py:825-1003). The reference goes through the `plyfile` package
  (absent from this image): a `vertex` element of float32 properties, binary little-endian. The writer here emits that
  layout directly (header + packed float32 rows); the reader parses any PLY whose vertex element has scalar properties
  (ascii / binary little- or big-endian; float, double, integer types), by property NAME like the reference does.
  Property names and order follow construct_list_of_attributes (:825-853): x y z nx ny nz f_dc_* f_rest_* opacity
  scale_* rot_* [base_color_* normal_* roughness_* incidents_dc_* incidents_rest_* visibility_dc_* visibility_rest_*].
  Reference quirks, handled explicitly:
    - SH blocks are stored channel-major: `_shs_dc [P,1,3]` / `_shs_rest [P,15,3]` are transposed to [P,3,K] and flattened.
    - with PBR attributes the reference's OWN writer cannot work: it declares `roughness_i` for i < _normal.shape[1]
      (= 12) but supplies _roughness's 4 columns (:846-847, :868), so numpy rejects the row tuples. This writer declares
      as many `roughness_i` as there are columns. Its reader fills roughness from the `normal_*` columns (:947-953);
      `load_ply(..., reference_roughness_quirk=True)` reproduces that, the default reads `roughness_*`.
* Checkpoints: `torch.save((gaussians.capture(), iteration), path)` (train.py). `capture` / `restore` use the same
  positional layout (:195-268): 15 entries, + 8 with PBR attributes; `opt_dict` is an Adam state_dict
  (optim.FusedAdam.state_dict() has the same shape).
Host-side code only (numpy / torch tensors on any device); no CUDA kernels involved.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


def _np(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=np.float32)


def _channel_major(t) -> np.ndarray:
    """[P,K,C] -> [P, C*K] as `.transpose(1, 2).flatten(start_dim=1)` (gaussian_model.py:861-862)."""
    a = _np(t)
    return np.ascontiguousarray(np.transpose(a, (0, 2, 1))).reshape(a.shape[0], -1)


def attribute_names(model: Dict[str, np.ndarray]) -> List[str]:
    """construct_list_of_attributes (gaussian_model.py:825-853) for the tensors present in `model`."""
    l = ["x", "y", "z", "nx", "ny", "nz"]
    l += ["f_dc_%d" % i for i in range(model["shs_dc"].shape[1] * model["shs_dc"].shape[2])]
    l += ["f_rest_%d" % i for i in range(model["shs_rest"].shape[1] * model["shs_rest"].shape[2])]
    l.append("opacity")
    l += ["scale_%d" % i for i in range(model["scaling"].shape[1])]
    l += ["rot_%d" % i for i in range(model["rotation"].shape[1])]
    if "base_color" in model:
        l += ["base_color_%d" % i for i in range(model["base_color"].shape[1])]
        l += ["normal_%d" % i for i in range(model["normal"].shape[1])]
        l += ["roughness_%d" % i for i in range(model["roughness"].shape[1])]   # reference: _normal.shape[1] (see module doc)
        for k in ("incidents_dc", "incidents_rest", "visibility_dc", "visibility_rest"):
            l += ["%s_%d" % (k, i) for i in range(model[k].shape[1] * model[k].shape[2])]
    return l


def save_ply(path: str, model: Dict[str, torch.Tensor], geo_normal: Optional[torch.Tensor] = None) -> None:
    """model: raw (pre-activation) tensors under the reference's names without the underscore: xyz [P,3], shs_dc [P,1,3],
    shs_rest [P,K,3], opacity [P,1], scaling [P,3], rotation [P,4], and optionally the PBR block base_color [P,12],
    normal [P,12], roughness [P,4], incidents_dc/rest, visibility_dc/rest. geo_normal [P,3] fills nx ny nz (the
    reference writes get_geo_normal there; zeros if omitted)."""
    m = {k: _np(v) for k, v in model.items()}
    P = m["xyz"].shape[0]
    gn = _np(geo_normal) if geo_normal is not None else np.zeros((P, 3), np.float32)
    cols = [m["xyz"], gn, _channel_major(m["shs_dc"]), _channel_major(m["shs_rest"]), m["opacity"].reshape(P, -1), m["scaling"],
            m["rotation"]]
    if "base_color" in m:
        cols += [m["base_color"], m["normal"], m["roughness"], _channel_major(m["incidents_dc"]), _channel_major(m["incidents_rest"]),
                 _channel_major(m["visibility_dc"]), _channel_major(m["visibility_rest"])]
    names = attribute_names(m)
    data = np.ascontiguousarray(np.concatenate(cols, axis=1), dtype="<f4")
    if data.shape[1] != len(names):
        raise ValueError("save_ply: %d columns for %d property names" % (data.shape[1], len(names)))
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P
    header += "".join("property float %s\n" % n for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(data.tobytes())


def read_ply_vertices(path: str) -> Dict[str, np.ndarray]:
    """All scalar properties of the `vertex` element as float64/int arrays by name (what
    `plydata.elements[0][name]` gives the reference)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("%s: not a PLY file" % path)
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError("%s: unterminated PLY header" % path)
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(cur)
            elif tok[0] == "property":
                if tok[1] == "list":
                    cur["props"].append(("list", tok[2], tok[3], tok[4]))
                else:
                    cur["props"].append((tok[1], tok[2]))
            elif tok[0] == "end_header":
                break
        out = None
        for el in elements:
            scalar = all(p[0] != "list" for p in el["props"])
            if fmt == "ascii":
                rows = [f.readline().split() for _ in range(el["count"])]
                if el["name"] == "vertex":
                    if not scalar:
                        raise ValueError("vertex element with list properties is not supported")
                    arr = np.array(rows, dtype=np.float64).reshape(el["count"], len(el["props"]))
                    out = {p[1]: arr[:, i] for i, p in enumerate(el["props"])}
                continue
            end = "<" if fmt == "binary_little_endian" else ">"
            if not scalar:
                if el["name"] == "vertex":
                    raise ValueError("vertex element with list properties is not supported")
                for _ in range(el["count"]):        # skip faces etc.
                    for p in el["props"]:
                        if p[0] == "list":
                            ct = np.dtype(end + _PLY_TYPES[p[1]])
                            n = int(np.frombuffer(f.read(ct.itemsize), ct)[0])
                            f.read(n * np.dtype(_PLY_TYPES[p[2]]).itemsize)
                        else:
                            f.read(np.dtype(_PLY_TYPES[p[0]]).itemsize)
                continue
            dt = np.dtype([(p[1], end + _PLY_TYPES[p[0]]) for p in el["props"]])
            buf = f.read(dt.itemsize * el["count"])
            if el["name"] == "vertex":
                rec = np.frombuffer(buf, dtype=dt, count=el["count"])
                out = {n: np.asarray(rec[n]) for n in dt.names}
        if out is None:
            raise ValueError("%s: no vertex element" % path)
        return out


def _block(v: Dict[str, np.ndarray], prefix: str, exact: bool = False) -> np.ndarray:
    names = [n for n in v if (n.startswith(prefix + "_") if exact else n.startswith(prefix))]
    names = sorted(names, key=lambda x: int(x.split("_")[-1]))
    return np.stack([v[n] for n in names], axis=1).astype(np.float32) if names else np.zeros((len(v["x"]), 0), np.float32)


def load_ply(path: str, max_sh_degree: int = 3, use_pbr: bool = False, vertex_num: int = 4,
             reference_roughness_quirk: bool = False) -> Dict[str, torch.Tensor]:
    """load_ply (gaussian_model.py:883-1003): returns raw tensors by the reference's names (without the underscore).
    `normal` is the geometric normal repeated vertex_num times (:921) unless the PBR block carries `normal_*`."""
    v = read_ply_vertices(path)
    P = len(v["x"])
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    xyz = np.stack([v["x"], v["y"], v["z"]], 1)
    normal = np.stack([v["nx"], v["ny"], v["nz"]], 1)
    K = (max_sh_degree + 1) ** 2
    dc = np.stack([v["f_dc_0"], v["f_dc_1"], v["f_dc_2"]], 1).reshape(P, 3, 1)
    rest = _block(v, "f_rest_")
    if rest.shape[1] != 3 * K - 3:
        raise ValueError("load_ply: %d f_rest_* properties, expected %d for SH degree %d" % (rest.shape[1], 3 * K - 3, max_sh_degree))
    out = {"xyz": f32(xyz), "normal": f32(np.tile(normal, (1, vertex_num))), "rotation": f32(_block(v, "rot")),
           "scaling": f32(_block(v, "scale_")), "opacity": f32(np.asarray(v["opacity"])[:, None]),
           "shs_dc": f32(np.transpose(dc, (0, 2, 1))), "shs_rest": f32(np.transpose(rest.reshape(P, 3, K - 1), (0, 2, 1))),
           "geo_normal": f32(normal), "active_sh_degree": max_sh_degree}
    if use_pbr:
        out["base_color"] = f32(_block(v, "base_color"))
        sn = _block(v, "normal_", exact=False)
        out["normal"] = f32(sn) if sn.shape[1] else out["normal"]
        out["roughness"] = f32(sn) if reference_roughness_quirk else f32(_block(v, "roughness"))
        inc_dc = np.stack([v["incidents_dc_%d" % i] for i in range(3)], 1).reshape(P, 3, 1)
        inc_rest = _block(v, "incidents_rest_")
        out["incidents_dc"] = f32(np.transpose(inc_dc, (0, 2, 1)))
        out["incidents_rest"] = f32(np.transpose(inc_rest.reshape(P, 3, -1), (0, 2, 1)))
        vis_dc = np.asarray(v["visibility_dc_0"]).reshape(P, 1, 1)
        vis_rest = _block(v, "visibility_rest_")
        out["visibility_dc"] = f32(np.transpose(vis_dc, (0, 2, 1)))
        out["visibility_rest"] = f32(np.transpose(vis_rest.reshape(P, 1, -1), (0, 2, 1)))
    return out


# ---- checkpoints ---------------------------------------------------------------------------------------------------
_BASE = ("active_sh_degree", "xyz", "normal", "shs_dc", "shs_rest", "scaling", "rotation", "opacity", "max_radii2D",
         "weights_accum", "xyz_gradient_accum", "normal_gradient_accum", "denom", "opt_dict", "spatial_lr_scale")
_PBR = ("base_color", "roughness", "incidents_dc", "incidents_rest", "visibility_dc", "visibility_rest", "radiances",
        "radiance_ratio")


def capture(model: Dict[str, object]) -> list:
    """GaussianModel.capture() (gaussian_model.py:195-225): positional list, PBR entries appended when present."""
    out = [model[k] for k in _BASE]
    if "base_color" in model:
        out += [model.get(k) for k in _PBR]
    return out


def restore(model_args) -> Dict[str, object]:
    """GaussianModel.restore() / create_from_ckpt() (:227-268, :601-640): names the positional entries; accepts the
    16-entry stage-0 (vanilla 3DGS) layout too (`from_gs`)."""
    a = list(model_args)
    if len(a) == 16:      # (sh_degree, xyz, f_dc, f_rest, scaling, rotation, opacity, max_radii2D, grad accums..., denom, opt, scale)
        names = ("active_sh_degree", "xyz", "shs_dc", "shs_rest", "scaling", "rotation", "opacity", "max_radii2D",
                 "xyz_gradient_accum", "scale_gradient_accum", "rot_gradient_accum", "opac_gradient_accum", "denom", "opt_dict",
                 "spatial_lr_scale")
        return dict(zip(names, a[:15]))
    m = dict(zip(_BASE, a[:15]))
    if len(a) > 15:
        m.update(dict(zip(_PBR, a[15:23])))
    return m


def save_checkpoint(path: str, model: Dict[str, object], iteration: int) -> None:
    """train.py: torch.save((gaussians.capture(), iteration), path)."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    torch.save((capture(model), int(iteration)), path)


def load_checkpoint(path: str, map_location="cpu") -> Tuple[Dict[str, object], int]:
    model_args, it = torch.load(path, map_location=map_location, weights_only=False)
    return restore(model_args), int(it)


def _quat_third_column(q: torch.Tensor) -> torch.Tensor:
    """quaternion2rotmat(q)[..., 2] for normalised (r, x, y, z): the surfel's geometric normal (get_geo_normal)."""
    r, x, y, z = q.unbind(-1)
    return torch.stack([2 * (x * z + r * y), 2 * (y * z - r * x), 1 - 2 * (x * x + y * y)], -1)


def surfel_model_tensors(raw: Dict[str, torch.Tensor], base_color_scale: Optional[torch.Tensor] = None,
                         vertex_num: int = 4) -> Dict[str, torch.Tensor]:
    """Activated tensors the hot path consumes (GaussianModel's getters, gaussian_model.py:112-125, 270-351):
    opacity = sigmoid; scaling = nan_to_num(exp, 1e-6); rotation = nan_to_num(normalise, 1e-6); shs = cat(dc, rest);
    base_color = (0.77 sigmoid + 0.03) * base_color_scale (repeat-interleaved over the vertices); roughness =
    nan_to_num(0.9 sigmoid + 0.09, 1e-8); shading normals = normalise(geo_normal + offsets) with the offsets stored as
    `_normal [P, 3*vertex_num]` = [P,3,V] (:286-293)."""
    rot = torch.nan_to_num(torch.nn.functional.normalize(raw["rotation"], dim=-1), nan=1e-6)
    o = {"xyz": raw["xyz"], "opacity": torch.sigmoid(raw["opacity"]),
         "scaling": torch.nan_to_num(torch.exp(raw["scaling"]), nan=1e-6), "rotation": rot,
         "shs": torch.cat([raw["shs_dc"], raw["shs_rest"]], dim=1).contiguous(), "geo_normal": _quat_third_column(rot)}
    if "base_color" in raw:
        bc = torch.sigmoid(raw["base_color"]) * 0.77 + 0.03
        if base_color_scale is not None:
            bc = bc * base_color_scale[None, :].repeat_interleave(repeats=vertex_num, dim=1)
        o["base_color"] = bc
        o["roughness"] = torch.nan_to_num(torch.sigmoid(raw["roughness"]) * 0.9 + 0.09, nan=1e-8)
        off = raw["normal"].reshape(-1, 3, vertex_num).transpose(1, 2)
        o["shading_normal"] = torch.nn.functional.normalize(o["geo_normal"][:, None].repeat(1, vertex_num, 1) + off, dim=-1)
    return o
