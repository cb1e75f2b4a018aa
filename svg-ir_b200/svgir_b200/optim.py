"""Optimiser step and densification for the surfel model, host side.

`FusedAdam` replaces the reference's `torch.optim.Adam(l, lr=..., eps=1e-15)` with one parameter group per tensor
(scene/gaussian_model.py:737-773) plus `replace_nangrad_to_zero` (:775-795): ONE kernel launch per step over all
groups, gradients read in place (e.g. from the all-reduced dist.FlatGradBucket views).
`DensificationState` + `densify_and_prune` replace add_densification_stats / densify_and_clone / densify_and_split /
prune_points / cat_tensors_to_optimizer (:1005-1276): decisions, compaction and the split transform run as CUDA
kernels (csrc/optim.cu); PyTorch supplies memory, one cumsum per count array and the N(0,1) samples.
No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib

ADAM_MAX_GROUPS = 16


class AdamGroup(C.Structure):
    """svgir_adam_group (include/svgir_b200.h)."""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_longlong), ("lr", C.c_float), ("nan_fix", C.c_int32), ("nan_value", C.c_float),
                ("reserved_", C.c_int32)]


class DensifyCfg(C.Structure):
    """svgir_densify_cfg."""
    _fields_ = [("P", C.c_int32), ("grad_threshold", C.c_float), ("grad_normal_threshold", C.c_float),
                ("percent_dense", C.c_float), ("extent", C.c_float), ("min_opacity", C.c_float),
                ("weights_threshold", C.c_float), ("use_screen_size", C.c_int32)]


_bound = False


def _L():
    global _bound
    L = _lib.lib()
    if not _bound:
        vp, ll = C.c_void_p, C.c_longlong
        L.svgir_adam_step.argtypes = [C.POINTER(AdamGroup), C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, vp]
        L.svgir_adam_step.restype = C.c_int
        L.svgir_densify_stats.argtypes = [C.c_int] + [vp] * 8
        L.svgir_densify_stats.restype = C.c_int
        L.svgir_densify_decide.argtypes = [C.POINTER(DensifyCfg)] + [vp] * 11
        L.svgir_densify_decide.restype = C.c_int
        L.svgir_densify_index.argtypes = [C.c_int] + [vp] * 6 + [ll, ll, ll, vp, vp, vp]
        L.svgir_densify_index.restype = C.c_int
        L.svgir_gather_rows.argtypes = [ll, C.c_int, vp, vp, vp, C.c_int, C.c_float, vp, vp]
        L.svgir_gather_rows.restype = C.c_int
        L.svgir_densify_split.argtypes = [ll, ll] + [vp] * 9
        L.svgir_densify_split.restype = C.c_int
        _bound = True
    return L


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


# NaN patches of replace_nangrad_to_zero (gaussian_model.py:775-795): group name -> replacement
NAN_FIX = {"xyz": 0.0, "f_dc": 0.0, "f_rest": 0.0, "scaling": 1e-6, "rotation": 1e-6, "opacity": 0.0, "roughness": 1e-6,
           "base_color": 0.0, "normal": 0.0}


class FusedAdam:
    """Adam over named parameter groups, one tensor per group (the reference's layout). `groups`: list of dicts
    {"name", "params": [tensor], "lr"} exactly like the list training_setup builds; state (`exp_avg`, `exp_avg_sq`,
    `step`) mirrors torch.optim.Adam's so checkpoints interchange (state_dict / load_state_dict)."""

    def __init__(self, groups: Sequence[dict], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-15, nan_fix: Optional[dict] = None):
        if not 0 < len(groups) <= ADAM_MAX_GROUPS:
            raise ValueError("FusedAdam handles 1..%d groups per launch" % ADAM_MAX_GROUPS)
        self.param_groups = []
        for g in groups:
            (p,) = g["params"]
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdam: parameters must be contiguous CUDA float32 tensors (no CPU fallback)")
            self.param_groups.append({"name": g.get("name", str(len(self.param_groups))), "params": [p], "lr": float(g.get("lr", lr))})
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.nan_fix = dict(NAN_FIX if nan_fix is None else nan_fix)
        self.state: Dict[str, dict] = {}
        self.step_count = 0
        for g in self.param_groups:
            p = g["params"][0]
            self.state[g["name"]] = {"exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}

    def set_lr(self, name: str, lr: float):
        """update_learning_rate (gaussian_model.py:797-804) sets the scheduled xyz rate every iteration."""
        for g in self.param_groups:
            if g["name"] == name:
                g["lr"] = float(lr)
                return lr
        raise KeyError(name)

    def zero_grad(self, set_to_none: bool = False):
        for g in self.param_groups:
            p = g["params"][0]
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self):
        L = _L()
        live = [g for g in self.param_groups if g["params"][0].grad is not None and g["params"][0].numel() > 0]
        if not live:
            return
        self.step_count += 1
        arr = (AdamGroup * len(live))()
        keep = []
        dev = live[0]["params"][0].device
        for a, g in zip(arr, live):
            p = g["params"][0]
            gr = p.grad
            if gr.dtype != torch.float32 or not gr.is_contiguous():
                gr = gr.float().contiguous()
                keep.append(gr)
            st = self.state[g["name"]]
            if st["exp_avg"].shape != p.shape:
                raise RuntimeError("FusedAdam: state of group %r does not match its parameter (use replace_tensors after a "
                                   "densification)" % g["name"])
            a.param, a.grad, a.exp_avg, a.exp_avg_sq = p.data_ptr(), gr.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            a.numel, a.lr = p.numel(), g["lr"]
            fix = self.nan_fix.get(g["name"])
            a.nan_fix, a.nan_value = (0, 0.0) if fix is None else (1, float(fix))
        with torch.cuda.device(dev):
            _lib.check(L.svgir_adam_step(arr, len(live), self.betas[0], self.betas[1], self.eps, self.step_count, _stream(dev)),
                       "adam_step")

    def replace_tensors(self, new_params: Dict[str, torch.Tensor], new_state: Optional[Dict[str, dict]] = None):
        """After a densification: swap in the compacted parameters and moments (cat_tensors_to_optimizer /
        _prune_optimizer, gaussian_model.py:1019-1081)."""
        for g in self.param_groups:
            if g["name"] in new_params:
                g["params"][0] = new_params[g["name"]]
                if new_state is not None and g["name"] in new_state:
                    self.state[g["name"]] = new_state[g["name"]]
                else:
                    self.state[g["name"]] = {"exp_avg": torch.zeros_like(new_params[g["name"]]),
                                             "exp_avg_sq": torch.zeros_like(new_params[g["name"]])}

    def state_dict(self) -> dict:
        """Shaped like torch.optim.Adam.state_dict() (what capture() stores as opt_dict, gaussian_model.py:195-232)."""
        return {"state": {i: {"step": torch.tensor(float(self.step_count)), "exp_avg": self.state[g["name"]]["exp_avg"],
                              "exp_avg_sq": self.state[g["name"]]["exp_avg_sq"]} for i, g in enumerate(self.param_groups)},
                # every key torch.optim.Adam's step reads from a group (a loaded group replaces the live one wholesale)
                "param_groups": [{"name": g["name"], "lr": g["lr"], "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0,
                                  "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                                  "differentiable": False, "fused": None, "decoupled_weight_decay": False, "params": [i]}
                                 for i, g in enumerate(self.param_groups)]}

    def load_state_dict(self, sd: dict):
        names = {pg.get("name", str(i)): pg["params"][0] for i, pg in enumerate(sd["param_groups"])}
        for g in self.param_groups:
            idx = names.get(g["name"])
            if idx is None or idx not in sd["state"]:
                continue
            st = sd["state"][idx]
            p = g["params"][0]
            self.state[g["name"]] = {"exp_avg": st["exp_avg"].to(p.device, torch.float32).contiguous(),
                                     "exp_avg_sq": st["exp_avg_sq"].to(p.device, torch.float32).contiguous()}
            self.step_count = max(self.step_count, int(float(st.get("step", 0))))
        for pg in sd["param_groups"]:
            for g in self.param_groups:
                if g["name"] == pg.get("name"):
                    g["lr"] = float(pg["lr"])


class DensificationState:
    """weights_accum / xyz_gradient_accum / normal_gradient_accum / denom / max_radii2D of training_setup
    (gaussian_model.py:738-742)."""

    def __init__(self, P: int, device):
        f32 = dict(dtype=torch.float32, device=device)
        self.weights_accum = torch.zeros((P, 1), **f32)
        self.xyz_gradient_accum = torch.zeros((P, 1), **f32)
        self.normal_gradient_accum = torch.zeros((P, 1), **f32)
        self.denom = torch.zeros((P, 1), **f32)
        self.max_radii2D = torch.zeros((P,), **f32)

    def add(self, viewspace_grad: torch.Tensor, radii: torch.Tensor, weights: torch.Tensor):
        """add_densification_stats (gaussian_model.py:1270-1276) + the max_radii2D update of train.py, one kernel.
        viewspace_grad [P,3] = means2D.grad (FusedTrainStep.result['viewspace_grad']), radii int32 [P], weights [P,1]."""
        L = _L()
        P = int(radii.shape[0])
        dev = radii.device
        vg, w = viewspace_grad.contiguous(), weights.contiguous()
        with torch.cuda.device(dev):
            _lib.check(L.svgir_densify_stats(P, vg.data_ptr(), radii.contiguous().data_ptr(), w.data_ptr(),
                                             self.weights_accum.data_ptr(), self.xyz_gradient_accum.data_ptr(),
                                             self.denom.data_ptr(), self.max_radii2D.data_ptr(), _stream(dev)), "densify_stats")


@torch.no_grad()
def densify_and_prune(tensors: Dict[str, torch.Tensor], stats: DensificationState, max_grad: float, min_opacity: float,
                      extent: float, max_screen_size, max_grad_normal: float, percent_dense: float = 0.01,
                      weights_threshold: float = 1e-5, optimizer: Optional[FusedAdam] = None,
                      normal_samples: Optional[torch.Tensor] = None, generator=None):
    """densify_and_prune (gaussian_model.py:1224-1250): clone small / split large surfels with a large accumulated
    screen-space gradient, then prune transparent, never-hit and oversized ones -- as device-side compaction.
    `tensors`: every per-surfel tensor of the model by group name; "xyz" [P,3], "scaling" [P,3] (log-scales),
    "rotation" [P,4] and "opacity" [P,1] (logits) are required, everything else ([P,...]) is carried along.
    Returns (new tensors dict, new DensificationState, info). With `optimizer` its parameters and moments are replaced
    (new surfels start with zero moments). `normal_samples` [2 n_split, 3] ~ N(0,1) may be supplied (tests); the reference
    draws torch.normal(0, stds) -- here N(0,1) samples are scaled by the std inside the kernel."""
    L = _L()
    xyz, scaling, rotation, opacity = tensors["xyz"], tensors["scaling"], tensors["rotation"], tensors["opacity"]
    dev = xyz.device
    P = int(xyz.shape[0])
    i32 = dict(dtype=torch.int32, device=dev)
    cfg = DensifyCfg(P, float(max_grad), float(max_grad_normal), float(percent_dense), float(extent), float(min_opacity),
                     float(weights_threshold), 1 if max_screen_size else 0)
    flags = torch.empty((P,), dtype=torch.uint8, device=dev)
    keep, nclone, nsplit = torch.empty((P,), **i32), torch.empty((P,), **i32), torch.empty((P,), **i32)
    for t in (scaling, rotation, opacity, xyz):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("densify_and_prune: contiguous float32 tensors expected")
    with torch.cuda.device(dev):
        _lib.check(L.svgir_densify_decide(C.byref(cfg), stats.xyz_gradient_accum.data_ptr(), stats.normal_gradient_accum.data_ptr(),
                                          stats.denom.data_ptr(), scaling.data_ptr(), opacity.data_ptr(), stats.weights_accum.data_ptr(),
                                          flags.data_ptr(), keep.data_ptr(), nclone.data_ptr(), nsplit.data_ptr(), _stream(dev)),
                   "densify_decide")
        ks, cs, ss = keep.cumsum(0), nclone.cumsum(0), nsplit.cumsum(0)      # int64 inclusive scans
        nA, nB, nC = (int(v) for v in torch.stack([ks[-1], cs[-1], ss[-1]]).tolist()) if P else (0, 0, 0)
        n_new = nA + nB + 2 * nC
        src = torch.empty((max(n_new, 1),), **i32)
        kind = torch.empty((max(n_new, 1),), dtype=torch.uint8, device=dev)
        _lib.check(L.svgir_densify_index(P, keep.data_ptr(), nclone.data_ptr(), nsplit.data_ptr(), ks.data_ptr(), cs.data_ptr(),
                                         ss.data_ptr(), nA, nB, nC, src.data_ptr(), kind.data_ptr(), _stream(dev)), "densify_index")

        def gather(t, zero_new=False, value=0.0):
            K = int(t.numel() // max(P, 1))
            out = torch.empty((n_new,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev)
            _lib.check(L.svgir_gather_rows(n_new, K, t.data_ptr(), src.data_ptr(), kind.data_ptr(), 1 if zero_new else 0,
                                           float(value), out.data_ptr(), _stream(dev)), "gather_rows")
            return out

        out = {}
        for name, t in tensors.items():
            if t.shape[0] != P:
                raise ValueError("densify_and_prune: tensor %r is not per-surfel" % name)
            tc = t.detach()
            if tc.dtype != torch.float32 or not tc.is_contiguous():
                tc = tc.float().contiguous()
            out[name] = gather(tc)
        if nC > 0:
            if normal_samples is None:
                normal_samples = torch.randn((2 * nC, 3), dtype=torch.float32, device=dev, generator=generator)
            if tuple(normal_samples.shape) != (2 * nC, 3):
                raise ValueError("normal_samples must be [2 n_split, 3] = [%d, 3]" % (2 * nC))
            ns = normal_samples.to(dev, torch.float32).contiguous()
            _lib.check(L.svgir_densify_split(n_new, nA + nB, src.data_ptr(), kind.data_ptr(), xyz.data_ptr(), scaling.data_ptr(),
                                             rotation.data_ptr(), ns.data_ptr(), out["xyz"].data_ptr(), out["scaling"].data_ptr(),
                                             _stream(dev)), "densify_split")
        if optimizer is not None:
            new_state = {}
            for g in optimizer.param_groups:
                if g["name"] in tensors:
                    st = optimizer.state[g["name"]]
                    new_state[g["name"]] = {"exp_avg": gather(st["exp_avg"], zero_new=True),
                                            "exp_avg_sq": gather(st["exp_avg_sq"], zero_new=True)}
            for name in list(out):
                if any(g["name"] == name for g in optimizer.param_groups):
                    out[name].requires_grad_(True)
            optimizer.replace_tensors({n: out[n] for n in new_state}, new_state)
    # densification_postfix resets every accumulator and the final prune zeroes weights_accum (:1119-1123, :1247)
    new_stats = DensificationState(n_new, dev)
    info = {"n_before": P, "n_after": n_new, "kept": nA, "cloned": nB, "split": nC, "flags": flags, "src": src[:n_new],
            "kind": kind[:n_new]}
    return out, new_stats, info
