"""Generates tests/golden/ref_req_sh_small.npz by running the REFERENCE's SH render_equation kernels
(rgss-rasterization/render_equation.cu compiled unmodified into oracle/_ref/libreq_ref.so, see oracle/Makefile)
on a GPU. Run on the GPU box:
    python tests/golden/make_golden_req_gpu.py   (writes gpurun_out/ref_req_sh_small.npz; copy it to tests/golden/)
dL_ddirect_shs of the reference is NOT stored: the reference accumulates it with an unsynchronised `+=` from
every thread (render_equation.cu:447-449), so its value is a race, not a specification. Case "b" keeps
S_direct <= S_incident: the reference's incident-SH gradient loop runs to S_direct (:453) and would write out of
bounds (into the next surfel's rows) otherwise.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_cuda  # noqa: E402
from oracle import render_equation_sh_oracle as RO  # noqa: E402


def main():
    r = ref_cuda.RefReq()
    out = {}
    for tag, (P, Si, Sd, Sv, Ns) in {"a": (300, 16, 16, 16, 24), "b": (130, 16, 9, 4, 40)}.items():
        t = RO.make_inputs(P, Si, Sd, Sv, seed=len(tag) + P, device="cuda")
        g = torch.Generator().manual_seed(5)
        rnd = torch.rand(P, Ns, 1, generator=g).cuda()
        g_pbr = torch.randn(P, 3, generator=g).cuda()
        g_dl = torch.randn(P, 3, generator=g).cuda()
        fw = r.forward(t, Ns, False)
        fwt = r.forward(t, Ns, True, rnd)
        fc = r.forward_complex(t, Ns)
        bw = r.backward(t, Ns, fw["incident_dirs"], g_pbr, g_dl)
        for k, v in t.items():
            out[f"{tag}_in_{k}"] = v.cpu().numpy()
        out[f"{tag}_meta"] = np.array([P, Si, Sd, Sv, Ns])
        out[f"{tag}_rand"] = rnd.cpu().numpy()
        out[f"{tag}_g_pbr"] = g_pbr.cpu().numpy()
        out[f"{tag}_g_dl"] = g_dl.cpu().numpy()
        for k, v in fw.items():
            out[f"{tag}_fw_{k}"] = v.cpu().numpy()
        for k, v in fwt.items():
            out[f"{tag}_fwt_{k}"] = v.cpu().numpy()
        for k, v in fc.items():
            out[f"{tag}_fc_{k}"] = v.cpu().numpy()
        for k, v in bw.items():
            if k != "dL_ddirect_shs":
                out[f"{tag}_bw_{k}"] = v.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "ref_req_sh_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
