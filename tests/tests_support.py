"""Helpers shared by the gpu tests, smoke() and bench.py: build the facade model for a golden case and run it."""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "st-nerf_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases as C  # noqa: E402


from stnerf_b200.config import make_cfg  # noqa: E402,F401


def build_case_model(name, precision="exact", chunk_rays=0, sd=None):
    """`name`: key of cases.CASES / cases.SCALE_CASES, or a case dict."""
    import modeling
    case = name if isinstance(name, dict) else (C.CASES[name] if name in C.CASES else C.SCALE_CASES[name])
    sd = sd if sd is not None else C.state_dict_for(case)
    if sd is None:
        return None
    model = modeling.build_layered_model(make_cfg(case["L"], case["n1"], case["n2"], case["space_time"], precision,
                                                  chunk_rays), 0, case.get("scale"), case.get("shift"))
    model.load_state_dict(sd)
    bkgd, frames = C.boxes_for(case)
    model.set_bkgd_bbox(bkgd)
    model.set_bboxes(frames)
    model.near = case.get("near", 0.0)
    model.alpha = case.get("alpha", 1.0)
    for i in case.get("hidden", []):
        model.hide_layer(i)
    return model.cuda()


def run_case_native(name, precision="exact", chunk_rays=0, inject=True, rays=None, uniforms=None):
    """Render a golden case through the facade (model.forward); returns the flat numpy dict of cases.flatten_outputs."""
    case = C.CASES[name]
    model = build_case_model(name, precision, chunk_rays)
    if model is None:
        return None
    rays = C.rays_for(case) if rays is None else rays
    jit, u = C.uniforms_for(case) if uniforms is None else uniforms
    dev = torch.device("cuda", 0)
    if inject:
        model.inject_uniforms(jit.to(dev), None if u is None else u.to(dev))
    with torch.no_grad():
        out = model(rays.to(dev), torch.zeros(rays.shape[0], device=dev), None,
                    only_coarse=case.get("only_coarse", False), density_threshold=case["thr"][0],
                    bkgd_density_threshold=case["thr"][1])
    torch.cuda.synchronize()
    return C.flatten_outputs(*out)


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)       # utils/metrics.py:16-17


def compare(flat, gold, keys=None):
    """Per-key max abs error + fraction of rays whose rgb error exceeds 1e-3."""
    rep = {}
    for k in (keys or gold):
        if k.startswith("ray_mask"):
            rep[k] = float((flat[k] != gold[k]).sum())
        else:
            rep[k] = float(np.abs(flat[k].astype(np.float64) - gold[k]).max())
    return rep
