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


# ---- host restatement of the in-kernel Philox4x32-10 stream (csrc/common.cuh: philox_uniform) ---------------------------------
# Lets a test hand the CPU oracle exactly the uniforms the kernels draw when none are injected: jitter of layer i = stream i,
# resampling uniforms of layer i = stream 64+i, counter = (ray id lo, ray id hi, idx >> 2, stream), key = seed, lane = idx & 3.
def philox_uniforms(seed: int, stream: int, ray_ids, count: int) -> np.ndarray:
    """(len(ray_ids), count) float32 uniforms in [0,1) with 24 random bits."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    ray_ids = np.asarray(ray_ids, dtype=np.uint64)[:, None]
    blk = (np.arange(count, dtype=np.uint64) >> np.uint64(2))[None, :]
    mask = np.uint64(0xFFFFFFFF)
    c0 = np.broadcast_to(ray_ids & mask, (ray_ids.shape[0], count)).copy()
    c1 = np.broadcast_to(ray_ids >> np.uint64(32), c0.shape).copy()
    c2 = np.broadcast_to(blk, c0.shape).copy()
    c3 = np.full(c0.shape, stream, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    lane = (np.arange(count) & 3)[None, :]
    v = np.where(lane == 0, c0, np.where(lane == 1, c1, np.where(lane == 2, c2, c3)))
    return ((v >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def philox_draws(seed: int, l: int, n_rays: int, n1: int, n2: int, ray_ids=None):
    """jitter (l, n_rays, n1), u (l, n_rays, n2) as torch tensors: the draws of a render call with this seed."""
    ids = np.arange(n_rays, dtype=np.uint64) if ray_ids is None else ray_ids
    jit = np.stack([philox_uniforms(seed, i, ids, n1) for i in range(l)], 0)
    u = np.stack([philox_uniforms(seed, 64 + i, ids, n2) for i in range(l)], 0)
    return torch.from_numpy(jit), torch.from_numpy(u)
