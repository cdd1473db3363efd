#!/usr/bin/env python
"""GPU diagnostic (test infrastructure): how far are the MLP modes from the truth, and is the tensor core's fp32 accumulation
round-to-nearest?  (1) `stnerf_selftest_umma_accum` for growing accumulation lengths; (2) sigma / rgb of the background and a
performer SpaceNet on the sample positions of real rays: fp32 SIMT, exact, mixed vs a float64 evaluation of the same weights,
next to the error of the oracle's (= reference's) own fp32 evaluation.  Prints one JSON line."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "st-nerf_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
import cases as C
from oracle import stnerf_oracle as O
from stnerf_b200 import _lib as L
from tests_support import build_case_model

res = {"accum_probe": {}}
for reps in (1, 4, 16, 64, 256):
    mx, ms = ctypes.c_float(), ctypes.c_float()
    L.check(L.lib().stnerf_selftest_umma_accum(reps, ctypes.byref(mx), ctypes.byref(ms)), "accum probe")
    res["accum_probe"][str(4 * reps) + " MMAs"] = {"max_abs_err": mx.value, "mean_signed_rel_err": ms.value, "in_ulps_2^-24": ms.value / 2.0 ** -24}

name = "scale_tkd2_16k"
case = C.SCALE_CASES[name]
rays, jit, u = C.scale_inputs(case)
n = 1024
sd = C.state_dict_for(case)
# sample positions of the background layer (coarse depths) of the first n rays
t = (torch.arange(64)[None] + jit[0, :n]) * 0.2 + 0.5
pos = (rays[:n, None, :3] + t[..., None] * rays[:n, None, 3:6]).reshape(-1, 3).contiguous()
dirs = rays[:n, None, 3:6].expand(-1, 64, -1).reshape(-1, 3).contiguous()
tm = torch.full((pos.shape[0], 1), 10.0)
res["nets"] = {}
def truth(wd, pos, dirs, tm, use_time):
    w64 = {k: v.double() for k, v in wd.items()}
    rgb, sig = O.spacenet_forward(w64, pos.double(), dirs.double(), tm.double() if use_time else None)
    return rgb, sig
for layer, fine, label in ((0, False, "bkgd coarse"), (1, False, "performer 1 coarse"), (0, True, "bkgd fine")):
    if True:
        pre = ("bkgd_spacenet_fine." if fine else "bkgd_spacenet.") if layer == 0 else ("spacenets_fine.%d." % (layer - 1) if fine else "spacenets.%d." % (layer - 1))
        wd = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    use_time = layer > 0 and case["space_time"]
    rgb64, sig64 = truth(wd, pos, dirs, tm, use_time)
    with torch.no_grad():
        rgb32, sig32 = O.spacenet_forward(wd, pos, dirs, tm if use_time else None)
    scale = sig64.abs().clamp(min=1.0)
    entry = {"sigma_range": [float(sig64.min()), float(sig64.max())],
             "reference fp32 (CPU)": {"sigma_rel_rms": float(((sig32.double() - sig64) / scale).pow(2).mean().sqrt()),
                                      "sigma_rel_max": float(((sig32.double() - sig64) / scale).abs().max()),
                                      "sigma_rel_mean_signed": float(((sig32.double() - sig64) / scale).mean())}}
    for prec in ("fp32", "exact", "exact_cf", "mixed"):
        model = build_case_model(dict(case, name=name), precision=prec)
        nat = model._ensure_native(torch.device("cuda", 0))
        rgb, sig = nat.spacenet(layer, fine, pos.cuda(), dirs.cuda(), tm.reshape(-1).cuda() if use_time else None)
        e = (sig.cpu().double().reshape(-1, 1) - sig64.reshape(-1, 1)) / scale.reshape(-1, 1)
        er = (rgb.cpu().double() - rgb64)
        entry[prec] = {"sigma_rel_rms": float(e.pow(2).mean().sqrt()), "sigma_rel_max": float(e.abs().max()), "sigma_rel_mean_signed": float(e.mean()),
                       "rgb_logit_abs_rms": float(er.pow(2).mean().sqrt()), "rgb_logit_abs_max": float(er.abs().max())}
        del model
    res["nets"][label] = entry
print(json.dumps(res))
