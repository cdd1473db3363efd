#!/usr/bin/env python
"""CPU study (oracle-based, test infrastructure): can any layer of the SpaceNets run with fewer than three MMA terms?
For every variant the oracle's Linear layers are replaced by the fp16 split with round-to-nearest accumulation
(`Ahi*Whi + Alo*Whi + Ahi*Wlo`), with ONE term dropped in a chosen set of layers, and the 16 384-ray taekwondo fixture of the
unmodified reference is rendered.  (The tensor core's truncating accumulation is not emulated, so these numbers are a LOWER
bound on what the GPU would show.)

    python tests/tools/emulate_term_dropping.py [n_rays]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "st-nerf_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
import torch.nn.functional as F
import cases as C
from oracle import stnerf_oracle as O

name = "scale_tkd2_16k"
case = C.SCALE_CASES[name]
n = int(sys.argv[1]) if len(sys.argv) > 1 else case["n_rays"]
torch.set_num_threads(os.cpu_count() or 1)
rays, jit, u = C.scale_inputs(case)
rays, jit, u = rays[:n], jit[:, :n], u[:, :n]
gold = C.load_golden(name)
sd = C.state_dict_for(case)
nets = O.split_state_dict(sd, case["L"])
sc = C.scene_for(case)
layer_of = {}
for k, v in sd.items():
    if k.endswith(".weight"):
        layer_of[v.data_ptr()] = k.split(".")[-3] + "." + k.split(".")[-2] if "motion" not in k else "motion." + k.split(".")[-2]
real_linear = F.linear
MODE = {"drop": {}}          # layer name -> "alo" | "wlo"


def split(x):
    hi = x.half().float()
    return hi, (x - hi).half().float()


def patched(x, w, b=None):
    if w.shape[0] <= 3:
        return real_linear(x, w, b)
    name_ = layer_of.get(w.data_ptr(), "?")
    xh, xl = split(x.clamp(-65504.0, 65504.0))
    wh, wl = split(w)
    acc = real_linear(xh, wh)
    drop = MODE["drop"].get(name_)
    if drop != "alo":
        acc = acc + real_linear(xl, wh)
    if drop != "wlo":
        acc = acc + real_linear(xh, wl)
    return acc if b is None else acc + b


def run():
    out = []
    with torch.no_grad():
        for c0 in range(0, n, 2048):
            w = O.render(nets, sc, rays[c0:c0 + 2048], case["n1"], case["n2"], jit[:, c0:c0 + 2048], u[:, c0:c0 + 2048],
                         density_threshold=case["thr"][0], bkgd_density_threshold=case["thr"][1])
            out.append(w["fine_mixed"][0])
    return torch.cat(out, 0).numpy()


F.linear = patched
O.F.linear = patched
TRUNK = ["stage1.0", "stage1.2", "stage1.4", "stage1.6", "stage2.0", "stage2.2", "stage2.4"]
variants = [("3 terms everywhere", {})]
for lay in TRUNK:
    variants.append(("no Ahi*Wlo in %s" % lay, {lay: "wlo"}))
for lay in ("stage1.0", "stage2.4"):
    variants.append(("no Alo*Whi in %s" % lay, {lay: "alo"}))
variants.append(("no Ahi*Wlo in the whole trunk", {l_: "wlo" for l_ in TRUNK}))
variants.append(("no Ahi*Wlo in rgb_net.1", {"rgb_net.1": "wlo"}))
variants.append(("no lo terms in rgb_net.1 (= mixed)", None))
res = {"rays": n, "note": "round-to-nearest accumulation (the GPU's truncating accumulator adds to all of these)"}
for label, drop in variants:
    if drop is None:
        MODE["drop"] = {}
        def patched2(x, w, b=None, _p=patched):
            if layer_of.get(w.data_ptr()) == "rgb_net.1":
                return real_linear(x.half().float(), w.half().float(), b)
            return _p(x, w, b)
        F.linear = patched2; O.F.linear = patched2
    else:
        MODE["drop"] = drop
    rgb = run()
    err = np.abs(rgb - gold["fine_mixed.rgb"][:n]).max(1)
    res[label] = {"max": float(err.max()), "pixels_over_1e-3": int((err > 1e-3).sum()), "mean": float(err.mean()), "p999": float(np.sort(err)[int(0.999 * n)])}
    print(label, res[label], flush=True)
print(json.dumps(res))
