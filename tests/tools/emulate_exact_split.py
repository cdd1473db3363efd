#!/usr/bin/env python
"""CPU study (oracle-based, test infrastructure): how close to fp32 is the 3-term fp16 split `Ahi*Whi + Alo*Whi + Ahi*Wlo` AS
THE KERNEL FORMS IT -- lo = fp16(x - fp16(x)), which for |x| < 2^-2 lands in fp16's SUBNORMAL range (spacing 2^-24) -- and what
do power-of-two operand scales buy?  Renders the rays of a parity-at-scale fixture with the oracle's Linear layers replaced by
the emulated split and compares with the UNMODIFIED reference's fixture (tests/golden/scale_*.npz).

    python tests/tools/emulate_exact_split.py [case] [n_rays]
"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "st-nerf_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
import torch.nn.functional as F
import cases as C
from oracle import stnerf_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "scale_tkd2_16k"
case = C.SCALE_CASES[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else case["n_rays"]
torch.set_num_threads(os.cpu_count() or 1)
rays, jit, u = C.scale_inputs(case)
rays, jit, u = rays[:n], jit[:, :n], u[:, :n]
gold = C.load_golden(name)
sd = C.state_dict_for(case)
nets = O.split_state_dict(sd, case["L"])
sc = C.scene_for(case)
real_linear = F.linear
MODE = {"kind": "fp32", "w_scale": False, "a_scale": 0}


def split(x):
    hi = x.half().float()
    lo = (x - hi).half().float()          # fp16 RN incl. subnormals, like cvt.rn.f16.f32
    return hi, lo


def patched(x, w, b=None):
    if MODE["kind"] == "fp32" or w.shape[0] <= 3:      # the 1- and 3-wide heads are fp32 FFMA work in the kernel
        return real_linear(x, w, b)
    sw = 1.0
    if MODE["w_scale"]:                                # per-layer power of two that puts max|W| just below 2^14
        sw = 2.0 ** math.floor(math.log2(16384.0 / float(w.abs().max())))
    sa = 2.0 ** MODE["a_scale"]
    xs = (x * sa).clamp(-65504.0, 65504.0)
    xh, xl = split(xs)
    wh, wl = split(w * sw)
    acc = real_linear(xh, wh) + real_linear(xl, wh) + real_linear(xh, wl)
    acc = acc * (1.0 / (sw * sa))
    return acc if b is None else acc + b


def run():
    outs = {"rgb": [], "acc": []}
    with torch.no_grad():
        for c0 in range(0, n, 2048):
            w = O.render(nets, sc, rays[c0:c0 + 2048], case["n1"], case["n2"], jit[:, c0:c0 + 2048], u[:, c0:c0 + 2048],
                         density_threshold=case["thr"][0], bkgd_density_threshold=case["thr"][1])
            outs["rgb"].append(w["fine_mixed"][0]); outs["acc"].append(w["fine_mixed"][2])
    return torch.cat(outs["rgb"], 0).numpy(), torch.cat(outs["acc"], 0).numpy()


F.linear = patched
O.F.linear = patched
res = {"case": name, "rays": n}
for label, kind, ws, a in (("oracle fp32", "fp32", False, 0), ("split as in the kernel", "split", False, 0),
                          ("split, weights scaled", "split", True, 0), ("split, weights scaled + activations x16", "split", True, 4)):
    MODE.update(kind=kind, w_scale=ws, a_scale=a)
    rgb, acc = run()
    err = np.abs(rgb - gold["fine_mixed.rgb"][:n]).max(1)
    ea = np.abs(acc - gold["fine_mixed.acc"][:n]).max(1)
    res[label] = {"max_rgb": float(err.max()), "pixels_over_1e-3": int((err > 1e-3).sum()), "acc_over_1e-3": int((ea > 1e-3).sum()),
                  "mean_rgb": float(err.mean()), "p999": float(np.sort(err)[int(0.999 * n)])}
    print(label, res[label], flush=True)
print(json.dumps(res))
