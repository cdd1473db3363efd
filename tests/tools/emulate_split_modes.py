#!/usr/bin/env python
"""CPU study (oracle-based, test infrastructure): which fp16 rounding source dominates the pixel error of a reduced-pass
tensor-core mode?  Emulates, inside the oracle's MLPs, rounding of (a) activations, (b) weights, (c) both to fp16 before
every Linear of the SpaceNets/MotionNets (fp32 accumulate), renders the same rays and compares with the fp32 oracle."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "st-nerf_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
import torch.nn.functional as F
import bench as B
from oracle import stnerf_oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.set_num_threads(min(32, os.cpu_count() or 1))
WL = WL["W"]ORKLOADS["taekwondo2"]
sd, data = B.load_weights(WL)
bkgd, frames, cams = B.scene_setup(WL)
K, T = cams[3]
full = O.generate_rays(K, T, WL["H"], WL["W"])
idx = torch.linspace(0, WL["H"] * WL["W"] - 1, n).long()
rays = torch.cat([full[idx], torch.tensor(WL["frame_ids"])[None].expand(n, -1)], 1).contiguous()
g = torch.Generator().manual_seed(11)
jit = torch.rand((3, n, WL["n1"]), generator=g); u = torch.rand((3, n, WL["n2"]), generator=g)
sc = O.resolve_scene(frames, bkgd, WL["frame_ids"], None, None)
sc.update(scale=None, shift=None, shown=[True] * 3, near=0.0, alpha=1.0, boarder=1e10)
nets = O.split_state_dict(sd, 2)

real_linear = F.linear
MODE = {"act": False, "w": False, "min_in": 0}
def q(x): return x.half().float()
def patched(x, w, b=None):
    if w.shape[0] > 3 or MODE.get("heads"):          # heads (1- and 3-wide) stay fp32 in the kernel
        if MODE["act"]: x = q(x)
        if MODE["w"]: w = q(w)
    return real_linear(x, w, b)

def run():
    outs = []
    with torch.no_grad():
        for c0 in range(0, n, 2048):
            w = O.render(nets, sc, rays[c0:c0 + 2048], WL["n1"], WL["n2"], jit[:, c0:c0 + 2048], u[:, c0:c0 + 2048],
                         density_threshold=0.0, bkgd_density_threshold=0.0)
            outs.append(w["fine_mixed"][0])
    return torch.cat(outs, 0)

ref = run()
F.linear = patched
O.F.linear = patched
res = {"rays": n}
for name, a, w in (("act_only (Ahi*Whi + Ahi*Wlo)", True, False), ("w_only (Ahi*Whi + Alo*Whi)", False, True), ("both (fast)", True, True)):
    MODE["act"], MODE["w"] = a, w
    got = run()
    err = (got - ref).abs().max(1)[0]
    res[name] = {"max": float(err.max()), "frac_over_1e-3": float((err > 1e-3).float().mean()), "mean": float(err.mean()),
                 "p99": float(err.kthvalue(int(0.99 * n))[0])}
print(json.dumps(res, indent=1))
