#!/usr/bin/env python
"""Parity of the GPU path vs the CPU oracle on a larger ray set (default 16384 rays of view 3, 64+128, taekwondo
checkpoint, injected uniforms).  Test-infrastructure script (imports the oracle); prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "st-nerf_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
import bench as B
from oracle import stnerf_oracle as O
from stnerf_b200.config import make_cfg
from stnerf_b200 import split_planes
import modeling

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.set_num_threads(min(32, os.cpu_count() or 1))
sd, data = B.load_weights()
bkgd, frames, cams = B.scene_setup()
K, T = cams[3]
full = O.generate_rays(K, T, B.H, B.W)
idx = torch.linspace(0, B.H * B.W - 1, n).long()
rays = torch.cat([full[idx], torch.tensor(B.FRAME_IDS)[None].expand(n, -1)], 1).contiguous()
g = torch.Generator().manual_seed(11)
jit = torch.rand((3, n, B.N1), generator=g); u = torch.rand((3, n, B.N2), generator=g)
sc = O.resolve_scene(frames, bkgd, B.FRAME_IDS, None, None)
sc.update(scale=None, shift=None, shown=[True] * 3, near=0.0, alpha=1.0, boarder=1e10)
t0 = time.time()
outs = []
with torch.no_grad():
    for c0 in range(0, n, 2048):
        w = O.render(O.split_state_dict(sd, 2), sc, rays[c0:c0 + 2048], B.N1, B.N2, jit[:, c0:c0 + 2048], u[:, c0:c0 + 2048],
                     density_threshold=0.0, bkgd_density_threshold=0.0)
        outs.append(w["fine_mixed"][0])
ref = torch.cat(outs, 0)
cpu_s = time.time() - t0
res = {"rays": n, "cpu_seconds": cpu_s, "data": data}
for prec in ("exact", "mixed", "fp32", "fast"):
    m = modeling.build_layered_model(make_cfg(2, B.N1, B.N2, True, prec)); m.load_state_dict(sd); m.set_bkgd_bbox(bkgd); m.set_bboxes(frames)
    m.inject_uniforms(jit.cuda(), u.cuda())
    with torch.no_grad():
        out = m(rays.cuda(), None, None, density_threshold=0.0, bkgd_density_threshold=0.0)
    got = out[0][0].float().cpu()
    err = (got - ref).abs().max(1)[0]
    mse = float(((got - ref) ** 2).mean())
    res[prec] = {"max": float(err.max()), "frac_over_1e-3": float((err > 1e-3).float().mean()), "mean": float(err.mean()),
                 "psnr_db": 99.0 if mse == 0 else 10 * np.log10(1 / mse)}
print(json.dumps(res))
