#!/usr/bin/env python
"""Small driver for ncu and A/B timing: renders `--rays` rays of view 0 of the bench workload `--calls` times and prints
the per-kernel-class CUDA-event times (stnerf_profile_*)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "st-nerf_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import bench as B
from stnerf_b200.config import make_cfg
import modeling
from stnerf_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=65536)
ap.add_argument("--precision", default="exact")
ap.add_argument("--calls", type=int, default=2)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--chunk", type=int, default=0)
a = ap.parse_args()
sd, _ = B.load_weights()
bkgd, frames, cams = B.scene_setup()
m = modeling.build_layered_model(make_cfg(B.LAYERS, B.N1, B.N2, True, a.precision, a.chunk))
m.load_state_dict(sd); m.set_bkgd_bbox(bkgd); m.set_bboxes(frames)
dev = torch.device("cuda", 0)
nat = m._ensure_native(dev)
nat.set_scene(m._resolve_scene(torch.tensor(B.FRAME_IDS), 0.0, 0.0))
K, T = cams[0]
rows = (a.rays + B.W - 1) // B.W
rays = ops.generate_rays(K, T, B.H, B.W, frame_ids=B.FRAME_IDS, row0=(B.H - rows) // 2, n_rows=rows)[:a.rays].contiguous()
for i in range(a.warm):
    nat.render(rays, B.N1, B.N2, seed=i + 1)
torch.cuda.synchronize()
nat.profile_begin()
for i in range(a.calls):
    nat.render(rays, B.N1, B.N2, seed=100 + i)
prof = nat.profile_end()
print(json.dumps({"lib": os.environ.get("STNERF_B200_LIB", "default"), "rays": rays.shape[0], "calls": a.calls, "chunk": a.chunk,
                  "ms_per_call": {k: round(v["ms"] / a.calls, 3) for k, v in prof.items()},
                  "points": {k: v["points"] / a.calls for k, v in prof.items()}}))
