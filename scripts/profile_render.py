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
ap.add_argument("--workload", default="taekwondo2", choices=list(B.WORKLOADS))
ap.add_argument("--fine-only", action="store_true", help="the renderer fast path: no coarse images (coarse rgb/sigma never reach HBM)")
a = ap.parse_args()
wl = B.WORKLOADS[a.workload]
H, W, N1, N2 = wl["H"], wl["W"], wl["n1"], wl["n2"]
sd, _ = B.load_weights(wl)
bkgd, frames, cams = B.scene_setup(wl)
m = modeling.build_layered_model(make_cfg(wl["layers"], N1, N2, wl["space_time"], a.precision, a.chunk))
m.load_state_dict(sd); m.set_bkgd_bbox(bkgd); m.set_bboxes(frames)
m.near = wl["near"]
dev = torch.device("cuda", 0)
nat = m._ensure_native(dev)
scene = m._resolve_scene(torch.tensor(wl["frame_ids"]), wl["thr"][0], wl["thr"][1])
nat.set_scene(scene)
K, T = cams[0]
rows = (a.rays + W - 1) // W
rays = ops.generate_rays(K, T, H, W, frame_ids=wl["frame_ids"], row0=(H - rows) // 2, n_rows=rows)[:a.rays].contiguous()
view = nat.make_view(K, T, wl["frame_ids"], scene, 7)


def call(seed):
    if a.fine_only:      # same rows through the views API with coarse_images = NULL
        nat.render_views([view], H, W, N1, N2, row0=(H - rows) // 2, row_step=1, n_rows=rows)
    else:
        nat.render(rays, N1, N2, seed=seed)


for i in range(a.warm):
    call(i + 1)
torch.cuda.synchronize()
nat.profile_begin()
for i in range(a.calls):
    call(100 + i)
prof = nat.profile_end()
print(json.dumps({"lib": os.environ.get("STNERF_B200_LIB", "default"), "rays": rays.shape[0], "calls": a.calls, "chunk": a.chunk,
                  "ms_per_call": {k: round(v["ms"] / a.calls, 3) for k, v in prof.items()},
                  "points": {k: v["points"] / a.calls for k, v in prof.items()}}))
