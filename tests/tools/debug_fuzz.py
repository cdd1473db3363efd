import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "st-nerf_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
import cases as C
from stnerf_b200.config import make_cfg
from stnerf_b200 import synthetic
from oracle import stnerf_oracle as O
import modeling
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 3
rs = np.random.RandomState(100 + seed)
L = int(rs.randint(1, 4)); n1 = int(rs.randint(3, 97)); n2 = int(rs.choice([0, rs.randint(1, 161)])); n_rays = int(rs.randint(2, 3000)); space_time = bool(rs.randint(0, 2))
print("L", L, "n1", n1, "n2", n2, "N", n_rays, "st", space_time)
sd = synthetic.synthetic_state_dict(L, space_time, seed=seed)
case = dict(L=L, n_rays=max(n_rays, 8), ray_seed=200 + seed, frame_ids=[0] + [10 + 0.5 * (seed % 2) + i for i in range(L)], n1=n1, n2=n2)
rays = C.rays_for(case)[:n_rays].contiguous()
g = torch.Generator().manual_seed(5)
jit = torch.rand((L + 1, n_rays, n1), generator=g); u = torch.rand((L + 1, n_rays, max(n2, 1)), generator=g)
bkgd, frames = synthetic.synthetic_boxes(L)
res = {}
for prec in ("exact", "fp32"):
    m = modeling.build_layered_model(make_cfg(L, n1, n2, space_time, prec)); m.load_state_dict(sd); m.set_bkgd_bbox(bkgd); m.set_bboxes(frames)
    m.inject_uniforms(jit.cuda(), u.cuda())
    with torch.no_grad():
        res[prec] = C.flatten_outputs(*m(rays.cuda(), None, None, only_coarse=(n2 == 0), density_threshold=0.0, bkgd_density_threshold=0.0))
sc = O.resolve_scene(frames, bkgd, case["frame_ids"], None, None); sc.update(scale=None, shift=None, shown=[True] * (L + 1), near=0.0, alpha=1.0, boarder=1e10)
w = O.render(O.split_state_dict(sd, L), sc, rays, n1, n2, jit, u if n2 else None, only_coarse=(n2 == 0), density_threshold=0.0, bkgd_density_threshold=0.0)
want = C.flatten_outputs(w["fine_mixed"], w["coarse_mixed"], w["fine_layer"], w["coarse_layer"], w["ray_mask"])
for k in sorted(want):
    if k.startswith("ray_mask"): continue
    e1 = np.abs(res["exact"][k] - want[k]).max(); e2 = np.abs(res["fp32"][k] - want[k]).max(); e3 = np.abs(res["exact"][k] - res["fp32"][k]).max()
    if max(e1, e2, e3) > 5e-4: print("%-22s exact-vs-oracle %.2e  fp32-vs-oracle %.2e  exact-vs-fp32 %.2e" % (k, e1, e2, e3))
k = "fine_mixed.rgb"
bad = np.where(np.abs(res["exact"][k] - res["fp32"][k]).max(1) > 1e-3)[0]
print("rays with exact-vs-fp32 > 1e-3:", bad[:10], "of", n_rays)
for r in bad[:3]:
    print(r, "exact", res["exact"][k][r], "fp32", res["fp32"][k][r], "oracle", want[k][r], "acc", want["fine_mixed.acc"][r])
