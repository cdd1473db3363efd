"""engine/render.py:30-77 of the reference for the layered model.

The reference function expects a single-layer `model(rays, bboxes, ...) -> 3-tuple` that no longer ships; this
adapter supplies what the layered model needs (per-layer frame-id columns) and keeps the signature and the return
value: `(stage2_final, stage1_final)`, each `[rgb (H,W,3), depth (H,W), alpha (H,W)]`."""
import torch

from stnerf_b200 import ops
from utils.batchify_rays import layered_batchify_ray


def render(model, K, T, img_size, ROI=None, bboxes=None, only_coarse=False, near_far=None, frame_ids=None,
           density_threshold=0, bkgd_density_threshold=0):
    model.eval()
    H, W = int(img_size[0]), int(img_size[1])
    l = model.layer_num + 1
    if frame_ids is None:
        frame_ids = [0.0] + [1.0] * (l - 1)
    mask = torch.ones(H, W)
    if ROI is not None:
        mask = torch.zeros(H, W)
        mask[ROI[0]:ROI[0] + ROI[2], ROI[1]:ROI[1] + ROI[3]] = 1.0                 # :34-36
    device = torch.device("cuda", torch.cuda.current_device())
    rays = ops.generate_rays(K, T, H, W, frame_ids=frame_ids, device=device)       # utils/ray_sampling.py:22-72
    sel = (mask > 0.5).reshape(-1).to(device)
    if ROI is not None:
        rays = rays[sel]
    labels = torch.zeros(rays.shape[0], device=device)
    with torch.no_grad():
        if only_coarse:
            out = model(rays, labels, None, only_coarse=True, density_threshold=density_threshold,
                        bkgd_density_threshold=bkgd_density_threshold)
        else:
            out = layered_batchify_ray(model, rays, labels, None, near_far=near_far, density_threshold=density_threshold,
                                       bkgd_density_threshold=bkgd_density_threshold)
    finals = []
    for stage in (out[0], out[1]):
        rgb = torch.zeros(H * W, 3, device=device); depth = torch.zeros(H * W, 1, device=device)
        alpha = torch.zeros(H * W, 1, device=device)
        rgb[sel], depth[sel], alpha[sel] = stage[0], stage[1], stage[2]            # :55-73
        finals.append([rgb.reshape(H, W, 3), depth.reshape(H, W), alpha.reshape(H, W)])
    return finals[0], finals[1]
