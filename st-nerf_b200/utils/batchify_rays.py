"""utils/batchify_rays.py of the reference, over the native renderer.

The reference splits rays into 3584-ray chunks, calls the model per chunk and concatenates 3*(2+2(L+1)) outputs
(utils/batchify_rays.py:57-140).  The native call streams any number of rays through its own internal chunking,
so the split is not observable -- provided all rays of a call share their frame-id columns, which holds for every
reference caller (one image per call, data/datasets/ray_dataset.py:276-281)."""
import torch


def layered_batchify_ray(model, rays, labels, bboxes, chuncks=512 * 7, near_far=None, near_far_points=[],
                         density_threshold=0, bkgd_density_threshold=0):
    """utils/batchify_rays.py:51-140."""
    N = rays.size(0)
    if N < chuncks:
        # quirk kept: small calls do not forward the thresholds, so the model defaults (1e-4, 0) apply (:53-54)
        return model(rays, labels, bboxes, near_far=near_far, near_far_points=near_far_points)
    stage2, stage1, stage2_layer, stage1_layer, ray_mask = model(
        rays, labels, bboxes, near_far=near_far, near_far_points=near_far_points,
        density_threshold=density_threshold, bkgd_density_threshold=bkgd_density_threshold)
    return stage2, stage1, stage2_layer, stage1_layer, ray_mask


def layered_batchify_ray_big(model, rays, labels, bboxes, chuncks=512 * 7, near_far=None, near_far_points=[],
                             density_threshold=0, bkgd_density_threshold=0, **unused):
    """The reference version (:144-233) passes kwargs forward() does not accept and cannot run; same semantics as
    layered_batchify_ray here."""
    return layered_batchify_ray(model, rays, labels, bboxes, chuncks, near_far, near_far_points, density_threshold,
                                bkgd_density_threshold)


def batchify_ray(model, rays, bboxes, chuncks=1024 * 7, near_far=None, near_far_points=[], density_threshold=0,
                 bkgd_density_threshold=0):
    """utils/batchify_rays.py:4-48: legacy 3-tuple form `(stage2, stage1, ray_mask)`.  The layered model's 5-tuple is
    reduced to it (mixed images + the background hit mask)."""
    labels = torch.zeros(rays.size(0), device=rays.device)
    if rays.size(0) < chuncks:
        out = model(rays, labels, bboxes, near_far=near_far, near_far_points=near_far_points)
    else:
        out = model(rays, labels, bboxes, near_far=near_far, near_far_points=near_far_points,
                    density_threshold=density_threshold, bkgd_density_threshold=bkgd_density_threshold)
    return out[0], out[1], out[4][0]
