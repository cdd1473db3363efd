"""Reference import name `layers` (layers/__init__.py:1-3): per-stage operators on the B200 kernels."""
import torch

from stnerf_b200 import ops


class RaySamplePoint(torch.nn.Module):
    """layers/RaySamplePoint.py:64-107.  forward(rays (N,>=6), bbox (N,L,8,3)) -> lists of t (N,C,1), xyz (N,C,3), mask (N).
    Boxes are read from row 0 (all rows are identical in every reference caller, layered_rfrender.py:195-208)."""

    def __init__(self, coarse_num=64):
        super().__init__()
        self.coarse_num = coarse_num

    def forward(self, rays, bbox, pdf=None, method='coarse'):
        n, l = rays.shape[0], bbox.shape[1]
        ts, pts, masks = [], [], []
        for i in range(l):
            corners = bbox[0, i].detach().cpu()
            jitter = torch.rand((n, self.coarse_num), device=rays.device)          # :98
            t, xyz, mask, _ = ops.intersect_sample(rays, corners[0], corners[6], self.coarse_num, jitter, is_bkgd=(i == 0))
            ts.append(t.unsqueeze(-1)); pts.append(xyz); masks.append(mask)
        return ts, pts, masks


class RaySamplePoint_Near_Far(torch.nn.Module):
    """Unusable in the reference's layered model (reads an attribute that is never set)."""

    def __init__(self, sample_num=75):
        super().__init__()
        self.sample_num = sample_num

    def forward(self, *a, **k):
        raise NotImplementedError("NEAR_FAR sampling is a dead branch of the reference (modeling/layered_rfrender.py:254)")


class VolumeRenderer(torch.nn.Module):
    """layers/render_layer.py:19-58.  forward(depth (N,L,1), rgb (N,L,3), sigma (N,L,1)) -> color, depth, acc, weights (N,L,1)."""

    def __init__(self, use_mask=False, boarder_weight=1e10):
        super().__init__()
        if use_mask:
            raise NotImplementedError("use_mask=True is never used by the layered renderer")
        self.boarder_weight = boarder_weight

    def forward(self, depth, rgb, sigma, noise=0):
        if noise > 0.:
            sigma = sigma + torch.randn_like(sigma) * noise                         # :42-43
        n, s = depth.shape[0], depth.shape[1]
        c, d, a, w = ops.composite(depth.reshape(n, s), rgb, sigma.reshape(n, s), self.boarder_weight)
        return c, d, a, w.unsqueeze(-1)


def make_loss(cfg):
    """layers/loss.py:4-5."""
    return torch.nn.MSELoss()


__all__ = ["RaySamplePoint", "RaySamplePoint_Near_Far", "VolumeRenderer", "make_loss"]
