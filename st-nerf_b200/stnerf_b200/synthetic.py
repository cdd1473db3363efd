"""Synthetic inputs of the benchmark / examples: the reference ships neither dataset nor cameras, so the
scene boxes, the 16-camera rig and (when the checkpoint copies are absent) seeded weights are generated here.
Pure input generation -- no part of the hot path.  The test-side CPU restatement carries an identical copy (the product
tree and the test infrastructure never import each other's arithmetic); tests/test_host_logic.py checks they stay in sync."""
from __future__ import annotations

import math

import torch

F32 = torch.float32


def corners_from_minmax(bmin, bmax) -> torch.Tensor:
    """Corner order of data/datasets/frame_dataset.py:187-188."""
    x0, y0, z0 = [float(v) for v in bmin]
    x1, y1, z1 = [float(v) for v in bmax]
    return torch.tensor([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0],
                         [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], dtype=F32)


def synthetic_boxes(layer_num: int, n_frames: int = 101):
    """bkgd box (-6,-6,-1)..(6,6,4); performers 0.8x0.8x1.8 on z=0, centres on x in [-2,2]."""
    bkgd = corners_from_minmax((-6, -6, -1), (6, 6, 4))[None]
    per = []
    for i in range(layer_num):
        cx = 0.0 if layer_num == 1 else -2.0 + 4.0 * i / (layer_num - 1)
        per.append(corners_from_minmax((cx - 0.4, -0.4, 0.0), (cx + 0.4, 0.4, 1.8)))
    per = torch.stack(per, 0)                                         # (L,8,3)
    # a slow drift so fractional frame ids exercise the bbox lerp
    frames = torch.stack([per + 0.002 * f * torch.tensor([1.0, 0.5, 0.0]) for f in range(n_frames)], 0)
    return bkgd, frames


def synthetic_camera(view: int, n_views: int, h: int, w: int):
    """circle radius 5, height 1.0, looking at (0,0,0.9), up +z, OpenCV c2w; fx=fy=0.78*W."""
    az = 2 * math.pi * view / n_views
    eye = torch.tensor([5 * math.cos(az), 5 * math.sin(az), 1.0], dtype=torch.float64)
    tgt = torch.tensor([0.0, 0.0, 0.9], dtype=torch.float64)
    fwd = tgt - eye; fwd = fwd / fwd.norm()
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    right = torch.linalg.cross(fwd, up); right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    T = torch.eye(4, dtype=torch.float64)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = right, down, fwd, eye
    K = torch.tensor([[0.78 * w, 0, w / 2], [0, 0.78 * w, h / 2], [0, 0, 1]], dtype=torch.float64)
    return K.to(F32), T.to(F32)


def synthetic_state_dict(layer_num: int, use_space_time: bool, seed: int = 0, gain: float = 1.6):
    """Seeded random weights with the reference's key names/shapes.

    ``gain`` scales the nn.Linear-style uniform init so activations do not collapse to
    zero through 8 layers (keeps sigma/rgb in a range where parity errors are visible).
    Generated with numpy's legacy RandomState so they are identical on every host.
    """
    import numpy as np
    rs = np.random.RandomState(seed)
    sd = {}

    def lin(name, out_f, in_f, g=gain):
        b = g / math.sqrt(in_f)
        sd[name + ".weight"] = torch.from_numpy(rs.uniform(-b, b, (out_f, in_f)).astype(np.float32))
        sd[name + ".bias"] = torch.from_numpy(rs.uniform(-b, b, (out_f,)).astype(np.float32))

    def spacenet(prefix, use_time):
        lin(prefix + "stage1.0", 256, 63)
        for i in (2, 4, 6):
            lin(prefix + "stage1.%d" % i, 256, 256)
        lin(prefix + "stage2.0", 256, 319)
        for i in (2, 4):
            lin(prefix + "stage2.%d" % i, 256, 256)
        lin(prefix + "density_net.0", 1, 256, g=gain * 4)
        lin(prefix + "rgb_net.1", 128, 256 + 27 + (21 if use_time else 0))
        lin(prefix + "rgb_net.3", 3, 128)

    spacenet("bkgd_spacenet.", False)
    spacenet("bkgd_spacenet_fine.", False)
    for i in range(layer_num):
        spacenet("spacenets.%d." % i, use_space_time)
    for i in range(layer_num):
        spacenet("spacenets_fine.%d." % i, use_space_time)
    for i in range(layer_num):
        p = "time_deform_nets.%d.motion_net." % i
        lin(p + "0", 128, 84)
        for j in (2, 4, 6, 8):
            lin(p + "%d" % j, 128, 128)
        lin(p + "10", 3, 128, g=0.1)
    return sd
