"""Checkpoint discovery / loading and camera-file parsing (host side only).

    get_iteration_path      data/datasets/utils.py:42-60      newest `layered_rfnr_checkpoint_<iter>.pt` of a directory
    load_checkpoint         render/layered_neural_renderer.py:109-117   `torch.load(...)['model']`, keys missing from the
                            file back-filled from the freshly initialised model, then `load_state_dict`
    read_intrinsics         data/datasets/utils.py:20-40      `K.txt`: 9 floats per line -> (M,3,3)
    campose_to_extrinsic    data/datasets/utils.py:6-17       `RT_c2w.txt` rows of 12 floats -> (M,4,4)
"""
from __future__ import annotations

import glob
import os

import numpy as np
import torch


def get_iteration_path(root_dir, fix_iter=-1):
    if fix_iter != -1:
        return os.path.join(root_dir, "frame", "layered_rfnr_checkpoint_%d.pt" % fix_iter)
    if not os.path.exists(root_dir):
        return None
    max_iter = -1
    for file_name in glob.glob(os.path.join(root_dir, "layered_rfnr_checkpoint_*.pt")):
        parts = file_name.split("/")[-1].split("_")
        if len(parts) != 4:                      # `..._<epoch>_<step>.pt` intermediate saves are skipped
            continue
        max_iter = max(max_iter, int(parts[-1].split(".")[0]))
    path = os.path.join(root_dir, "layered_rfnr_checkpoint_%d.pt" % max_iter)
    return path if os.path.exists(path) else None


def load_checkpoint(model, path, map_location="cpu"):
    """Load `path` into `model` with the reference's back-fill of keys the file lacks (its loader never fails on a
    checkpoint saved before a sub-network existed).  Returns the list of back-filled keys."""
    model_dict = torch.load(path, map_location=map_location)["model"]
    fresh = model.state_dict()
    missing = [k for k in fresh if k not in model_dict]
    for k in missing:
        model_dict[k] = fresh[k]
    model.load_state_dict(model_dict)
    return missing


def read_intrinsics(fn_intrinsic):
    Ks = []
    with open(fn_intrinsic) as fo:
        for line in fo.readlines():
            v = [float(x) for x in line.split()[0:9]]
            Ks.append(np.vstack([np.array(v[0:3]), np.array(v[3:6]), np.array(v[6:9])]))
    return np.stack(Ks)


def campose_to_extrinsic(camposes):
    if camposes.shape[1] != 12:
        raise Exception(" wrong campose data structure!")
    res = np.zeros((camposes.shape[0], 4, 4))
    res[:, 0, :] = camposes[:, 0:4]
    res[:, 1, :] = camposes[:, 4:8]
    res[:, 2, :] = camposes[:, 8:12]
    res[:, 3, 3] = 1.0
    return res


# ---- checkpoint copies for hosts without the reference tree -----------------------------------------------------------
_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHECKPOINT_DIRS = [os.path.join(_REPO_ROOT, "oracle", "_ref", "ckpt"), "/root/reference/outputs"]


def find_checkpoint(scene: str, dirs=None):
    """Path of the shipped checkpoint of `scene` ('taekwondo' | 'walking'): the git-ignored copy that travels with the
    repo (`<repo>/oracle/_ref/ckpt/<scene>.pt`) or the reference tree's `outputs/<scene>/layered_rfnr_checkpoint_1.pt`."""
    for d in (dirs or CHECKPOINT_DIRS):
        for p in (os.path.join(d, scene + ".pt"), os.path.join(d, scene, "layered_rfnr_checkpoint_1.pt")):
            if os.path.isfile(p):
                return p
    return None


def replicate_layers(sd: dict, L: int) -> dict:
    """configurations with more performers than the checkpoint reuse its nets round-robin."""
    have = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("spacenets."))
    out = {k: v for k, v in sd.items() if k.startswith("bkgd_")}
    for i in range(L):
        for grp in ("spacenets", "spacenets_fine", "time_deform_nets"):
            src = "%s.%d." % (grp, i % have)
            for k, v in sd.items():
                if k.startswith(src):
                    out["%s.%d.%s" % (grp, i, k[len(src):])] = v
    return out
