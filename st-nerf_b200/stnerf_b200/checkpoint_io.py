"""Checkpoint discovery / loading and camera-file parsing (host side only).

    get_iteration_path      data/datasets/utils.py:42-60      newest `layered_rfnr_checkpoint_<iter>.pt` of a directory
    load_checkpoint         render/layered_neural_renderer.py:109-117   `torch.load(...)['model']`, keys missing from the
                            file back-filled from the freshly initialised model, then `load_state_dict`
    load_checkpoint_cached  the same, with the packed tensor-core weight image cached next to the `.pt` (8f row 3)
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


# ---- packed-weight cache next to the checkpoint -------------------------------------------------------------
CACHE_MAGIC = b"STNB200C"
CACHE_SUFFIX = ".b200w"


def file_sha256(path, chunk=1 << 20) -> bytes:
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(chunk)
            if not b:
                break
            h.update(b)
    return h.digest()


def read_weight_cache(pt_path, cache_path=None):
    """The packed image cached for `pt_path`, or None if there is none / it belongs to other file contents / is damaged."""
    cache_path = cache_path or pt_path + CACHE_SUFFIX
    if not os.path.isfile(cache_path):
        return None
    with open(cache_path, "rb") as f:
        head = f.read(8 + 32 + 8)
        if len(head) != 48 or head[:8] != CACHE_MAGIC or head[8:40] != file_sha256(pt_path):
            return None
        n = int.from_bytes(head[40:48], "little")
        image = f.read(n + 1)
    return image if len(image) == n else None


def write_weight_cache(pt_path, image: bytes, cache_path=None):
    cache_path = cache_path or pt_path + CACHE_SUFFIX
    tmp = cache_path + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as f:
        f.write(CACHE_MAGIC + file_sha256(pt_path) + len(image).to_bytes(8, "little") + image)
    os.replace(tmp, cache_path)                      # atomic: concurrent ranks either see the old file or the whole new one
    return cache_path


def load_checkpoint_cached(model, path, device=None, cache_path=None):
    """`load_checkpoint` with the packed MMA-layout image cached next to the `.pt` (`<path>.b200w`, keyed by the SHA-256 of
    the checkpoint file).  First call: torch.load + back-fill + pack on the device + write the cache.  Later calls: the
    image goes straight to the device (no unpickling, no re-packing); a changed `.pt`, a damaged cache or an image the
    library rejects (other layer configuration / library version) falls back to the first path and rewrites the cache.
    Returns "cache" or "checkpoint"."""
    from ._lib import StnerfError
    if not (hasattr(model, "load_packed") and hasattr(model, "export_packed")):
        raise TypeError("load_checkpoint_cached needs a stnerf_b200 model (build_layered_model); got %s" % type(model).__name__)
    if not torch.cuda.is_available():                  # nothing to pack on: plain load; rendering will fail loudly later
        load_checkpoint(model, path)
        return "checkpoint"
    image = read_weight_cache(path, cache_path)
    if image is not None:
        model.load_packed(image, state_dict_source=lambda: torch.load(path, map_location="cpu")["model"])
        try:
            model._ensure_native(_device(device))        # uploads the image; the library validates it against this model
            return "cache"
        except StnerfError:
            pass                                         # other layer configuration / library version: rebuild below
    load_checkpoint(model, path)
    try:
        write_weight_cache(path, model.export_packed(_device(device)), cache_path)
    except OSError:
        pass                                         # read-only checkpoint directory: run without the cache
    return "checkpoint"


def _device(device):
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


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
