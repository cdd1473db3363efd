"""Checkpoint discovery / loading and camera-file parsing.  CPU only.  When the reference tree is
present (build container) its own functions (data/datasets/utils.py, a numpy-only module) are executed side by side."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import cases as C
from tests_support import make_cfg

REF_UTILS = "/root/reference/data/datasets/utils.py"


def _ref():
    if not os.path.isfile(REF_UTILS):
        return None
    spec = importlib.util.spec_from_file_location("_ref_dataset_utils", REF_UTILS)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_get_iteration_path(tmp_path):
    from stnerf_b200 import checkpoint_io as io
    ref = _ref()
    d = str(tmp_path)
    assert io.get_iteration_path(os.path.join(d, "nope")) is None
    assert io.get_iteration_path(d) is None                                   # empty directory
    for name in ("layered_rfnr_checkpoint_3.pt", "layered_rfnr_checkpoint_12.pt", "layered_rfnr_checkpoint_5_200.pt",
                 "other.pt"):
        open(os.path.join(d, name), "w").close()
    assert io.get_iteration_path(d) == os.path.join(d, "layered_rfnr_checkpoint_12.pt")
    assert io.get_iteration_path(d, fix_iter=7) == os.path.join(d, "frame", "layered_rfnr_checkpoint_7.pt")
    if ref is not None:
        assert io.get_iteration_path(d) == ref.get_iteration_path(d)
        assert io.get_iteration_path(d, 7) == ref.get_iteration_path(d, 7)
        assert ref.get_iteration_path(os.path.join(d, "nope")) is None


def test_camera_file_parsing(tmp_path):
    from stnerf_b200 import checkpoint_io as io
    ref = _ref()
    rs = np.random.RandomState(3)
    Ks = rs.rand(5, 9) * 1000
    fn = os.path.join(str(tmp_path), "K.txt")
    np.savetxt(fn, Ks)
    got = io.read_intrinsics(fn)
    assert got.shape == (5, 3, 3) and np.array_equal(got, np.loadtxt(fn).reshape(5, 3, 3))
    poses = rs.rand(5, 12)
    ext = io.campose_to_extrinsic(poses)
    assert ext.shape == (5, 4, 4) and np.array_equal(ext[:, :3, :].reshape(5, 12), poses) and np.all(ext[:, 3] == [0, 0, 0, 1])
    with pytest.raises(Exception):
        io.campose_to_extrinsic(rs.rand(5, 11))
    if ref is not None:
        assert np.array_equal(got, ref.read_intrinsics(fn))
        assert np.array_equal(ext, ref.campose_to_extrinsic(poses))


def test_load_checkpoint_backfills_missing_keys(tmp_path):
    import modeling
    from stnerf_b200 import checkpoint_io as io
    case = C.CASES["syn_L2_64_128"]
    sd = C.state_dict_for(case)
    partial = {k: v for k, v in sd.items() if not k.startswith("time_deform_nets.1.")}
    path = os.path.join(str(tmp_path), "layered_rfnr_checkpoint_1.pt")
    torch.save({"model": partial}, path)
    m = modeling.build_layered_model(make_cfg(2, 64, 128, True))
    fresh = m.state_dict()
    missing = io.load_checkpoint(m, io.get_iteration_path(str(tmp_path)))
    assert sorted(missing) == sorted(k for k in sd if k.startswith("time_deform_nets.1."))
    now = m.state_dict()
    assert torch.equal(now["spacenets.0.stage1.0.weight"], sd["spacenets.0.stage1.0.weight"])
    assert torch.equal(now["time_deform_nets.1.motion_net.0.weight"], fresh["time_deform_nets.1.motion_net.0.weight"])
