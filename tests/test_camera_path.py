"""Camera-path / retiming front end against vectors produced by executing the reference's own
method bodies (tests/golden/make_golden.py: run_camera_path).  CPU only."""
import numpy as np
import pytest

import cases as C

GOLD = C.load_golden("camera_path")


@pytest.mark.parametrize("name", list(C.CAMERA_PATH_SCENARIOS))
def test_camera_path_matches_reference(name):
    from stnerf_b200.camera_path import CameraPath
    assert GOLD is not None, "tests/golden/camera_path.npz missing"
    sc = C.CAMERA_PATH_SCENARIOS[name]
    gt_poses, gt_Ks = C.camera_path_inputs()
    p = CameraPath(gt_poses.numpy(), [k.numpy() for k in gt_Ks], layer_num=2, frame_num=101, frame_offset=sc["offset"],
                   s_shift=sc.get("s_shift"), s_scale=sc.get("s_scale"), s_alpha=sc.get("s_alpha"))
    for i in sc.get("hidden", []):
        p.hide_layer(i)
    C.drive_camera_path(p, sc)
    np.testing.assert_allclose(np.stack(p.poses), GOLD[name + ".poses"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.stack([np.asarray(k, np.float64) for k in p.Ks]), GOLD[name + ".Ks"], rtol=0, atol=1e-9)
    pairs = np.array([[f for (_, f) in pair] for pair in p.layer_frame_pairs], dtype=np.float64)
    assert np.array_equal(pairs, GOLD[name + ".pairs"])
    if sc.get("s_shift") is not None:
        assert np.array_equal(np.array(p.s_shift_frame), GOLD[name + ".s_shift_frame"])
        assert np.allclose(np.array(p.s_alpha_frame), GOLD[name + ".s_alpha_frame"], rtol=0, atol=0)
    assert len(p.poses) == sc["steps"] and len(p.layer_frame_pairs) == sc["steps"] + 1


def test_per_frame_state_pushes_edits():
    from stnerf_b200.camera_path import CameraPath
    import types
    sc = C.CAMERA_PATH_SCENARIOS["around_smooth"]
    gt_poses, gt_Ks = C.camera_path_inputs()
    p = CameraPath(gt_poses.numpy(), [k.numpy() for k in gt_Ks], 2, 101, s_shift=sc["s_shift"], s_alpha=sc["s_alpha"])
    p.set_smooth_path_poses(5)
    model = types.SimpleNamespace(shift=None, scale=None, alpha=1)
    p.per_frame_state()(4, model)
    assert model.shift == [[0, 0, 0], [0, 2, 0], [0, -2, 0]] and abs(model.alpha - 0.25) < 1e-12 and model.scale is None
