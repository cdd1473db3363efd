"""CPU tests of the host-side mirror of the reference interface (no device needed)."""
import types

import numpy as np
import pytest
import torch

import cases as C
from oracle import stnerf_oracle as O
from tests_support import make_cfg


def _model(case):
    import modeling
    m = modeling.build_layered_model(make_cfg(case["L"], case["n1"], case["n2"], case["space_time"]), 0,
                                     case.get("scale"), case.get("shift"))
    bkgd, frames = C.boxes_for(case)
    m.set_bkgd_bbox(bkgd); m.set_bboxes(frames)
    m.near, m.alpha = case.get("near", 0.0), case.get("alpha", 1.0)
    for i in case.get("hidden", []):
        m.hide_layer(i)
    return m


@pytest.mark.parametrize("name", list(C.CASES))
def test_scene_prologue_matches_oracle(name):
    """forward()'s prologue (layered_rfrender.py:190-242): edited boxes, pivot, edit flags."""
    case = C.CASES[name]
    m = _model(case)
    sc = m._resolve_scene(torch.tensor(case["frame_ids"]), case["thr"][0], case["thr"][1])
    want = C.scene_for(case)
    l = case["L"] + 1
    got_min = np.array([[sc.bmin[i][a] for a in range(3)] for i in range(l)], np.float32)
    got_max = np.array([[sc.bmax[i][a] for a in range(3)] for i in range(l)], np.float32)
    assert np.array_equal(got_min, want["bmin"].numpy()) and np.array_equal(got_max, want["bmax"].numpy())
    if case.get("scale") is not None:
        assert np.array_equal(np.array(list(sc.pivot), np.float32), want["pivot"].numpy())
    assert [sc.shown[i] for i in range(l)] == [1 if s else 0 for s in want["shown"]]
    assert sc.apply_thresholds == 1 and sc.near_plane == np.float32(case.get("near", 0.0))


def test_edit_flag_quirks():
    """None shift entries skip the fine-pass scale as well (`continue` at layered_rfrender.py:468-469)."""
    case = dict(C.CASES["tkd_edit_frac"], shift=[[0, 0, 0], None, [0, -2, 0]])
    m = _model(case)
    sc = m._resolve_scene(torch.tensor(case["frame_ids"]), 0.0, 0.0)
    assert [sc.shift_on[i] for i in range(3)] == [1, 0, 1]
    assert [sc.scale_coarse_on[i] for i in range(3)] == [1, 1, 1]
    assert [sc.scale_fine_on[i] for i in range(3)] == [1, 0, 1]


def test_state_dict_surface():
    """Key names / shapes / order equal the shipped checkpoints; missing-key back-fill flow of
    render/layered_neural_renderer.py:109-117 works."""
    m = _model(C.CASES["tkd_64_128"])
    sd = m.state_dict()
    assert sum(v.numel() for v in sd.values()) == 2950942
    p = C.find_checkpoint("taekwondo")
    if p is not None:
        ck = torch.load(p, map_location="cpu")["model"]
        assert list(ck.keys()) == list(sd.keys())
        assert all(tuple(ck[k].shape) == tuple(sd[k].shape) for k in ck)
        partial = {k: v for k, v in ck.items() if not k.startswith("time_deform_nets.1.")}
        model_dict = m.state_dict()
        model_dict.update({k: v for k, v in partial.items() if k in model_dict})
        m.load_state_dict(model_dict)
        assert torch.equal(m.state_dict()["bkgd_spacenet.stage1.0.weight"], ck["bkgd_spacenet.stage1.0.weight"])
    with pytest.raises(RuntimeError):
        m.load_state_dict({"nope": torch.zeros(1)})


def test_weight_blob_order():
    from stnerf_b200.native import _blob, SPACENET_KEYS, MOTIONNET_KEYS
    sd = O.synthetic_state_dict(1, True, seed=3)
    b = _blob(sd, "spacenets.0.", SPACENET_KEYS)
    assert b.numel() == 466948
    assert torch.equal(b[:256 * 63].reshape(256, 63), sd["spacenets.0.stage1.0.weight"])
    assert torch.equal(b[-3:], sd["spacenets.0.rgb_net.3.bias"])
    assert _blob(sd, "bkgd_spacenet.", SPACENET_KEYS).numel() == 464260
    assert _blob(sd, "time_deform_nets.0.", MOTIONNET_KEYS).numel() == 77315


def test_batchify_threshold_quirk():
    """utils/batchify_rays.py:53-54: calls smaller than `chuncks` do not forward the thresholds."""
    import utils
    seen = {}

    def fake(rays, labels, bboxes, **kw):
        seen.clear(); seen.update(kw)
        return (1, 2, 3, 4, 5)

    utils.layered_batchify_ray(fake, torch.zeros(10, 9), None, None, density_threshold=20, bkgd_density_threshold=0.8)
    assert "density_threshold" not in seen
    utils.layered_batchify_ray(fake, torch.zeros(4000, 9), None, None, density_threshold=20, bkgd_density_threshold=0.8)
    assert seen["density_threshold"] == 20 and seen["bkgd_density_threshold"] == 0.8


def test_reference_names_import():
    import modeling, utils, layers, engine
    assert modeling.build_model is modeling.build_layered_model
    for n in ("Trigonometric_kernel", "sample_pdf", "ray_sampling", "batchify_ray", "layered_batchify_ray", "generate_rays"):
        assert hasattr(utils, n)
    for n in ("RaySamplePoint", "RaySamplePoint_Near_Far", "VolumeRenderer", "make_loss"):
        assert hasattr(layers, n)
    assert callable(engine.render)


def test_unsupported_configs_fail_loudly():
    import modeling
    cfg = make_cfg(2, 64, 128, True)
    cfg.MODEL.SAMPLE_METHOD = "NEAR_FAR"
    with pytest.raises(NotImplementedError):
        modeling.build_layered_model(cfg)
    m = _model(C.CASES["tkd_64_128"])
    with pytest.raises(Exception):
        m(torch.zeros(8, 9), None)            # CPU rays: the product path has no CPU fallback


def test_synthetic_generators_in_sync_with_oracle_copy():
    """stnerf_b200.synthetic (bench / examples inputs) and the oracle-side copy used by the tests must agree."""
    from stnerf_b200 import synthetic as S
    b1, f1 = S.synthetic_boxes(3); b2, f2 = O.synthetic_boxes(3)
    assert torch.equal(b1, b2) and torch.equal(f1, f2)
    for v in (0, 5):
        K1, T1 = S.synthetic_camera(v, 16, 1080, 1920); K2, T2 = O.synthetic_camera(v, 16, 1080, 1920)
        assert torch.equal(K1, K2) and torch.equal(T1, T2)
    s1, s2 = S.synthetic_state_dict(1, True, seed=4), O.synthetic_state_dict(1, True, seed=4)
    assert list(s1) == list(s2) and all(torch.equal(s1[k], s2[k]) for k in s1)
