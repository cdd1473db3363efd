"""Scene ingestion: stnerf_b200.scene_data against the reference's FrameLayerDataset outputs
(tests/golden/dataset.npz, produced by make_golden.py: run_dataset) on the same synthetic scene directory."""
import os

import numpy as np
import pytest
import torch

import cases as C
from stnerf_b200 import scene_data as SD

SP = C.DATASET_SPEC


@pytest.fixture(scope="module")
def scene_dir(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("scene"))
    C.write_synthetic_dataset(root)
    return root


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(os.path.dirname(C.__file__), "dataset.npz"))


@pytest.mark.parametrize("fmt", ["ascii", "le_f4", "be_f4", "le_f8_extra"])
def test_ply_reader_round_trip(tmp_path, fmt):
    pts = C.dataset_points(1, 4)
    p = str(tmp_path / "c.ply")
    C.write_ply(p, pts, fmt)
    got = SD.read_ply_points(p)
    assert got.dtype == np.float64 and got.shape == pts.shape
    if fmt == "ascii":                                  # 9 significant digits identify the fp32 value, not its fp64 widening
        assert np.array_equal(got.astype(np.float32), pts.astype(np.float32))
    else:
        assert np.array_equal(got, pts)                 # binary encodings carry the values exactly


def test_ply_reader_rejects_garbage(tmp_path):
    p = tmp_path / "x.ply"
    p.write_bytes(b"not a ply\n")
    with pytest.raises(ValueError):
        SD.read_ply_points(str(p))
    p.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nend_header\n0 0\n")
    with pytest.raises(ValueError):
        SD.read_ply_points(str(p))


@pytest.mark.parametrize("tag,fixed", [("auto", (-1.0, -1.0)), ("fixed", (0.5, 20.0))])
def test_frame_layer_data_matches_reference(scene_dir, golden, tag, fixed):
    for layer_id in range(SP["layer_num"] + 1):
        for frame_id in range(1 + SP["frame_offset"], SP["frame_offset"] + SP["frame_num"] + 1):
            d = SD.FrameLayerData(scene_dir, frame_id, layer_id, scale=SP["scale"], fixed_near=fixed[0], fixed_far=fixed[1],
                                  use_cache=False)
            k = "%s.l%d.f%d." % (tag, layer_id, frame_id)
            assert np.array_equal(np.asarray(d.bbox), golden[k + "bbox"]), k          # bit-exact: fp32 min/max * scale
            assert np.array_equal(np.asarray(d.center, dtype=np.float64), golden[k + "center"]), k
            np.testing.assert_allclose(np.asarray(d.near), golden[k + "near"], rtol=0, atol=2e-6, err_msg=k)
            np.testing.assert_allclose(np.asarray(d.far), golden[k + "far"], rtol=0, atol=2e-6, err_msg=k)
    assert np.array_equal(np.asarray(d.Ts), golden["Ts"]) and np.array_equal(np.asarray(d.Ks), golden["Ks"])
    assert tuple(golden["original_size"]) == d.get_original_size() == SP["original"]


def test_bbox_cache_round_trip(scene_dir):
    a = SD.FrameLayerData(scene_dir, 3, 1, scale=SP["scale"], use_cache=True)
    assert os.path.isfile(os.path.join(scene_dir, "bbox_tmp", "frame3", "layer1", "bbox.pt"))
    assert os.path.isfile(os.path.join(scene_dir, "near_far_tmp", "frame3", "layer1", "near.pt"))
    b = SD.FrameLayerData(scene_dir, 3, 1, scale=SP["scale"], use_cache=True)         # second time: from the cache files
    assert b.pointcloud is None
    assert torch.equal(a.bbox, b.bbox) and torch.equal(a.near, b.near) and np.array_equal(a.center, b.center)
    import shutil
    shutil.rmtree(os.path.join(scene_dir, "bbox_tmp")); shutil.rmtree(os.path.join(scene_dir, "near_far_tmp"))


def test_render_dataset_assembly(scene_dir, golden):
    ds = SD.RenderDataset(scene_dir, SP["layer_num"], SP["frame_num"], SP["frame_offset"], size_test=SP["size_test"],
                          scale=SP["scale"], fixed_near=0.5, fixed_far=20.0, use_cache=False)
    F = SP["frame_num"] + SP["frame_offset"]
    assert ds.bboxes.shape == (F, SP["layer_num"], 8, 3)
    for layer_id in (1, 2):
        for frame_id in range(1 + SP["frame_offset"], F + 1):
            assert np.array_equal(np.asarray(ds.bboxes[frame_id - 1, layer_id - 1]), golden["fixed.l%d.f%d.bbox" % (layer_id, frame_id)][0])
    assert torch.count_nonzero(ds.bboxes[:SP["frame_offset"]]) == 0                   # frames before the offset stay zero (ray_dataset.py:228)
    assert np.array_equal(np.asarray(ds.bkgd_bbox), golden["fixed.l0.f3.bbox"])
    # ray_dataset.py:243-248: fx, fy, cx, cy scaled by W_test / W_orig (fp32 arithmetic on the fp32 K)
    r = SP["size_test"][0] / SP["original"][0]
    K0 = torch.from_numpy(golden["Ks"]).clone()
    for (i, j) in ((0, 0), (1, 1), (0, 2), (1, 2)):
        K0[:, i, j] = K0[:, i, j] * r
    assert torch.equal(ds.Ks, K0)
    assert (ds.width, ds.height) == SP["size_test"] and ds.camera_num == SP["cams"]
    assert ds.frame_ids([(0, 3), (2, 4.5)]) == [3.0, 0.0, 4.5]
    assert ds.K[2, 2] == 1 and abs(float(ds.K[0, 0]) - float(golden["Ks"][0, 0, 0]) * 0.5) < 1e-4

    class M:
        def set_bkgd_bbox(self, b): self.bk = b
        def set_bboxes(self, b): self.bb = b
    m = ds.apply_to(M())
    assert m.bk is ds.bkgd_bbox and m.bb is ds.bboxes


def test_render_dataset_missing_cloud(tmp_path):
    root = str(tmp_path / "s")
    C.write_synthetic_dataset(root)
    os.remove(os.path.join(root, "frame4", "pointclouds", "2.ply"))
    with pytest.raises(FileNotFoundError):
        SD.RenderDataset(root, SP["layer_num"], SP["frame_num"], SP["frame_offset"], size_test=SP["size_test"], use_cache=False)


@pytest.mark.gpu
def test_dataset_rays_and_render(scene_dir):
    """Rays for a dataset pose come from the native generator (within 2e-6 of the oracle's generate_rays) and drive the
    model built from the dataset's boxes."""
    from oracle import stnerf_oracle as O
    from stnerf_b200.config import make_cfg
    import modeling
    ds = SD.RenderDataset(scene_dir, SP["layer_num"], SP["frame_num"], SP["frame_offset"], size_test=SP["size_test"],
                          scale=SP["scale"], fixed_near=0.5, fixed_far=20.0, use_cache=False)
    pair = [(0, 3), (1, 3), (2, 4)]
    rays, labels, bboxes, nf = ds.get_rays_by_pose_and_K(ds.poses[1], ds.Ks[1], pair)
    ref = O.generate_rays(ds.Ks[1], ds.poses[1], ds.height, ds.width)
    assert rays.shape == (ds.height * ds.width, 9)
    assert float((rays[:, :6].cpu() - ref).abs().max()) <= 2e-6 and rays[0, 6:].tolist() == [3.0, 3.0, 4.0]   # same bar as test_gpu_stages
    m = modeling.build_layered_model(make_cfg(2, 16, 16, True, "exact"))
    m.load_state_dict(O.synthetic_state_dict(2, True, seed=3))
    ds.apply_to(m)
    with torch.no_grad():
        out = m(rays, None, None, density_threshold=0.0, bkgd_density_threshold=0.0)
    assert out[0][0].shape == (rays.shape[0], 3) and bool(torch.isfinite(out[0][0]).all())
    assert int(out[4][1].sum()) > 0 and int(out[4][2].sum()) > 0        # both performer boxes are seen from this camera
