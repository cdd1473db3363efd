"""`render.LayeredNeuralRenderer` (the object the reference's demos drive) assembled over the native path: construction from a
scene directory + checkpoint directory, path / timeline methods, and -- on the GPU -- `render_path` with the reference's
output layout, frame for frame equal to calling the model on the same rays."""
import os

import numpy as np
import pytest
import torch

import cases as C
from stnerf_b200.config import make_render_cfg
from stnerf_b200.synthetic import synthetic_state_dict

SP = C.DATASET_SPEC


@pytest.fixture(scope="module")
def dirs(tmp_path_factory):
    scene = str(tmp_path_factory.mktemp("scene"))
    out = str(tmp_path_factory.mktemp("outputs"))
    C.write_synthetic_dataset(scene)
    torch.save({"model": synthetic_state_dict(SP["layer_num"], True, seed=4)}, os.path.join(out, "layered_rfnr_checkpoint_3.pt"))
    torch.save({"model": synthetic_state_dict(SP["layer_num"], True, seed=5)}, os.path.join(out, "layered_rfnr_checkpoint_12.pt"))
    return scene, out


def _renderer(dirs, **kw):
    import render
    scene, out = dirs
    cfg = make_render_cfg(out, scene, SP["layer_num"], SP["frame_num"], SP["size_test"], n1=16, n2=24,
                          frame_offset=SP["frame_offset"], scale=SP["scale"], fixed_near=0.5, fixed_far=20.0)
    return render.LayeredNeuralRenderer(cfg, **kw)


def test_construction_and_path_methods(dirs):
    r = _renderer(dirs, s_shift=[[[0, 0, 0]] * 3, [[0, 1, 0]] * 3])
    # newest checkpoint of OUTPUT_DIR (get_iteration_path), boxes from the scene directory
    want = synthetic_state_dict(SP["layer_num"], True, seed=5)
    assert torch.equal(r.model.state_dict()["spacenets.0.stage1.0.weight"], want["spacenets.0.stage1.0.weight"])
    assert tuple(r.model.bboxes.shape) == (SP["frame_num"] + SP["frame_offset"], SP["layer_num"], 8, 3)
    assert r.model.shift == [[0, 0, 0]] * 3                      # s_shift[0] seeds the model (:28-29)
    assert (r.height, r.width) == (SP["size_test"][1], SP["size_test"][0]) and r.camera_num == SP["cams"]
    assert r.min_frame == [1 + SP["frame_offset"]] * 3 and r.max_frame == [SP["frame_num"] + SP["frame_offset"]] * 3
    r.set_smooth_path_poses(5, around=False)
    assert len(r.poses) == 5 and len(r.Ks) == 5 and len(r.layer_frame_pairs) == 6 and len(r.s_shift_frame) == 5
    r.hide_layer(1)
    assert not r.is_shown_layer(1) and not r.model.is_shown_layer(1)
    r.show_layer(1)
    r.set_path_gt_poses()
    assert len(r.poses) == 5 + SP["cams"] and len(r.Ks) == 5 + SP["cams"]
    r.set_path_fixed_gt_poses(2, num=4)
    assert len(r.poses) == 9 + SP["cams"] and torch.equal(torch.as_tensor(r.poses[-1]), r.dataset.poses[2])
    before = r.gt_poses.copy()
    c = np.asarray(r.get_center_frame_layer(0, 1), dtype=np.float32)
    r.zoom_in(1, 0, 2.0)
    assert np.allclose(r.gt_poses[:, :3, 3], c + 0.5 * (before[:, :3, 3] - c), atol=1e-6)
    r.set_near(4.0); r.set_fps(30); r.set_save_dir("edit")
    assert r.model.near == 4.0 and r.fps == 30 and r.dir_name == "edit"


def test_missing_checkpoint_is_an_error(dirs, tmp_path):
    import render
    cfg = make_render_cfg(str(tmp_path), dirs[0], SP["layer_num"], SP["frame_num"], SP["size_test"], frame_offset=SP["frame_offset"])
    with pytest.raises(FileNotFoundError):
        render.LayeredNeuralRenderer(cfg)


@pytest.mark.gpu
def test_render_path_layout_and_pixels(dirs):
    from PIL import Image
    r = _renderer(dirs, shift=[[0, 0, 0], [0, 0.2, 0], [0, -0.2, 0]])
    r.set_save_dir("shift")
    r.set_smooth_path_poses(3, around=False)
    r.hide_layer(2)
    r.model.seed = 100
    r.render_path(inverse_y_axis=False, density_threshold=0, auto_save=True)
    base = os.path.join(r.output_dir, "shift", "video_0")
    assert sorted(os.listdir(base)) == ["0", "1", "mixed"]                       # hidden layer 2 gets no folder (:470-471)
    assert sorted(os.listdir(os.path.join(base, "mixed"))) == ["Ks", "color", "depth", "poses"]
    assert sorted(os.listdir(os.path.join(base, "mixed", "color"))) == ["0.jpg", "1.jpg", "2.jpg"]
    assert len(r.images) == 3 and tuple(r.images[0].shape) == (r.height, r.width, 3) and tuple(r.depths[0].shape) == (r.height, r.width, 1)
    png = np.asarray(Image.open(os.path.join(base, "mixed", "depth", "1.png")))
    want = (np.clip(r.depths[1].numpy()[..., 0], 0, 1) * 255 + 0.5).astype(np.uint8)
    assert png.shape == want.shape and np.array_equal(png, want)
    # frame 1 again through render_pose, and through the model on the dataset's rays: same pixels (same Philox seed)
    r.model.seed = 101
    color, depth, color_layer, depth_layer = r.render_pose(r.poses[1], r.Ks[1], r.layer_frame_pairs[1], 0, 0)
    assert torch.equal(color.cpu(), r.images[1])
    # depth = raw / far: the device path divides on the GPU (reciprocal multiply), the path generator on the host -> 1 ulp
    assert float((depth.cpu() - r.depths[1]).abs().max()) <= 1e-7
    rays, labels, bboxes, nf = r.dataset.get_rays_by_pose_and_K(r.poses[1], r.Ks[1], r.layer_frame_pairs[1])
    r.model.seed = 101
    import utils
    with torch.no_grad():
        stage2, stage1, stage2_layer, stage1_layer, _ = utils.layered_batchify_ray(r.model, rays, labels.cuda(), bboxes.cuda(),
                                                                                    density_threshold=0, bkgd_density_threshold=0)
    assert float((stage2[0].reshape(r.height, r.width, 3).cpu() - r.images[1]).abs().max()) < 2e-6
    r.save_video()                                                # without imageio: counts the take, leaves the frames
    assert r.save_count == 1
    # the walking variant: flat folders, every layer written, plus the layer-2-over-background composite "02" (:550-618)
    r.show_layer(2)
    r.render_path_walking(auto_save=True)
    for leaf in ("mixed", "0", "1", "2", "02"):
        assert sorted(os.listdir(os.path.join(r.output_dir, leaf, "color"))) == ["0.jpg", "1.jpg", "2.jpg"], leaf
    assert not os.path.exists(os.path.join(r.output_dir, "02", "depth")) and len(r.images) == 3


# demo/taekwondo_demo.py:39-72: three edit sessions (origin / shift / scale), each = retime two performers by key frames,
# a 101-pose smooth path, render_path with auto_save, save_video.  Executed VERBATIM from the reference's demo script when
# a reference checkout (or the archive of oracle/stash_reference.py) is available, restated otherwise.
DEMO_SESSIONS = '''
key_frames_layer_1 = [21,49,74,87]
key_frames_layer_2 = [13,42,80,90]
key_frames = [20,50,74,85]
density_threshold = 0
inverse_y_axis = False
for kw, name in (({}, 'origin'), ({'shift': [[0,0,0],[0,2,0],[0,-2,0]]}, 'shift'), ({'scale': [1,0.75,1.5]}, 'scale')):
    neural_renderer = LayeredNeuralRenderer(cfg, **kw)
    neural_renderer.set_save_dir(name)
    neural_renderer.retime_by_key_frames(1, key_frames_layer_1, key_frames)
    neural_renderer.retime_by_key_frames(2, key_frames_layer_2, key_frames)
    neural_renderer.set_fps(25)
    neural_renderer.set_smooth_path_poses(101, around=False)
    neural_renderer.render_path(inverse_y_axis,density_threshold,auto_save=True)
    neural_renderer.save_video()
'''


@pytest.mark.gpu
def test_taekwondo_demo_call_sequence(tmp_path):
    import sys
    import render
    sys.path.insert(0, C.ROOT)
    from oracle import stash_reference
    spec = dict(SP, frame_num=101, frame_offset=0, size_test=(48, 27), original=(96, 54))
    scene, out = str(tmp_path / "scene"), str(tmp_path / "outputs")
    os.makedirs(out)
    C.write_synthetic_dataset(scene, spec=spec)
    torch.save({"model": synthetic_state_dict(2, True, seed=9)}, os.path.join(out, "layered_rfnr_checkpoint_1.pt"))
    cfg = make_render_cfg(out, scene, 2, spec["frame_num"], spec["size_test"], n1=12, n2=20, frame_offset=0,
                          scale=spec["scale"], fixed_near=0.5, fixed_far=20.0)
    ref = stash_reference.reference_root()
    block, verbatim = DEMO_SESSIONS, False
    if ref is not None:
        lines = open(os.path.join(ref, "demo", "taekwondo_demo.py")).read().split("\n")
        block, verbatim = "\n".join(lines[38:72]), True                  # demo/taekwondo_demo.py:39-72
        assert "LayeredNeuralRenderer(cfg, scale=[1,0.75,1.5])" in block and "cfg.merge_from_file" not in block
    scope = {"cfg": cfg, "LayeredNeuralRenderer": render.LayeredNeuralRenderer}
    exec(block, scope)
    r = scope["neural_renderer"]                                         # the last session ('scale')
    assert r.model.scale == [1, 0.75, 1.5] and len(r.poses) == 101 and len(r.images) == 101 and r.save_count == 1
    root = os.path.join(out, "rendered")
    assert sorted(os.listdir(root)) == ["origin", "scale", "shift"], (verbatim, os.listdir(root))
    for name in ("origin", "shift", "scale"):
        base = os.path.join(root, name, "video_0")
        assert sorted(os.listdir(base)) == ["0", "1", "2", "mixed"]
        assert len(os.listdir(os.path.join(base, "mixed", "color"))) == 101 and len(os.listdir(os.path.join(base, "2", "depth"))) == 101
    # the retimed timelines reached the renderer: layer 1 shows frame 21 where the new timeline says 20, ... (:46-47)
    pairs = dict((int(a), float(b)) for a, b in r.layer_frame_pairs[20])
    assert set(pairs) == {0, 1, 2}
    # frame 50 of the scale session again through render_pose: same pixels as the saved take (same seed)
    r.model.seed = r.model.seed - (101 - 50)
    color, depth, _, _ = r.render_pose(r.poses[50], r.Ks[50], r.layer_frame_pairs[50], 0, 0)
    assert torch.equal(color.cpu(), r.images[50])
