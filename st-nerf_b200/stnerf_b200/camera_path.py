"""Camera-path and retiming front end of the free-viewpoint renderer: host-side scheduling only.

Mirrors the path / timeline methods of `LayeredNeuralRenderer` (render/layered_neural_renderer.py) without its
dataset, file and video IO, so the demo edit sessions (demo/taekwondo_demo.py:39-72, demo/walking_demo.py:40-68) can be
scripted on top of `PoseRenderer`:

    set_frame_duration / set_pose_duration   :672-683     set_smooth_path_poses   :230-319
    retime_by_key_frames                     :495-544     invert_poses            :685-687
    hide_layer / show_layer / is_shown_layer :653-664     load_path_poses         :321-337

State and names follow the reference (`poses`, `Ks`, `layer_frame_pairs`, `s_shift_frame`, ...).  Inputs that the
reference reads from its dataset object are constructor arguments here: the ground-truth camera poses `gt_poses (M,4,4)`,
intrinsics `gt_Ks (M,3,3)`, `frame_num`, `frame_offset`, `layer_num`.
Arithmetic is numpy/scipy exactly as in the reference (Slerp of the rotations, cubic B-spline through the camera
centres, linear interpolation of K); parity is pinned in tests/test_camera_path.py against vectors produced by executing
the reference's own method bodies (tests/golden/make_golden.py: run_camera_path).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
from scipy.interpolate import splev, splprep
from scipy.spatial.transform import Rotation as R
from scipy.spatial.transform import Slerp


class CameraPath:
    def __init__(self, gt_poses, gt_Ks, layer_num: int, frame_num: int, frame_offset: int = 0, s_shift=None,
                 s_scale=None, s_alpha=None):
        self.gt_poses = np.asarray(gt_poses, dtype=np.float32)
        self.gt_Ks = [np.asarray(k) for k in gt_Ks]
        self.layer_num = int(layer_num)
        self.camera_num = self.gt_poses.shape[0]
        self.min_camera_id, self.max_camera_id = 0, self.camera_num - 1                     # :73-74
        self.min_frame = [1 + frame_offset for _ in range(layer_num + 1)]                   # :61
        self.max_frame = [frame_num + frame_offset for _ in range(layer_num + 1)]           # :62
        self.display_layers = {i: 1 for i in range(layer_num + 1)}                          # :47-51
        self.s_shift, self.s_scale, self.s_alpha = s_shift, s_scale, s_alpha
        self.poses: List[np.ndarray] = []
        self.Ks: List[np.ndarray] = []
        self.layer_frame_pairs: List[list] = []
        self.s_shift_frame = self.s_scale_frame = self.s_alpha_frame = None

    # ---- small state setters ---------------------------------------------------------------------------------
    def hide_layer(self, layer_id): self.display_layers[layer_id] = 0
    def show_layer(self, layer_id): self.display_layers[layer_id] = 1
    def is_shown_layer(self, layer_id): return self.display_layers[layer_id] == 1

    def set_frame_duration(self, min_frame, max_frame, layer_id=-1):
        ids = range(self.layer_num + 1) if layer_id == -1 else [layer_id]
        for i in ids:
            self.min_frame[i], self.max_frame[i] = min_frame, max_frame

    def set_pose_duration(self, min_camera_id, max_camera_id):
        self.min_camera_id, self.max_camera_id = min_camera_id, max_camera_id

    def invert_poses(self):
        self.poses.reverse()
        self.Ks.reverse()

    # ---- layer / frame schedule shared by every path constructor (:306-318, :160-168, :178-186) ----------------
    def _append_layer_frame_pairs(self, n_poses: int, smooth_time: bool = False):
        for idx in range(n_poses + 1):                 # the reference appends len(poses)+1 entries
            pair = []
            for layer_id in range(self.layer_num + 1):
                if self.is_shown_layer(layer_id):
                    span = (self.max_frame[layer_id] - self.min_frame[layer_id]) / n_poses * idx
                    frame_id = (span if smooth_time else int(span)) + self.min_frame[layer_id]
                    pair.append((layer_id, frame_id))
            self.layer_frame_pairs.append(pair)

    # ---- :230-319 ----------------------------------------------------------------------------------------------
    def set_smooth_path_poses(self, step_num: int, around: bool = False, smooth_time: bool = False):
        if self.s_shift is not None:
            s0, s1 = np.array(self.s_shift[0]), np.array(self.s_shift[1])
            shift_step = (s1 - s0) / (step_num - 1)
            self.s_shift_frame = []
        if self.s_alpha is not None:
            a0, a1 = self.s_alpha[0], self.s_alpha[1]
            alpha_step = (a1 - a0) / (step_num - 1)
            self.s_alpha_frame = []
        lo, hi = self.min_camera_id, self.max_camera_id
        Rs = self.gt_poses[lo:hi + 1, :3, :3]
        Ts = self.gt_poses[lo:hi + 1, :3, 3]
        key_frames = list(range(lo, hi + 1))
        if not around:                                 # only the first and the last camera orientation
            Rs = np.array([Rs[0], Rs[-1]])
            key_frames = [lo, hi]
        interp_frames = [(i * (hi - lo) / (step_num - 1) + lo) for i in range(step_num)]
        interp_Rs = Slerp(key_frames, R.from_matrix(Rs))(interp_frames).as_matrix()
        tck, _ = splprep([Ts[:, 0], Ts[:, 1], Ts[:, 2]])
        new_points = np.stack(splev([i / (step_num - 1) for i in range(step_num)], tck), axis=1)
        K0, K1 = self.gt_Ks[lo], self.gt_Ks[hi]
        if self.s_scale is not None:
            c0, c1 = np.array(self.s_scale[0]), np.array(self.s_scale[1])
            scale_step = (c1 - c0) / (step_num - 1)
            self.s_scale_frame = []
        poses = []
        for i in range(step_num):
            pose = np.zeros((4, 4))
            pose[:3, :3] = interp_Rs[i]
            pose[:3, 3] = new_points[i]
            pose[3, 3] = 1
            poses.append(pose)
            self.Ks.append((K1 - K0) * i / (step_num - 1) + K0)
            if self.s_scale is not None:
                self.s_scale_frame.append((c0 + i * scale_step).tolist())
            if self.s_shift is not None:
                self.s_shift_frame.append((s0 + i * shift_step).tolist())
            if self.s_alpha is not None:
                self.s_alpha_frame.append(a0 + i * alpha_step)
        self.poses = self.poses + poses
        self._append_layer_frame_pairs(len(poses), smooth_time)

    # ---- :321-337 ----------------------------------------------------------------------------------------------
    def load_path_poses(self, poses: Sequence[np.ndarray]):
        self.poses = list(poses)
        step_num = len(poses)
        K0, K1 = self.gt_Ks[self.min_camera_id], self.gt_Ks[self.max_camera_id - 1]
        for i in range(step_num):
            self.Ks.append((K1 - K0) * i / (step_num - 1) + K0)
        self._append_layer_frame_pairs(len(poses))

    # ---- :495-544 ----------------------------------------------------------------------------------------------
    def retime_by_key_frames(self, layer_id: int, key_frames_layer: Sequence[int], key_frames: Sequence[int]):
        assert len(key_frames_layer) == len(key_frames)
        for i in range(len(self.layer_frame_pairs)):
            for j in range(len(self.layer_frame_pairs[i])):
                layer, frame = self.layer_frame_pairs[i][j]
                if layer != layer_id:
                    continue
                idx_start, idx_end, weight = -1, -1, 0
                for idx in range(len(key_frames)):
                    if frame <= key_frames[idx]:
                        idx_end, idx_start = idx, idx - 1
                        end = key_frames[idx]
                        start = self.min_frame[layer] if idx == 0 else key_frames[idx - 1]
                        weight = (frame - start) / (end - start)
                        break
                if idx_start == -1 and idx_end == 0:
                    weight = (frame - self.min_frame[layer]) / (key_frames[0] - self.min_frame[layer])
                    new_start, new_end = self.min_frame[layer], key_frames_layer[0]
                elif idx_start >= -1 and idx_end != -1:
                    new_start, new_end = key_frames_layer[idx_start], key_frames_layer[idx_start + 1]
                elif idx_start == -1 and idx_end == -1:
                    weight = (frame - key_frames[-1]) / (self.max_frame[layer] - key_frames[-1])
                    new_start, new_end = key_frames_layer[-1], self.max_frame[layer]
                else:
                    raise ValueError("Undefined branch: start idx %d, end idx %d" % (idx_start, idx_end))   # ref: exit(-1)
                self.layer_frame_pairs[i][j] = (layer, round(weight * (new_end - new_start) + new_start))

    # ---- driving PoseRenderer (render_path :401-488 without the file IO) -----------------------------------------------
    def per_frame_state(self):
        """Callback for PoseRenderer.render_path: the per-frame edit state the reference pushes into the model (:435-440)."""
        def apply(idx, model):
            if self.s_shift is not None:
                model.shift = self.s_shift_frame[idx]
            if self.s_scale is not None:
                model.scale = self.s_scale_frame[idx]
            if self.s_alpha is not None:
                model.alpha = self.s_alpha_frame[idx]
        return apply
