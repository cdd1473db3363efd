"""Scene ingestion for the renderer: camera files, per-layer point clouds -> bounding boxes, render rays.

Host-side mirror of what the reference's render data path reads from a captured scene, without open3d / torchvision:

    FrameLayerData   <- FrameLayerDataset.__init__      data/datasets/frame_dataset.py:94-247
                        pose/RT_c2w.txt + pose/K.txt (:124-132), translation scale (:128), `<layer>.ply` lookup order
                        (:109-115, :146-150), axis-aligned box from the scaled points in the corner order the ray/box test
                        expects (:169-191; corner 0 = min, corner 6 = max), `bbox_tmp/frame<f>/layer<l>/{center,bbox}.pt`
                        cache (:152-165, :193-206), camera-space near/far from the points or the fixed pair (:209-245)
    RenderDataset    <- Ray_Dataset_Render               data/datasets/ray_dataset.py:212-300
                        `bboxes (frame_num+frame_offset, layer_num, 8, 3)` indexed [frame_id-1, layer_id-1] (:228-237), poses,
                        intrinsics rescaled to the render width (:243-248), `get_rays_by_pose_and_K` / `get_rays_by_pose`
                        (:268-318) with the rays generated on the device by the native ray generator instead of on the host

The point clouds are read by a small PLY reader (`read_ply_points`: ascii, binary_little_endian, binary_big_endian; the
`vertex` element's x/y/z of any scalar type), which replaces `o3d.io.read_point_cloud(...).points`.
"""
from __future__ import annotations

import os
import struct
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .checkpoint_io import campose_to_extrinsic, read_intrinsics

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
    "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


def read_ply_points(path: str) -> np.ndarray:
    """Vertex positions of a PLY file as float64 (N,3) -- what `np.asarray(o3d.io.read_point_cloud(path).points)` holds."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("%s: not a PLY file" % path)
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError("%s: PLY header not terminated" % path)
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(cur)
            elif tok[0] == "property":
                if tok[1] == "list":
                    cur["props"].append(("list", tok[2], tok[3], tok[4]))
                else:
                    cur["props"].append((tok[1], tok[2]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise ValueError("%s: unsupported PLY format %r" % (path, fmt))
        for el in elements:
            scalar = all(p[0] != "list" for p in el["props"])
            if el["name"] != "vertex":
                if fmt == "ascii":
                    for _ in range(el["count"]):
                        f.readline()
                    continue
                if scalar:
                    f.seek(el["count"] * sum(np.dtype(_PLY_TYPES[p[0]]).itemsize for p in el["props"]), os.SEEK_CUR)
                    continue
                end = "<" if fmt == "binary_little_endian" else ">"
                for _ in range(el["count"]):            # list properties before the vertices: walk them
                    for p in el["props"]:
                        if p[0] == "list":
                            cnt_t, it_t = np.dtype(_PLY_TYPES[p[1]]), np.dtype(_PLY_TYPES[p[2]])
                            (k,) = struct.unpack(end + cnt_t.char, f.read(cnt_t.itemsize))
                            f.seek(int(k) * it_t.itemsize, os.SEEK_CUR)
                        else:
                            f.seek(np.dtype(_PLY_TYPES[p[0]]).itemsize, os.SEEK_CUR)
                continue
            if not scalar:
                raise ValueError("%s: list property in the vertex element" % path)
            names = [p[1] for p in el["props"]]
            if not all(a in names for a in "xyz"):
                raise ValueError("%s: vertex element lacks x/y/z" % path)
            n = el["count"]
            if fmt == "ascii":
                rows = np.loadtxt(f, dtype=np.float64, max_rows=n, ndmin=2) if n else np.zeros((0, len(names)))
                return np.ascontiguousarray(rows[:, [names.index(a) for a in "xyz"]], dtype=np.float64)
            end = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(p[1], end + _PLY_TYPES[p[0]]) for p in el["props"]])
            rec = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
            return np.stack([rec[a].astype(np.float64) for a in "xyz"], axis=1)
    raise ValueError("%s: no vertex element" % path)


def bbox_from_points(xyz: np.ndarray, scale: float = 1.0) -> Tuple[torch.Tensor, np.ndarray, torch.Tensor]:
    """frame_dataset.py:171-191: fp32 points * scale -> (bbox (1,8,3), center (3,), scaled points)."""
    pts = torch.Tensor(np.asarray(xyz)) * scale
    mx, mn = torch.max(pts, dim=0)[0], torch.min(pts, dim=0)[0]
    pad = (mx - mn) * 0.0                                           # the reference's margin factor is 0.0 (:180)
    mx, mn = mx + pad, mn - pad
    bbox = torch.Tensor([[mn[0], mn[1], mn[2]], [mx[0], mn[1], mn[2]], [mx[0], mx[1], mn[2]], [mn[0], mx[1], mn[2]],
                         [mn[0], mn[1], mx[2]], [mx[0], mn[1], mx[2]], [mx[0], mx[1], mx[2]], [mn[0], mx[1], mx[2]]])
    center = np.array([(mn[0] + mx[0]) / 2, (mn[1] + mx[1]) / 2, (mn[2] + mx[2]) / 2])
    return bbox.reshape(1, 8, 3), center, pts


class FrameLayerData:
    """One (frame, layer) of a captured scene: cameras, bounding box, near/far.  See the module docstring for the mapping."""

    def __init__(self, dataset_path: str, frame_id: int, layer_id: int, scale: float = 1.0, fixed_near: float = -1.0,
                 fixed_far: float = -1.0, camera_num: int = 0, use_cache: bool = True):
        self.frame_id, self.layer_id = frame_id, layer_id
        frame_dir = os.path.join(dataset_path, "frame" + str(frame_id))
        self.image_path = os.path.join(frame_dir, "images")
        if layer_id != 0:
            cloud_dirs = [os.path.join(frame_dir, "pointclouds"), os.path.join(dataset_path, "background")]
        else:
            cloud_dirs = [os.path.join(dataset_path, "background"), "None"]
        pose_path = os.path.join(dataset_path, "pose")
        self.Ts = torch.Tensor(campose_to_extrinsic(np.loadtxt(os.path.join(pose_path, "RT_c2w.txt"), ndmin=2)))
        self.Ts[:, 0:3, 3] = self.Ts[:, 0:3, 3] * scale
        self.Ks = torch.Tensor(read_intrinsics(os.path.join(pose_path, "K.txt")))
        self.cam_num = self.Ts.shape[0] if camera_num == 0 else camera_num

        name = os.path.join(cloud_dirs[0], "%d.ply" % layer_id)
        if not os.path.exists(name):
            name = os.path.join(cloud_dirs[1], "%d.ply" % layer_id)
        cache = os.path.join(dataset_path, "bbox_tmp", "frame" + str(frame_id), "layer" + str(layer_id))
        cached = os.path.exists(os.path.join(cache, "center.pt"))
        self.pointcloud = None
        self.bbox, self.center = None, torch.Tensor([0, 0, 0])
        if cached and use_cache:
            self.center = torch.load(os.path.join(cache, "center.pt"), weights_only=False)
            self.bbox = torch.load(os.path.join(cache, "bbox.pt"), weights_only=False)
        elif os.path.exists(name):
            self.bbox, self.center, self.pointcloud = bbox_from_points(read_ply_points(name), scale)
            if use_cache:
                os.makedirs(cache, exist_ok=True)
                torch.save(self.center, os.path.join(cache, "center.pt"))
                torch.save(self.bbox, os.path.join(cache, "bbox.pt"))

        if fixed_near == -1.0 and fixed_far == -1.0:
            nf = os.path.join(dataset_path, "near_far_tmp", "frame" + str(frame_id), "layer" + str(layer_id))
            if use_cache and os.path.exists(os.path.join(nf, "near.pt")):
                self.near = torch.load(os.path.join(nf, "near.pt"), weights_only=False)
                self.far = torch.load(os.path.join(nf, "far.pt"), weights_only=False)
            else:
                if self.pointcloud is None:
                    self.pointcloud = torch.Tensor(read_ply_points(name)) * scale
                inv = torch.inverse(self.Ts)                                       # world -> camera
                z = self.pointcloud @ inv[:, 2, :3].T + inv[:, 2, 3][None]         # (N, M) camera-space depth
                self.near, self.far = z.min(dim=0)[0], z.max(dim=0)[0]
                if use_cache:
                    os.makedirs(nf, exist_ok=True)
                    torch.save(self.near, os.path.join(nf, "near.pt"))
                    torch.save(self.far, os.path.join(nf, "far.pt"))
        else:
            self.near = torch.ones(self.Ts.shape[0]) * fixed_near
            self.far = torch.ones(self.Ts.shape[0]) * fixed_far

    def __len__(self):
        return self.cam_num

    def get_original_size(self) -> Tuple[int, int]:
        """(width, height) of the first captured image (:319-329); needs Pillow only when images are present."""
        for nm in ("%03d.png" % 0, "%d.png" % 0):
            p = os.path.join(self.image_path, nm)
            if os.path.exists(p):
                from PIL import Image
                with Image.open(p) as im:
                    return im.size
        raise FileNotFoundError("no image 000.png / 0.png under %s" % self.image_path)


class RenderDataset:
    """What the free-viewpoint renderer needs from a scene directory (`Ray_Dataset_Render`): `bboxes`, `poses`, `Ks`,
    `camera_num`, `height`/`width`, and rays for a pose.  `size_test = (W, H)` as `cfg.INPUT.SIZE_TEST`; `original_size`
    (W, H) may be given when the scene ships no images."""

    def __init__(self, dataset_path: str, layer_num: int, frame_num: int, frame_offset: int = 0,
                 size_test: Sequence[int] = (1920, 1080), scale: float = 1.0, fixed_near: float = -1.0,
                 fixed_far: float = -1.0, camera_num: int = 0, original_size: Optional[Sequence[int]] = None,
                 use_cache: bool = True, use_time: bool = True):
        self.layer_num, self.frame_num, self.frame_offset = layer_num, frame_num, frame_offset
        self.use_time = use_time
        self.datasets: List[List[FrameLayerData]] = []
        self.bboxes = torch.zeros(frame_num + frame_offset, layer_num, 8, 3)
        for layer_id in range(layer_num + 1):
            row = []
            for frame_id in range(1 + frame_offset, frame_offset + frame_num + 1):
                d = FrameLayerData(dataset_path, frame_id, layer_id, scale, fixed_near, fixed_far, camera_num, use_cache)
                row.append(d)
                if layer_id != 0:
                    if d.bbox is None:
                        raise FileNotFoundError("no point cloud / cached bbox for layer %d frame %d" % (layer_id, frame_id))
                    self.bboxes[frame_id - 1, layer_id - 1] = d.bbox
            self.datasets.append(row)
        first = self.datasets[0][0]
        self.bkgd_bbox = first.bbox
        self.camera_num = first.cam_num
        self.poses = first.Ts
        col, row_px = original_size if original_size is not None else first.get_original_size()
        self.Ks = first.Ks.clone()
        r = size_test[0] / col                                                      # ray_dataset.py:243-248
        self.Ks[:, 0, 0] *= r; self.Ks[:, 1, 1] *= r; self.Ks[:, 0, 2] *= r; self.Ks[:, 1, 2] *= r
        # The default intrinsic / image size come from `get_data(0)` through the test-time transform, whose deterministic
        # branch (data/transforms/random_transforms.py:57-160 with range = rotation = 0, ratio = 1) is K * (H_test / H_orig)
        # with K[2,2] = 1 and an output of exactly SIZE_TEST.
        self.width, self.height = int(size_test[0]), int(size_test[1])
        s = self.height / row_px
        self.K = first.Ks[0].clone() * s
        self.K[2, 2] = 1
        self.near_far = torch.Tensor([fixed_near, fixed_far]).reshape(1, 2)

    def frame_ids(self, layer_frame_pair) -> List[float]:
        ids = [0.0] * (self.layer_num + 1)
        for layer_id, frame_id in layer_frame_pair:
            ids[layer_id] = float(frame_id)
        return ids

    def get_rays_by_pose_and_K(self, T, K, layer_frame_pair, device="cuda"):
        """(rays (H*W, 6 [+ layer_num+1]), labels, bboxes, near_fars) as ray_dataset.py:268-293; the rays are produced on
        `device` by the native generator (within 2e-6 of utils/render_helpers.py:42-126, tests/test_gpu_stages.py)."""
        from . import ops
        ids = self.frame_ids(layer_frame_pair) if self.use_time else None
        rays = ops.generate_rays(torch.as_tensor(K, dtype=torch.float32), torch.as_tensor(np.asarray(T), dtype=torch.float32),
                                 self.height, self.width, frame_ids=ids, device=device)
        n = rays.shape[0]
        return rays, torch.zeros(n), torch.zeros(n, 8, 3), self.near_far.repeat(n, 1)

    def get_rays_by_pose(self, T, layer_frame_pair, device="cuda"):
        return self.get_rays_by_pose_and_K(T, self.K, layer_frame_pair, device)

    def apply_to(self, model):
        """render/layered_neural_renderer.py:107-108."""
        model.set_bkgd_bbox(self.bkgd_bbox)
        model.set_bboxes(self.bboxes)
        return model
